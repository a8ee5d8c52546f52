#!/bin/bash
# Round-2 evidence run on ONE B200: bench (all legs) + reference arm, ncu launch list of one eager step, tensor-pipe counters of
# every tcgen05 launch of a step, GEMM DRAM traffic, ncu --set full summaries of the main kernels.  Usage: tools/gpu_final_r2.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-graph"
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
ncu --query-metrics 2>/dev/null | grep -i "tensor" > gpurun_out/${tag}_tensor_metrics_available.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,gpu__time_duration.sum --clock-control none -k regex:"gemm_tc_kernel|attn_tc_" -c 500 --csv --log-file gpurun_out/${tag}_tensor_pipe.csv $B > gpurun_out/${tag}_ncu_tensor_pipe.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc_kernel -c 500 --csv --log-file gpurun_out/${tag}_gemm_dram.csv $B > gpurun_out/${tag}_ncu_gemm_dram.log 2>&1
NCU="ncu --set full --import-source on --clock-control none"
# (template instantiations are told apart by their MANGLED names: ILi<BN>ELi<MODE>ELi<CTAS>ELi<EPI>E)
for spec in "attn_fwd:attn_tc_fwd_kernel:21" "attn_bwd_q:attn_tc_bwd_q:21" "attn_bwd_kv3:attn_tc_bwd_kv3:21" "gemm_wgrad_tn:gemm_tc_kernelILi256ELi2ELi2ELi128E:80" "gemm_o:gemm_tc_kernelILi256ELi0ELi2ELi32E:18" "gemm_dgrad_mlp:gemm_tc_kernelILi256ELi1ELi2ELi80E:20" "ln_bwd:ln_bwd_rows:40" "ctc:ctc_kernel:3"; do
  IFS=: read name kern skip <<< "$spec"
  timeout 600 $NCU --kernel-name-base mangled --kernel-name "regex:$kern" --launch-skip $skip --launch-count 1 -f -o gpurun_out/${tag}_prof_$name $B > gpurun_out/${tag}_ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/${tag}_prof_$name.ncu-rep --source 30 > gpurun_out/${tag}_ncu_${name}_summary.txt 2>&1
  case $name in attn_bwd_kv3) ;; *) rm -f gpurun_out/${tag}_prof_$name.ncu-rep ;; esac
done
ls -la gpurun_out/${tag}_* | head -40
