#!/bin/bash
# Final evidence run: bench (no profiler), ncu launch list of the same command, GEMM DRAM traffic, ncu --set full of the main kernels.
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc_kernel -c 400 --csv --log-file gpurun_out/${tag}_gemm_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_gemm_dram.log 2>&1
NCU="ncu --set full --import-source on --clock-control none"
for spec in "gemm_o:gemm_tc_kernel:215" "gemm_qkv:gemm_tc_kernel:214" "gemm_down:gemm_tc_kernel:213" "gemm_dgrad_down:gemm_tc_kernel:230" "attn_fwd:attn_tc_fwd_kernel:16" "attn_bwd_kv:attn_tc_bwd_kv2:16" "attn_bwd_q:attn_tc_bwd_q:16" "ln_bwd:ln_bwd_rows:40" "ln_fwd:ln_fwd_rows:40" "ctc:ctc_kernel:3" "adamw:adamw_fused:3"; do
  IFS=: read name kern skip <<< "$spec"
  timeout 600 $NCU --kernel-name regex:$kern --launch-skip $skip --launch-count 1 -f -o gpurun_out/${tag}_prof_$name python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/${tag}_prof_$name.ncu-rep --source 30 > gpurun_out/${tag}_ncu_${name}_summary.txt 2>&1
  # gpurun brings back at most 64 MiB: keep the raw report of a few kernels only, the text summary of all
  case $name in gemm_o|attn_fwd|ln_bwd) ;; *) rm -f gpurun_out/${tag}_prof_$name.ncu-rep ;; esac
done
ls -la gpurun_out/${tag}_*
