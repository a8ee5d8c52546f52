#!/bin/bash
# ncu launch list + tensor-pipe counters of one eager step (the two cheap passes of tools/gpu_final_r2.sh).  Usage: tools/gpu_launchlist.sh [tag]
tag=${1:-r02d}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,gpu__time_duration.sum --clock-control none -k regex:"gemm_tc_kernel|attn_tc_" -c 500 --csv --log-file gpurun_out/${tag}_tensor_pipe.csv $B > gpurun_out/${tag}_ncu_tensor_pipe.log 2>&1
ls -la gpurun_out/${tag}_*
