#!/bin/bash
# ncu --set full captures (with source counters) of single kernels.  Usage: tools/gpu_profile.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none"
# the GELU / bf16-out GEMM of the self-test (second case: launches 25..48 of gemm_tc_kernel)
timeout 600 $NCU --kernel-name regex:gemm_tc_kernel --launch-skip 30 --launch-count 1 -f -o gpurun_out/prof_gemm_gelu_$tag ./build/gemm_selftest 32 > gpurun_out/ncu_gemm_gelu_$tag.log 2>&1
# kernels of one training step
timeout 900 $NCU --kernel-name regex:ctc_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_ctc_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ctc_$tag.log 2>&1
timeout 900 $NCU --kernel-name regex:attn_tc_fwd_kernel --launch-skip 16 --launch-count 1 -f -o gpurun_out/prof_attnfwd_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attnfwd_$tag.log 2>&1
ls -la gpurun_out/*.ncu-rep
