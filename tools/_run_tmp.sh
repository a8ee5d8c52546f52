NCU="ncu --set full --import-source on --clock-control none"
timeout 600 $NCU --kernel-name regex:gemm_tc_kernel --launch-skip 230 --launch-count 1 -f -o gpurun_out/prof_dgelu_r34 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dgelu_r34.log 2>&1
timeout 600 $NCU --kernel-name regex:gemm_tc_kernel --launch-skip 232 --launch-count 1 -f -o gpurun_out/prof_dgradup_r34 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dgradup_r34.log 2>&1
