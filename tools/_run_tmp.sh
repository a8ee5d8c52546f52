./build/gemm_selftest 32 > gpurun_out/selftest_r16.log 2>&1
NDT1_GEMM_BN=128 ./build/gemm_selftest 32 > gpurun_out/selftest_r16_bn128.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r16.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r16.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r16.log 2> gpurun_out/bench_r16.err
NDT1_GEMM_BN=128 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r16_bn128.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r16.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r16.log 2>&1
tail -n 3 gpurun_out/pytest_r16.log
