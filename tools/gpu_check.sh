#!/bin/bash
# One GPU round trip: GEMM self-test, GPU parity suite, bench, ncu launch list.  Usage: tools/gpu_check.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
(timeout 300 ./build/gemm_selftest 32 > gpurun_out/selftest_$tag.log 2>&1; echo "exit $?" >> gpurun_out/selftest_$tag.log)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
tail -n 4 gpurun_out/selftest_$tag.log gpurun_out/pytest_$tag.log gpurun_out/bench_$tag.log
