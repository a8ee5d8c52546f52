#!/bin/bash
# Round-2 GPU round trip: parity suite, bench (all legs), reference arm.  Usage: tools/gpu_r2.sh <tag> [pytest -k expr]
tag=${1:-x}
mkdir -p gpurun_out
if [ -n "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -s -k "$2" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
else
  timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?" >> gpurun_out/bench_$tag.err
tail -n 15 gpurun_out/pytest_$tag.log; tail -n 3 gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json
