#!/bin/bash
# Round-2 GPU round trip: parity suite, bench (all legs).  Usage: tools/gpu_r2.sh <tag> [pytest -k expr] [extra bench args]
tag=${1:-x}
mkdir -p gpurun_out
if [ -n "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -s -k "$2" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
else
  timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 $3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?" >> gpurun_out/bench_$tag.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline > gpurun_out/bench_${tag}_nograph.json 2> gpurun_out/bench_${tag}_nograph.err
tail -n 12 gpurun_out/pytest_$tag.log; tail -n 3 gpurun_out/bench_$tag.err; python - <<PY
import json
for f in ("gpurun_out/bench_$tag.json","gpurun_out/bench_${tag}_nograph.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "gemm frac", round(d["roofline"]["frac"],4), "graph", d["config"].get("cuda_graph"))
    except Exception as e:
        print(f, "unreadable", e)
PY
