#!/bin/bash
# GEMM experiment round trip: self-test (tcgen05 vs CUDA-core GEMM, timings) with and without ragged tiles, then bench.
tag=${1:-x}
mkdir -p gpurun_out
GEMM_TL="NT 1024x1024 bias" timeout 300 build/gemm_selftest 32 > gpurun_out/selftest_${tag}_ragged.log 2>&1; echo "exit $?" >> gpurun_out/selftest_${tag}_ragged.log
NDT1_GEMM_RAGGED=0 GEMM_TL="NT 1024x1024 bias" timeout 300 build/gemm_selftest 32 > gpurun_out/selftest_${tag}_regular.log 2>&1; echo "exit $?" >> gpurun_out/selftest_${tag}_regular.log
paste -d'\n' gpurun_out/selftest_${tag}_ragged.log gpurun_out/selftest_${tag}_regular.log | grep -E "OK|FAIL|exit|mean" | cut -c1-170
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"
NDT1_GEMM_RAGGED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager > gpurun_out/bench_${tag}_regular.json 2> gpurun_out/bench_${tag}_regular.err; echo "bench exit $?"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_${tag}.log 2>&1; tail -n 3 gpurun_out/pytest_${tag}.log
python - <<PY
import json
for f in ("gpurun_out/bench_${tag}.json","gpurun_out/bench_${tag}_regular.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "gemm frac", round(d["roofline"]["frac"],4), "gemm ms", round(d["roofline"]["gemm_ms_per_step"],4))
        for e in d["kernel_ms_per_step_alone"]["top"]:
            if "gemm" in e["kernel"] or "ctc" in e["kernel"]: print("   ", e["kernel"], round(e["ms_per_step"]*1000/e["launches_per_step"],1), "x", e["launches_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
