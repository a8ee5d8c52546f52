#!/bin/bash
# 8-GPU end-to-end diagnosis: default, NVLS off, half the bytes.  Usage: tools/gpu_e2e8.sh <N> <tag>
N=${1:-8}; tag=${2:-x}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager "$@" > gpurun_out/bench_${N}gpu_${tag}_$name.json 2> gpurun_out/bench_${N}gpu_${tag}_$name.err; }
run default
NCCL_NVLS_ENABLE=0 run nvls0
run half --e2e-copy-frac 0.5
python - <<PY
import json
for n in ("default","nvls0","half"):
    f="gpurun_out/bench_${N}gpu_${tag}_%s.json" % n
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "e2e ms", round(d["e2e"]["ms_per_step"],4), "h2d in loop", d["e2e"].get("h2d_ms_per_step_in_loop"), "alone", d["e2e"].get("h2d_ms_per_step_alone"))
    except Exception as e:
        print(f, "unreadable", e)
PY
