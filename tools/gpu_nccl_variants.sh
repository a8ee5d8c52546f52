#!/bin/bash
# NCCL algorithm variants of the N-GPU bench (end-to-end vs resident).  Usage: tools/gpu_nccl_variants.sh <N> <tag>
N=${1:-4}; tag=${2:-x}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29595 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager > gpurun_out/bench_${N}gpu_${tag}_$name.json 2> gpurun_out/bench_${N}gpu_${tag}_$name.err; }
run ring NCCL_ALGO=Ring
run tree NCCL_ALGO=Tree
run nvls1 NCCL_NVLS_ENABLE=1
run ring_simple NCCL_ALGO=Ring NCCL_PROTO=Simple
python - <<PY
import json
for n in ("ring","tree","nvls1","ring_simple"):
    f="gpurun_out/bench_${N}gpu_${tag}_%s.json" % n
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "e2e ms", round(d["e2e"]["ms_per_step"],4), "h2d in loop", round(d["e2e"].get("h2d_ms_per_step_in_loop") or 0,3))
    except Exception as e:
        print(f, "unreadable", e)
PY
