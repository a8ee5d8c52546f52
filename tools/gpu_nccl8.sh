#!/bin/bash
# 8-GPU bench: the trainer's default NCCL settings (NVLS off; NCCL_DEBUG=INFO of rank 0 kept to see the algorithm it picks) and Tree.
N=${1:-8}; tag=${2:-x}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29597 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager > gpurun_out/bench_${N}gpu_${tag}_$name.json 2> gpurun_out/bench_${N}gpu_${tag}_$name.err; }
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING,COLL NCCL_DEBUG_FILE=gpurun_out/nccl_${tag}_%h_%p.log
run tree NCCL_ALGO=Tree
ls gpurun_out/nccl_${tag}_* 2>/dev/null | head -1 | xargs -r -I{} sh -c 'grep -i -m 12 "algo\|Tree\|Ring\|NVLS" {} | cut -c1-200' > gpurun_out/nccl_${tag}_algo.txt
ls gpurun_out/nccl_${tag}_*.log 2>/dev/null | tail -n +2 | xargs -r rm -f
f=$(ls gpurun_out/nccl_${tag}_*.log 2>/dev/null | head -1); [ -n "$f" ] && head -c 300000 "$f" > gpurun_out/nccl_${tag}_rank.log && rm -f "$f"
python - <<PY
import json
for n in ("default","tree"):
    f="gpurun_out/bench_${N}gpu_${tag}_%s.json" % n
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "e2e ms", round(d["e2e"]["ms_per_step"],4), "h2d in loop", round(d["e2e"].get("h2d_ms_per_step_in_loop") or 0,3))
    except Exception as e:
        print(f, "unreadable", e)
PY
