#!/bin/bash
# One multi-GPU bench line, launched the way the driver launches it.  Usage: tools/gpu_bench_n.sh <N> <tag> [extra bench args]
N=${1:-2}; tag=${2:-x}; shift; shift
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager "$@" > gpurun_out/bench_${N}gpu_$tag.json 2> gpurun_out/bench_${N}gpu_$tag.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${N}gpu_$tag.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "e2e ms", round(d["e2e"]["ms_per_step"],4), "h2d in loop", d["e2e"].get("h2d_ms_per_step_in_loop"), "nvls", d["e2e"].get("nccl_nvls"))
PY
