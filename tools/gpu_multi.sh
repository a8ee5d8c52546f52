#!/bin/bash
# Multi-GPU round trip (gpurun --gpus N): NCCL parity of the data-parallel trainer, then the N-GPU bench.  Usage: tools/gpu_multi.sh <N> <tag>
N=${1:-2}; tag=${2:-x}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/dp_parity.py --backend nccl > gpurun_out/dp_parity_${N}gpu_$tag.txt 2>&1; echo "dp_parity exit $?" >> gpurun_out/dp_parity_${N}gpu_$tag.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu_$tag.json 2> gpurun_out/bench_${N}gpu_$tag.err; echo "bench exit $?" >> gpurun_out/bench_${N}gpu_$tag.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus $N --steps 20 --warmup 5 --no-graph > gpurun_out/bench_${N}gpu_${tag}_nograph.json 2> gpurun_out/bench_${N}gpu_${tag}_nograph.err
tail -n 6 gpurun_out/dp_parity_${N}gpu_$tag.txt; tail -n 3 gpurun_out/bench_${N}gpu_$tag.err
python - <<PY
import json
for f in ("gpurun_out/bench_${N}gpu_$tag.json","gpurun_out/bench_${N}gpu_${tag}_nograph.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "e2e ms", round(d["e2e"]["ms_per_step"],4), "graph", d["config"].get("cuda_graph"))
    except Exception as e:
        print(f, "unreadable", e)
PY
