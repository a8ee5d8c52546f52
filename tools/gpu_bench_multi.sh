N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r01_bench_${N}gpu.json 2> gpurun_out/r01_bench_${N}gpu.err
tail -c 1500 gpurun_out/r01_bench_${N}gpu.json; tail -n 3 gpurun_out/r01_bench_${N}gpu.err
