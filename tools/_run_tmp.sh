timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r21.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r21.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r21.log 2> gpurun_out/bench_r21.err
tail -n 12 gpurun_out/pytest_r21.log; cat gpurun_out/bench_r21.log; tail -n 3 gpurun_out/bench_r21.err
