timeout 600 python -m pytest tests -m gpu -x -q -k "attention or dropout or full_size" > gpurun_out/pytest_r31.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r31.log
tail -n 3 gpurun_out/pytest_r31.log
python tools/attn_timeline.py 2>&1 | head -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r31.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r31.log 2>&1
