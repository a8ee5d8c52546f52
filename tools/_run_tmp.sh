timeout 600 python -m pytest tests -m gpu -x -q -k "attention or dropout" > gpurun_out/pytest_r26.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r26.log
tail -n 3 gpurun_out/pytest_r26.log
python tools/attn_timeline.py > gpurun_out/attn_timeline_r26.log 2>&1; cat gpurun_out/attn_timeline_r26.log
