timeout 600 python -m pytest tests -m gpu -x -q -k "attention or full_size or ctc_small or benchmark_size" > gpurun_out/pytest_r11.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r11.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r11.log 2> gpurun_out/bench_r11.err
NCU="ncu --set full --import-source on --clock-control none"
timeout 900 $NCU --kernel-name regex:ctc_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_ctc_r11 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ctc_r11.log 2>&1
timeout 900 $NCU --kernel-name regex:attn_tc_fwd_kernel --launch-skip 16 --launch-count 1 -f -o gpurun_out/prof_attnfwd_r11 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attnfwd_r11.log 2>&1
timeout 900 $NCU --kernel-name regex:smooth_noise --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_smooth_r11 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_smooth_r11.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r11.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r11.log 2>&1
tail -n 3 gpurun_out/pytest_r11.log gpurun_out/bench_r11.log
