python -m pytest tests -m gpu -x -q > gpurun_out/pytest13.log 2>&1; tail -3 gpurun_out/pytest13.log
python bench.py > gpurun_out/bench14.json 2> gpurun_out/bench14.err; tail -c 300 gpurun_out/bench14.err
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/plain_r1d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 546 -c 186 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_l_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:adaptive_threshold -s 4 -c 2 -o gpurun_out/r01d_a1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_a1_r1d.log 2>&1
ncu -i gpurun_out/r01d_a1.ncu-rep --page raw --csv > gpurun_out/r01d_a1_full_raw.csv 2>/dev/null
python tools/stage_bench.py --json gpurun_out/stage_roofline_d.json > gpurun_out/stage_bench_d.log 2>&1; tail -3 gpurun_out/stage_bench_d.log
