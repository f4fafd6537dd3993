python -m pytest tests -m gpu -x -q > gpurun_out/pytest14.log 2>&1; tail -3 gpurun_out/pytest14.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench16.json 2> gpurun_out/bench16.err; tail -c 300 gpurun_out/bench16.err
python tools/stage_bench.py --json gpurun_out/stage_roofline_e.json > gpurun_out/stage_bench_e.log 2>&1; grep -E "a-7|a-9" gpurun_out/stage_bench_e.log
