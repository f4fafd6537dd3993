python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/layer_profile.py seg 148 2>&1 | head -12
python tools/e2e_timeline.py > gpurun_out/e2e_timeline2.txt 2>&1; grep -v "host copy\|^cores" gpurun_out/e2e_timeline2.txt
