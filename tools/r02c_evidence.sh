# Round-2 final evidence run (one GPU): tests, bench line, ncu launch list, ncu --set full of one segmentation forward.
# Every command runs plain (exit 0) before it runs under ncu; numbers printed under ncu are never bench values.
# The big .ncu-rep is exported to CSV and removed on the box (gpurun returns at most 64 MiB).
set -o pipefail
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 || exit 1
python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 300 gpurun_out/r02c_bench.err
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02c_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02c_launches_bench_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02c_ncu_launches.log 2>&1
python tools/layer_profile.py seg 148 > gpurun_out/r02c_layer_profile_seg148.txt 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc|dec_tail|stem_tc' -s 44 -c 44 -f -o /tmp/r02c_seg_convs python tools/layer_profile.py seg 148 > gpurun_out/r02c_ncu_seg_convs.log 2>&1
ncu -i /tmp/r02c_seg_convs.ncu-rep --page raw --csv > gpurun_out/r02c_seg_convs_full_raw.csv 2>/dev/null
python tools/layer_profile.py ocr 64 > gpurun_out/r02c_layer_profile_ocr64.txt 2>&1
python tools/stages.py > gpurun_out/r02c_stages.txt 2>&1
tail -3 gpurun_out/r02c_layer_profile_seg148.txt; wc -l gpurun_out/r02c_launches_bench_c2.csv gpurun_out/r02c_seg_convs_full_raw.csv; du -sh gpurun_out
