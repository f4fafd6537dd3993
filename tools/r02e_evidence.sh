# Round-2 closing evidence (one GPU).  Every command runs plain (exit 0) before it runs under ncu; numbers printed under
# ncu are never bench values.  .ncu-rep files are exported to CSV on the box (gpurun returns at most 64 MiB).
set -o pipefail
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 || exit 1
timeout 900 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -c 300 gpurun_out/r02e_bench.err
timeout 200 python tools/stages.py > gpurun_out/r02e_stages.txt 2>&1
timeout 100 python tools/layer_profile.py ocr 64 > gpurun_out/r02e_layer_profile_ocr64.txt 2>&1
timeout 100 python tools/layer_profile.py seg 148 > gpurun_out/r02e_layer_profile_seg148.txt 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02e_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02e_launches_bench_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02e_ncu_launches.log 2>&1
# the kernels this session changed: a-7 canvas + threshold, greedy CTC (bulk copies), squeeze-excite gate / apply
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'adaptive_threshold|ocr_canvas|ctc_greedy_bulk' -c 8 -f -o /tmp/r02e_a7a9 python tools/stages.py > gpurun_out/r02e_ncu_a7a9.log 2>&1
ncu -i /tmp/r02e_a7a9.ncu-rep --page raw --csv > gpurun_out/r02e_a7a9_full_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'se_partial_gate|se_apply' -s 6 -c 6 -f -o /tmp/r02e_se python tools/layer_profile.py ocr 64 > gpurun_out/r02e_ncu_se.log 2>&1
ncu -i /tmp/r02e_se.ncu-rep --page raw --csv > gpurun_out/r02e_se_full_raw.csv 2>/dev/null
head -8 gpurun_out/r02e_stages.txt; head -4 gpurun_out/r02e_layer_profile_ocr64.txt; head -2 gpurun_out/r02e_layer_profile_seg148.txt
wc -l gpurun_out/r02e_launches_bench_c2.csv gpurun_out/r02e_a7a9_full_raw.csv gpurun_out/r02e_se_full_raw.csv; du -sh gpurun_out
python - <<'P'
import json
d = json.load(open("gpurun_out/r02e_bench.json"))
print("C2", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "iob", round(d["e2e_io_binding"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), "whole", round(d["roofline"]["whole_step_frac"], 3), d["clocks"])
c1 = d["c1"]; print("c1", round(c1["lines_per_sec"]), c1["ms_per_step_loops"], round(c1["lines_per_sec_batch_of_512"]), round(c1["e2e_lines_per_sec"]), round(c1["roofline"]["frac"], 3))
c3 = d["c3"]; print("c3", round(c3["value"], 1), c3.get("cpu_baseline"))
print("c5", {k: v for k, v in d["c5"].items() if k in ("pages_per_sec", "host_cores_busy", "limiter")})
print("c4", d["other_workloads"]["c4"] if not isinstance(d["other_workloads"]["c4"], dict) else {k: v for k, v in d["other_workloads"]["c4"].items() if "per_sec" in k})
print("errors", d["other_workloads"].get("errors"))
P
