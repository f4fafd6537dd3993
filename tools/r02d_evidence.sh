# Round-2 closing evidence (one GPU): the full GPU test suite, the squeeze-excite A/B (weights staged in shared memory
# against read straight from L2: same bits), the stage rooflines and the bench line of the final build.
set -o pipefail
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for k in 1 2; do
  BDRC_SE_STAGED=1 timeout 120 python tools/layer_profile.py ocr 64 > gpurun_out/r02d_ocr64_se_staged_$k.txt 2>&1
  timeout 120 python tools/layer_profile.py ocr 64 > gpurun_out/r02d_layer_profile_ocr64_$k.txt 2>&1
done
grep -h "wall\|ocr.se" gpurun_out/r02d_ocr64_se_staged_*.txt gpurun_out/r02d_layer_profile_ocr64_*.txt
timeout 300 python tools/stages.py > gpurun_out/r02d_stages.txt 2>&1; head -8 gpurun_out/r02d_stages.txt
timeout 900 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 300 gpurun_out/r02d_bench.err
python - <<'P'
import json
d = json.load(open("gpurun_out/r02d_bench.json"))
print("C2", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3))
for k in ("c1", "c3", "c5"):
    b = d.get(k) or {}
    print(k, {x: (round(v, 1) if isinstance(v, float) else v) for x, v in b.items() if x in ("lines_per_sec", "pages_per_sec", "ms_per_step", "e2e_lines_per_sec", "error")})
print("errors", d["other_workloads"].get("errors"))
P
