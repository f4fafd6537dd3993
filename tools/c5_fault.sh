# 2-GPU check of bench.py under torchrun: (1) the normal run, (2) rank 1 fails inside C5 -- rank 0 must still print
# the ONE JSON line (c5 = {"error": ...}) and every rank must leave with exit code 0 once the deadline passes.
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps $2 --warmup 3; }
run 29511 10 > gpurun_out/r02d_bench_2gpu.json 2> gpurun_out/r02d_bench_2gpu.err; echo "normal run rc=$?"
BDRC_BENCH_C5_FAULT=1 BDRC_C5_DEADLINE_S=30 run 29512 3 > gpurun_out/r02d_bench_2gpu_fault.json 2> gpurun_out/r02d_bench_2gpu_fault.err; echo "fault run rc=$?"
python - <<'P'
import json
for f in ("r02d_bench_2gpu", "r02d_bench_2gpu_fault"):
    txt = open(f"gpurun_out/{f}.json").read().strip().splitlines()
    print(f, "lines on stdout:", len(txt))
    d = json.loads(txt[-1])
    print("  C2", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "c5", {k: v for k, v in (d["c5"] or {}).items() if k in ("pages_per_sec", "host_cores_busy", "host_cores", "limiter", "error")})
P
tail -c 400 gpurun_out/r02d_bench_2gpu_fault.err
