"""C1 alone, three times, from the package in the current directory (A/B of two builds on one box):
cd <tree> && python /root/repo/tools/c1_only.py"""
import os
import sys

sys.path.insert(0, os.getcwd())
import bench  # noqa: E402
from ocr_inference_b200.engine import Engine  # noqa: E402

if __name__ == "__main__":
    eng = Engine(0)
    for rep in range(3):
        r = bench.bench_c1(eng, 20, 3, False)
        k = r["ms_by_kernel"]
        print(f"{os.getcwd()} rep {rep}: {r['ms_per_step']:.4f} ms/step {r['lines_per_sec']:.0f} lines/s  kernel sum {sum(k.values()):.4f}  "
              f"a-7 {sum(v for n, v in k.items() if n.startswith('a-7')):.4f} se {k.get('ocr.se', 0):.4f}  512: {r['lines_per_sec_batch_of_512']:.0f}  fixed 80x2000: {r.get('lines_per_sec_fixed_size_80x2000', 0):.0f}", flush=True)
