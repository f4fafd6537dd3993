"""C3 (page -> text) with several host processes sharing one B200: python tools/c3_procs.py [n_procs threads]..."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402

if __name__ == "__main__":
    args = [int(a) for a in sys.argv[1:]] or [4, 2, 8, 1]
    for n, th in zip(args[0::2], args[1::2]):
        print(n, th, json.dumps(bench.bench_c3_procs(0, n, threads=th)), flush=True)
