"""Build libbdrc_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m ocr_inference_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbdrc_b200.so")
OBJDIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--use_fast_math" if False else "-DBDRC_NO_FAST_MATH",  # exact expf/rounding needed for parity
    "-I", os.path.join(HERE, "..", "include"),
    *os.environ.get("BDRC_NVCC_EXTRA", "").split(),   # e.g. -DBDRC_EXPERIMENTS for the timing experiments of conv_tc.cu
]


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(paths: list[str]) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sources()
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "bdrc_b200.h"))
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_stamp = _stamp(hdrs)
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(OBJDIR, os.path.basename(s)[:-3] + ".o")
        st = os.path.join(OBJDIR, os.path.basename(s)[:-3] + ".stamp")
        want = _stamp([s]) + hdr_stamp
        have = open(st).read() if os.path.exists(st) else ""
        objs.append(o)
        if force or have != want or not os.path.exists(o):
            jobs.append((s, o, st, want))

    def compile_one(job):
        s, o, st, want = job
        cmd = [NVCC, *FLAGS, "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr)
        with open(st, "w") as f:
            f.write(want)
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(OUT):
        cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
