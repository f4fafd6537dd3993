"""Host-side packing of ragged line crops into the pinned upload buffer: python tools/pack_bench.py
numpy slice assignment (one core) against bdrc_host_gather with 1..8 native threads, and the whole run_batch call."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ocr_inference_b200 import synthetic  # noqa: E402
from ocr_inference_b200.inference import get_engine  # noqa: E402

eng = get_engine(0)
crops = synthetic.glyph_lines(1001, 64)[0]
print("lines", len(crops), "MB", sum(c.nbytes for c in crops) / 1e6, "cores", len(os.sched_getaffinity(0)))
for th in (0, 1, 2, 4, 8):
    eng._gather_threads = th            # 0: numpy slice assignment on the calling thread
    crops_t = crops
    for _ in range(3):
        eng.pack_crops(crops_t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        eng.pack_crops(crops_t)
    torch.cuda.synchronize()
    print("numpy path" if th == 0 else f"gather x{th}", f"{(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per pack_crops (incl. H2D)")

# the whole call: pack -> upload -> a-7 -> Easter2 -> greedy CTC -> text
from tools.pipeline_profile import build_pipeline  # noqa: E402

ocr = build_pipeline(eng).ocr_inference
for th in (0, 2, 4, 8):
    eng._gather_threads = th
    for _ in range(3):
        ocr.run_batch(crops, pre_pad=False)
    t0 = time.perf_counter()
    for _ in range(20):
        texts = ocr.run_batch(crops, pre_pad=False)
    dt = (time.perf_counter() - t0) / 20
    print(f"run_batch, gather threads {th}: {dt * 1e3:.3f} ms per 64 lines = {64 / dt:.0f} lines/s")
