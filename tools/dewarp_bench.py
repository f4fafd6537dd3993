"""TPS dewarp of one 2048x3072x3 page on a B200, bilinear against cubic B-spline sampling: python tools/dewarp_bench.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ocr_inference_b200.inference import get_engine  # noqa: E402

eng = get_engine(0)
img = torch.randint(0, 256, (2048, 3072, 3), dtype=torch.uint8, device="cuda")
ctrl = np.array([[1000., 200.], [1000, 900], [1000, 1500], [1000, 2200], [1000, 2900], [0, 0], [2048, 0], [0, 3072],
                 [2048, 3072]])
par = np.zeros((12, 2))
par[10, 0] = par[11, 1] = 1.0
par[9] = [0.3, 0.4]


def timed(name, fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t) / reps * 1e3:.3f} ms")


timed("spline prefilter (u8 -> fp64 coefficients, both axes)", lambda: eng.spline3_prefilter(img))
timed("TPS + cubic sampling, prefilter included", lambda: eng.tps_remap_cubic(img, ctrl, par))
timed("TPS + bilinear sampling", lambda: eng.tps_remap(img, ctrl, par))
