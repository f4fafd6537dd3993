"""Depthwise 3x3 kernel variants against each other, bit for bit, repeatedly (race detector):
python tools/dw_check.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ocr_inference_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)


def run(x, w, dil):
    y = torch.empty_like(x)
    N, H, W, Cn = x.shape
    rc = ctx.lib.bdrc_depthwise3x3_test(ctx.h, x.data_ptr(), N, H, W, Cn, dil, w.ctypes.data, y.data_ptr(),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
    ctx.check(rc, "dw test")
    return y


bad_total = 0
for (N, H, W, Cn) in [(48, 128, 128, 304), (48, 32, 32, 256), (5, 64, 64, 64), (148, 128, 128, 304)]:
    g = torch.Generator(device="cpu").manual_seed(H + Cn)
    x = torch.randn((N, H, W, Cn), generator=g).to(torch.bfloat16).cuda()
    w = np.ascontiguousarray(np.random.default_rng(Cn).standard_normal((9, Cn)).astype(np.float32))
    os.environ["BDRC_NO_DW_RING"] = "1"
    ref = run(x, w, 1)
    del os.environ["BDRC_NO_DW_RING"]
    first = run(x, w, 1)
    # the two kernels sum the 9 taps in different orders: equal up to fp32 rounding of bf16 outputs
    err = float((first.float() - ref.float()).abs().max())
    scale = float(ref.float().abs().max())
    bad = 0
    for i in range(20):
        y = run(x, w, 1)
        bad += int((y != first).sum().item()) > 0
    ok = err <= 0.02 * scale
    print(f"[{N},{H},{W},{Cn}] ring vs strip kernel: max |delta| {err:.4f} (scale {scale:.2f}) {'OK' if ok else 'MISMATCH'}; "
          f"{bad}/20 ring runs differ from the first", flush=True)
    bad_total += bad + (0 if ok else 1)
sys.exit(1 if bad_total else 0)
