"""GPU check of the implicit-GEMM conv engine against torch conv2d (fp32 math on bf16-rounded data).

    python tools/conv_check.py            # all cases, tcgen05 engine and SIMT engine
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ocr_inference_b200 import _lib  # noqa: E402


def run_case(ctx, name, N, H, W, Cin, Cout, R, S, stride, pad, dil, residual, relu, out_fp32, engine,
             seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, H, W, Cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, R, S, generator=g) / (Cin * R * S) ** 0.5).float()
    b = torch.randn(Cout, generator=g).float()
    sh, sw = stride
    ph, pw = pad
    dh, dw = dil
    Ho = (H + 2 * ph - dh * (R - 1) - 1) // sh + 1
    Wo = (W + 2 * pw - dw * (S - 1) - 1) // sw + 1
    res = torch.randn(N, Ho, Wo, Cout, generator=g).to(torch.bfloat16) if residual else None
    xd = x.cuda()
    resd = res.cuda() if residual else None
    y = torch.full((N, Ho, Wo, Cout), float("nan"), device="cuda",
                   dtype=torch.float32 if out_fp32 else torch.bfloat16)
    wc = np.ascontiguousarray(w.numpy())
    bc = np.ascontiguousarray(b.numpy())
    ctx.lib.bdrc_set_conv_engine(ctx.h, engine)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = ctx.lib.bdrc_conv2d_test(
        ctx.h, xd.data_ptr(), N, H, W, Cin, wc.ctypes.data, bc.ctypes.data, Cout, R, S, sh, sw, ph, pw,
        dh, dw, resd.data_ptr() if residual else None, int(relu), int(out_fp32), y.data_ptr(),
        torch.cuda.current_stream().cuda_stream)
    dt = time.perf_counter() - t0
    if rc != 0:
        print(f"[{name}] engine={engine} FAILED rc={rc}: {ctx.lib.bdrc_last_error(ctx.h).decode()}")
        return False
    # reference: fp32 conv on the bf16-rounded operands
    wr = w.to(torch.bfloat16).float().cuda()
    ref = torch.nn.functional.conv2d(xd.float().permute(0, 3, 1, 2), wr, b.cuda(), stride=stride,
                                     padding=pad, dilation=dil).permute(0, 2, 3, 1)
    if residual:
        ref = ref + resd.float()
    if relu:
        ref = torch.relu(ref)
    yf = y.float()
    err = (yf - ref).abs()
    tol = 2e-2 * ref.abs().max().item() + 1e-3
    nbad = int((~(err <= tol)).sum().item())
    ok = nbad == 0
    print(f"[{name}] engine={engine} N={N} HxW={H}x{W} Cin={Cin} Cout={Cout} k={R}x{S} s={stride} "
          f"d={dil} -> max_err={err.nan_to_num(1e9).max().item():.4g} tol={tol:.3g} bad={nbad} "
          f"nan={int(torch.isnan(yf).sum())} {'OK' if ok else 'MISMATCH'} ({dt*1e3:.1f} ms)")
    if not ok:
        bad = (~(err <= tol)).nonzero()[:5]
        for idx in bad:
            i = tuple(int(v) for v in idx)
            print("   ", i, float(yf[i]), float(ref[i]))
    return ok


CASES = [
    # name, N, H, W, Cin, Cout, R, S, stride, pad, dil, residual, relu, out_fp32
    ("gemm_1x1_k64", 1, 1, 128, 64, 64, 1, 1, (1, 1), (0, 0), (1, 1), False, False, True),
    ("gemm_1x1_k256", 1, 1, 256, 256, 128, 1, 1, (1, 1), (0, 0), (1, 1), False, False, True),
    ("gemm_1x1_n256", 2, 1, 384, 512, 256, 1, 1, (1, 1), (0, 0), (1, 1), False, True, False),
    ("gemm_1x1_n512", 1, 32, 32, 512, 512, 1, 1, (1, 1), (0, 0), (1, 1), True, True, False),
    ("c3x3_64_128sq", 2, 128, 128, 64, 64, 3, 3, (1, 1), (1, 1), (1, 1), True, True, False),
    ("c3x3_64_128sq_n5", 5, 128, 128, 64, 64, 3, 3, (1, 1), (1, 1), (1, 1), False, False, False),
    ("c3x3_64_128x8", 3, 8, 128, 64, 64, 3, 3, (1, 1), (1, 1), (1, 1), True, False, False),
    ("c3x3_128_64sq", 2, 64, 64, 128, 128, 3, 3, (1, 1), (1, 1), (1, 1), True, True, False),
    ("c3x3_256_32sq", 3, 32, 32, 256, 256, 3, 3, (1, 1), (1, 1), (1, 1), False, True, False),
    ("c3x3_d2_512", 2, 32, 32, 512, 512, 3, 3, (1, 1), (2, 2), (2, 2), True, True, False),
    ("c3x3_s2_64_128", 2, 128, 128, 64, 128, 3, 3, (2, 2), (1, 1), (1, 1), False, True, False),
    ("c3x3_s2_128_256", 2, 64, 64, 128, 256, 3, 3, (2, 2), (1, 1), (1, 1), False, True, False),
    ("c1x1_s2_64_128", 2, 128, 128, 64, 128, 1, 1, (2, 2), (0, 0), (1, 1), False, False, False),
    ("c1x1_304_256", 1, 128, 128, 304, 256, 1, 1, (1, 1), (0, 0), (1, 1), False, True, False),
    ("c1x1_64_48", 1, 128, 128, 64, 48, 1, 1, (1, 1), (0, 0), (1, 1), False, True, False),
    ("head_256_1", 2, 128, 128, 256, 1, 1, 1, (1, 1), (0, 0), (1, 1), False, False, True),
    ("c1d_k3_s2_128", 3, 1, 2000, 128, 128, 1, 3, (1, 2), (0, 1), (1, 1), False, True, False),
    ("c1d_k5_256", 3, 1, 500, 256, 256, 1, 5, (1, 1), (0, 2), (1, 1), False, True, False),
    ("c1d_k9_256", 2, 1, 500, 256, 256, 1, 9, (1, 1), (0, 4), (1, 1), False, False, False),
    ("c1d_k11_d2", 2, 1, 500, 256, 512, 1, 11, (1, 1), (0, 10), (1, 2), False, True, False),
    ("c1d_head_77", 2, 1, 500, 512, 77, 1, 1, (1, 1), (0, 0), (1, 1), False, False, True),
]


def main():
    os.environ["BDRC_ALLOW_PARTIAL_LIB"] = "1"
    ctx = _lib.Context(0)
    print(ctx.lib.bdrc_version().decode(), torch.cuda.get_device_name(0))
    only = sys.argv[1:] or None
    results = {}
    for engine in (1, 0):
        for case in CASES:
            if only and case[0] not in only:
                continue
            try:
                ok = run_case(ctx, *case, engine=engine)
            except Exception as e:  # noqa: BLE001
                print(f"[{case[0]}] engine={engine} EXCEPTION {e!r}")
                ok = False
                try:
                    torch.cuda.synchronize()
                except Exception as e2:  # noqa: BLE001
                    print("CUDA context is dead:", e2)
                    print("RESULT", results)
                    sys.exit(2)
            results[(case[0], engine)] = ok
    nfail = sum(1 for v in results.values() if not v)
    print(f"SUMMARY {len(results) - nfail}/{len(results)} passed")
    sys.exit(1 if nfail else 0)


if __name__ == "__main__":
    main()
