"""Per-launch CUDA-event profile of the two networks (B200): python tools/layer_profile.py [seg|ocr] [N]"""
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ocr_inference_b200.engine import Engine  # noqa: E402
from ocr_inference_b200.synthetic_weights import easter2_weights, seg_weights  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "seg"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else (48 if which == "seg" else 64)
    eng = Engine(0)
    if which == "seg":
        model = eng.load_model(1, seg_weights(11, 1, 3), 1)
        x = (torch.rand((n, 512, 512), device="cuda") > 0.2).to(torch.uint8)
        run = lambda: eng.seg_forward(model, x, lowres=True)  # noqa: E731
    else:
        model = eng.load_model(2, easter2_weights(7, 77, 80), 77)
        x = torch.sign(torch.randn((n, 2000, 128), device="cuda")).to(torch.bfloat16)
        x[:, :, 80:] = 0
        run = lambda: eng.ocr_forward(model, x)  # noqa: E731
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    reps = 5
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    prof = eng.profile_results()
    agg = collections.OrderedDict()
    for name, ms, fl in prof:
        a = agg.setdefault(name, [0.0, 0.0, 0])
        a[0] += ms / reps
        a[1] += fl / reps
        a[2] += 1
    tot = sum(a[0] for a in agg.values())
    print(f"{which} n={n}: wall {e0.elapsed_time(e1) / reps:.3f} ms per forward, sum of launches {tot:.3f} ms")
    for name, (ms, fl, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        tf = fl / (ms / 1e3) / 1e12 if ms > 0 and fl > 0 else 0
        print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{cnt // reps:<3d} {tf:7.1f} TF/s  {name}")
    tcf = sum(a[1] for k, a in agg.items() if a[1] > 0 and not k.startswith(("stem", "aspp.dw", "dec.dw", "aspp.pool")))
    tcm = sum(a[0] for k, a in agg.items() if a[1] > 0 and not k.startswith(("stem", "aspp.dw", "dec.dw", "aspp.pool")))
    print(f"conv_tc aggregate: {tcf / (tcm / 1e3) / 1e12:.1f} TF/s over {tcm:.3f} ms")


if __name__ == "__main__":
    main()
