"""Run-to-run bit reproducibility of the two forward passes (a data race shows up as differing bits):
python tools/determinism_check.py [n_tiles] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ocr_inference_b200.engine import Engine  # noqa: E402
from ocr_inference_b200.synthetic_weights import easter2_weights, seg_weights  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
eng = Engine(0)
if os.environ.get("BDRC_CONV_ENGINE"):
    eng.set_conv_engine(int(os.environ["BDRC_CONV_ENGINE"]))
seg = eng.load_model(1, seg_weights(11, 1, 3), 1)
x = (torch.rand((n, 512, 512), device="cuda") > 0.2).to(torch.uint8)
ref = eng.seg_forward(seg, x, lowres=True).clone()
bad = 0
for i in range(reps):
    y = eng.seg_forward(seg, x, lowres=True)
    d = int((y != ref).sum().item())
    if d:
        bad += 1
        print(f"seg rep {i}: {d} differing values, max |delta| {float((y - ref).abs().max()):.4f}")
print(f"seg n={n}: {bad}/{reps} runs differ from the first")
ocr = eng.load_model(2, easter2_weights(7, 77, 80), 77)
xl = torch.sign(torch.randn((6, 2000, 128), device="cuda")).to(torch.bfloat16)
xl[:, :, 80:] = 0
ref = eng.ocr_forward(ocr, xl).clone()
bad = 0
for i in range(reps):
    y = eng.ocr_forward(ocr, xl)
    d = int((y != ref).sum().item())
    if d:
        bad += 1
        print(f"ocr rep {i}: {d} differing values")
print(f"ocr L=6: {bad}/{reps} runs differ from the first")
