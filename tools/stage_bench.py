"""Per-stage roofline of every SURVEY 8(a) row on a B200: python tools/stage_bench.py [--json out.json]

Each stage runs through the C ABI on device-resident inputs at BASELINE.json sizes, timed with CUDA events
over `reps` calls that rotate through input sets larger than the 126 MB L2.  `achieved` = ALGORITHMIC bytes
(the bytes the stage must read and write, not what the kernel happens to move) or FLOPs per
call / time; `frac` = achieved / MEASURED_PEAKS.json (HBM copy bandwidth, or sustained bf16 TFLOP/s).
"""
import json
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from ocr_inference_b200 import synthetic  # noqa: E402
from ocr_inference_b200.engine import LAYOUT_TV, POST_SIGMOID, POST_SOFTMAX, Engine  # noqa: E402
from ocr_inference_b200.synthetic_weights import (easter2_flops_per_line, easter2_weights_fitted,  # noqa: E402
                                                  seg_flops_per_tile, seg_weights_fitted)

H, W = 2048, 3072


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1400.0, "fallback"


def timeit(fn, sets, reps=24, warm=4):
    for i in range(warm):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    eng = Engine(0)
    dev = eng.device
    hbm, tf, src = peaks()
    rows = []

    def add(stage, what, ms, bytes_=None, flops=None, unit_n=1, unit="page"):
        r = {"stage": stage, "what": what, "ms_per_call": ms, f"{unit}s_per_sec": unit_n / (ms / 1e3)}
        if bytes_ is not None:
            r.update(bound="hbm", achieved_gbs=bytes_ / (ms / 1e3) / 1e9, frac=bytes_ / (ms / 1e3) / 1e9 / hbm,
                     algorithmic_mb=bytes_ / 1e6)
        if flops is not None:
            r.update(bound="tensor", achieved_tflops=flops / (ms / 1e3) / 1e12, frac=flops / (ms / 1e3) / 1e12 / tf,
                     algorithmic_gflop=flops / 1e9)
        rows.append(r)
        print(f"{stage:5s} {what:58s} {ms:8.3f} ms  " +
              (f"{r['achieved_gbs']:7.0f} GB/s  {100 * r['frac']:5.1f} % of HBM" if bytes_ is not None else
               f"{r['achieved_tflops']:7.0f} TF/s  {100 * r['frac']:5.1f} % of bf16 sustained"), flush=True)

    pages_h = [synthetic.text_page(3003 + i)[0] for i in range(2)]
    pages = [torch.from_numpy(np.roll(pages_h[i % 2], 97 * i, axis=1).copy()).to(dev) for i in range(8)]   # 151 MB

    # a-1 ------------------------------------------------------------------------------------------
    tiles_buf = torch.empty((24, 512, 512), dtype=torch.uint8, device=dev)
    ms = timeit(lambda p: eng.seg_preprocess(p, out=tiles_buf), pages)
    add("a-1", "gray + adaptive threshold + tile, 2048x3072x3 page -> 24 u8 tiles", ms, bytes_=5 * H * W)
    eng.profile(True)
    for p in pages:
        eng.seg_preprocess(p, out=tiles_buf)
    torch.cuda.synchronize()
    ev = [t for (n, t, f) in eng.profile_results() if n.startswith("a-1")]
    eng.profile(False)
    add("a-1", "  same, kernel only (CUDA events around the launch, no host call overhead)", sum(ev) / len(ev),
        bytes_=5 * H * W)

    # a-2 ------------------------------------------------------------------------------------------
    seg = eng.load_model(1, seg_weights_fitted(11), 1)
    tsets = [(torch.rand((148, 512, 512), device=dev) > 0.2).to(torch.uint8) for _ in range(3)]
    ms = timeit(lambda t: eng.seg_forward(seg, t, lowres=True), tsets, reps=9, warm=3)
    add("a-2", "DeepLabV3+/R34 forward, 148 tiles (one chunk, 6.2 pages), bf16 tcgen05", ms,
        flops=148 * seg_flops_per_tile(1), unit_n=148 / 24)

    # a-3 ------------------------------------------------------------------------------------------
    meta = dict(H0=H, W0=W, Hc=H, Wc=W, nty=4, ntx=6, pad_x=0, pad_y=0)
    lows = [torch.randn((24, 128, 128, 1), device=dev) * 3 for _ in range(8)]
    mask = torch.empty((H, W), dtype=torch.uint8, device=dev)
    ms = timeit(lambda x: eng.seg_postprocess(x, meta, POST_SIGMOID, 0.9, lowres=True, out=mask), lows)
    add("a-3", "line mask from low-res logits (fused x4 upsample + threshold + stitch)", ms,
        bytes_=24 * 128 * 128 * 4 + H * W)
    fulls = [torch.randn((24, 1, 512, 512), device=dev) * 3 for _ in range(6)]
    ms = timeit(lambda x: eng.seg_postprocess(x, meta, POST_SIGMOID, 0.9, out=mask), fulls)
    add("a-3", "line mask from full-res fp32 logits [24,1,512,512] (ONNX-shaped output)", ms, bytes_=5 * H * W)
    full5 = [torch.randn((24, 5, 512, 512), device=dev) * 3 for _ in range(2)]
    mask5 = torch.empty((H, W, 5), dtype=torch.uint8, device=dev)
    ms = timeit(lambda x: eng.seg_postprocess(x, meta, POST_SOFTMAX, 0.8, out=mask5), full5, reps=12)
    add("a-3", "layout masks, softmax over 5 classes > 0.8", ms, bytes_=5 * 5 * H * W)

    # a-4 / a-6 ------------------------------------------------------------------------------------
    M = cv2.getRotationMatrix2D((W / 2, H / 2), 1.3, 1)
    ms = timeit(lambda p: eng.warp_affine(p, M), pages)
    add("a-4", "warpAffine rotation of the page (cv2 fixed point, bit exact)", ms, bytes_=6 * H * W)
    yy, xx = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32),
                            torch.arange(W, device=dev, dtype=torch.float32), indexing="ij")
    coords = torch.stack([yy + 3.3 * torch.sin(xx / 300), xx + 0.4])
    ms = timeit(lambda p: eng.remap_bilinear(p, coords), pages)
    add("a-6", "bilinear remap with a materialised coordinate field", ms, bytes_=(6 + 8) * H * W)
    from ocr_inference_b200.image_dewarping import ThinPlateSpline

    rng = np.random.default_rng(0)
    out_pts = np.array([[y, x] for y in np.linspace(100, H - 100, 6) for x in np.linspace(100, W - 100, 5)])
    in_pts = out_pts + rng.normal(0, 4.0, out_pts.shape)
    corners = np.array([[0, 0], [0, W - 1], [H - 1, 0], [H - 1, W - 1]], dtype=np.float64)
    tps = ThinPlateSpline(alpha=0.5)
    tps.fit(np.vstack([out_pts, corners]), np.vstack([in_pts, corners]))     # output grid -> input coordinates
    ctrl, params = tps.control_points, tps.parameters
    ms = timeit(lambda p: eng.tps_remap(p, ctrl, params), pages)
    add("a-6", "thin-plate-spline dewarp, coordinates evaluated in the kernel", ms, bytes_=6 * H * W)

    # a-5 ------------------------------------------------------------------------------------------
    page_h, boxes, _ = synthetic.text_page(3003)
    band = synthetic.band_mask(boxes, H, W, grow=2, thin=0.45)
    contours, _ = cv2.findContours(band, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    from ocr_inference_b200 import line_detection as ld

    rois, masks, hs = [], [], []
    for c in contours:
        roi, m, h = ld._roi_mask(c, H, W)
        rois.append(roi), masks.append(m), hs.append(h)
    pack = eng.pack_masks(masks, rois)
    ses = [ld._se_for(h, 2.0) for h in hs]
    idx = list(range(len(rois)))
    crops, hw, off = eng.extract_lines(pages[0], pack, idx, ses)
    hw_h = hw.cpu().numpy()
    crop_px = int((hw_h[:, 0].astype(np.int64) * hw_h[:, 1]).sum())
    ms = timeit(lambda p: eng.extract_lines(p, pack, idx, ses), pages)
    add("a-5", f"masked crops of {len(idx)} lines ({crop_px / 1e6:.1f} Mpx): dilate, AND, delete empty rows/cols, gather",
        ms, bytes_=6 * crop_px)

    # a-7 / a-8 / a-9 --------------------------------------------------------------------------------
    L = 64
    crop_sets = []
    for s in range(8):
        cs, _ = synthetic.glyph_lines(1001 + s, L)
        crop_sets.append(eng.pack_crops(cs) + (sum(c.shape[0] * c.shape[1] for c in cs),))
    torch.cuda.synchronize()
    src_px = crop_sets[0][3]
    ms = timeit(lambda c: eng.ocr_prepare_lines(c[0], c[1], c[2], L, 80, 2000, False), crop_sets)
    add("a-7", "64 ragged crops -> resize, pad, adaptive threshold -> bf16 [64,2000,128]", ms,
        bytes_=3 * src_px + L * 2 * 80 * 2000 * 128 // 80, unit_n=L, unit="line")
    ocr = eng.load_model(2, easter2_weights_fitted(), 77)
    lsets = [eng.ocr_prepare_lines(c[0], c[1], c[2], L, 80, 2000, False)[0] for c in crop_sets[:4]]
    ms = timeit(lambda x: eng.ocr_forward(ocr, x), lsets)
    add("a-8", "Easter2 forward, 64 lines x 2000 px, bf16 tcgen05", ms, flops=L * easter2_flops_per_line(77, 80, 2000),
        unit_n=L, unit="line")
    big = [torch.randn((1024, 500, 77), device=dev) for _ in range(2)]      # 158 MB each
    ms = timeit(lambda x: eng.ctc_greedy(x, layout=LAYOUT_TV), big, reps=12)
    add("a-9", "argmax + CTC collapse, 1024 lines [500,77] fp32", ms, bytes_=1024 * 500 * 77 * 4, unit_n=1024, unit="line")
    lg = [torch.randn((L, 500, 77), device=dev) for _ in range(8)]
    ms = timeit(lambda x: eng.ctc_greedy(x, layout=LAYOUT_TV), lg)
    add("a-9", "argmax + CTC collapse, 64 lines (C1 size: launch-latency bound)", ms, bytes_=L * 500 * 77 * 4, unit_n=L,
        unit="line")

    out = {"peaks": {"hbm_gbs": hbm, "bf16_tflops_sustained": tf, "source": src}, "stages": rows}
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
