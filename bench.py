#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 page->text hot path.

Workload (BASELINE.json configs[1], "C2"): a batch of 16 synthetic 2048x3072 pages per GPU per
step -> 384 tiles of 512x512 -> binarise/tile (a-1) -> DeepLabV3+/R34 forward (a-2) -> mask
threshold (a-3).  Metric: pages/s (whole job, all ranks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`value`  : inputs (u8 pages) resident in HBM, masks left in HBM, timed with CUDA events.
`e2e`    : LineDetection.predict_batch(host numpy pages) -> host numpy masks (pinned H2D / D2H inside
           the timed region).
`roofline`: tcgen05 implicit-GEMM conv kernel, algorithmic FLOPs / CUDA-event time per launch,
           recorded live during the timed steps, vs MEASURED_PEAKS.json.
`cpu_baseline`: the oracle (reference host code restated + torch-CPU fp32 network) on the host
           cores, rank 0, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAGES_PER_STEP = 16
PAGE_H, PAGE_W = 2048, 3072
SEG_SEED, DATA_SEED = 11, 2002
METRIC, UNIT = "pages_per_sec", "pages/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1371.0), d.get("hbm_gbs", 6544.0), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


_SAMPLER_SRC = r"""
import sys, time
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
    print("max", float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)), flush=True)
    while True:
        mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        try:
            pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
        except Exception:
            pw = -1.0
        print("s", time.time(), mhz, mask, pw, flush=True)
        time.sleep(0.025)
except Exception as e:
    print("err", repr(e), flush=True)
"""


class ClockSampler:
    """SM clock, power and throttle reasons during the timed regions, read through NVML by a SEPARATE
    process started before the CUDA context exists: NVML calls (and nvidia-smi forks) made from the
    process that launches kernels contend with the launch path and stalled it for tens of
    milliseconds, perturbing the very region being timed.  Samples are time-stamped; summary() keeps
    those inside the marked [t0, t1] windows."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        import subprocess

        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES, NVML indexes physical GPUs
        phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
        self.windows: list[tuple[float, float]] = []
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(phys)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception as e:  # noqa: BLE001
            self.proc, self.err = None, repr(e)

    def mark(self, t0: float, t1: float):
        self.windows.append((t0, t1))

    def summary(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"sampler unavailable: {self.err}"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
            out, _ = self.proc.communicate()
        max_mhz, rows, err = None, [], None
        for ln in out.splitlines():
            f = ln.split()
            if f and f[0] == "max":
                max_mhz = float(f[1])
            elif f and f[0] == "s" and len(f) >= 5:
                rows.append((float(f[1]), float(f[2]), int(f[3]), float(f[4])))
            elif f and f[0] == "err":
                err = ln
        inside = [r for r in rows if any(a <= r[0] <= b for a, b in self.windows)]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": max_mhz, "reasons": [f"no samples inside the timed region ({err})"],
                    "samples": 0}
        sm = sorted(r[1] for r in inside)
        seen = 0
        for r in inside:
            seen |= r[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": max_mhz,
                "power_w_max": max(r[3] for r in inside),
                "reasons": [n for bit, n in self.REASONS.items() if seen & bit], "samples": len(inside)}


def make_pages(n: int, seed: int):
    from ocr_inference_b200 import synthetic

    base = [synthetic.page(seed + i, PAGE_H, PAGE_W)[0] for i in range(min(n, 4))]
    # 16 distinct pages from 4 generated ones (shift / contrast), keeps set-up time short
    pages = []
    for i in range(n):
        p = base[i % len(base)]
        if i >= len(base):
            p = np.roll(p, 37 * i, axis=1)
            p = np.clip(p.astype(np.int16) + (i - 8), 0, 255).astype(np.uint8)
        pages.append(np.ascontiguousarray(p))
    return pages


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle) -- the only place bench.py executes oracle/
# ------------------------------------------------------------------------------------------------
def cpu_reference_pages_per_sec(n_pages: int, steps: int = 1, warmup: int = 0):
    import cv2
    import torch

    from oracle import hostops as ho
    from oracle import nets

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cv2.setNumThreads(-1)
    w = nets.seg_weights(SEG_SEED, 1, 3)
    pages = make_pages(n_pages, DATA_SEED)

    def one_pass():
        for p in pages:
            tiles, meta = ho.seg_preprocess(p, closed_form=False)
            x = np.repeat(tiles[:, None].astype(np.float32), 3, axis=1)
            logits = np.concatenate([nets.seg_forward(w, x[i:i + 8]) for i in range(0, len(x), 8)])
            ho.line_postprocess(logits, meta, 0.9, closed_form=False)

    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = time.perf_counter() - t0
    return n_pages * steps / dt, dt / steps, cores


def cpu_reference_page_to_text(page, gpu_result):
    """The oracle's whole page -> text pipeline (tile/binarise, torch-CPU fp32 DeepLabV3+ with the fitted head,
    mask, the reference's cv2/numpy line extraction, line prep, torch-CPU Easter2, greedy CTC) on ONE page, on all
    host cores, beside the GPU pipeline's result for the same page."""
    import cv2
    import torch

    from ocr_inference_b200 import synthetic
    from ocr_inference_b200.data import OpStatus
    from ocr_inference_b200.synthetic_weights import easter2_weights_fitted, seg_weights_fitted
    from oracle import geometry as og
    from oracle import hostops as ho
    from oracle import nets

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cv2.setNumThreads(-1)
    ws, we = seg_weights_fitted(11), easter2_weights_fitted()
    t0 = time.perf_counter()
    tiles, meta = ho.seg_preprocess(page, closed_form=False)
    x = np.repeat(tiles[:, None].astype(np.float32), 3, axis=1)
    logits = np.concatenate([nets.seg_forward(ws, x[i:i + 8]) for i in range(0, len(x), 8)])
    mask = ho.line_postprocess(logits, meta, 0.9, closed_form=False)
    r = og.page_to_lines(page, mask, 2.0, 2.5)
    texts = [] if isinstance(r, str) else og.recognise(r[2], we, synthetic.CHARSET_76)[0]
    dt = time.perf_counter() - t0
    out = {"value": 1.0 / dt, "unit": "pages/s", "lines_per_sec": len(texts) / dt, "cores": cores, "kind": "port",
           "sample": "1 of the pages: oracle/ page -> text (cv2/numpy host code of the reference restated, torch-CPU fp32 "
                     "networks), all host threads", "seconds_per_page": dt, "lines": len(texts)}
    if gpu_result[0] == OpStatus.SUCCESS:
        got = [l.text for l in gpu_result[1][2]]
        want = [t.strip().replace("\u00a7", " ") for t in texts]
        out["gpu_lines"] = len(got)
        out["gpu_lines_identical_to_cpu_oracle"] = sum(a == b for a, b in zip(got, want)) if len(got) == len(want) else None
        # whole-page identity is decided by the few mask pixels whose logit sits inside the bf16 tolerance of the
        # threshold (a different contour polygon -> a crop shifted by a pixel -> another resize phase); report how far
        # apart the two pipelines actually are: line boxes and character-level agreement
        if not isinstance(r, str):
            gb = [(l.bbox.x, l.bbox.y, l.bbox.w, l.bbox.h) for l in gpu_result[1][1]]
            ob = [tuple(l.bbox) for l in r[1]]
            out["gpu_line_boxes_identical"] = sum(a == b for a, b in zip(gb, ob)) if len(gb) == len(ob) else None
            out["gpu_line_boxes_max_abs_diff_px"] = (max(max(abs(p - q) for p, q in zip(a, b)) for a, b in zip(gb, ob))
                                                     if len(gb) == len(ob) and gb else None)
        if len(got) == len(want) and got:
            dist = sum(_edit_distance(a, b) for a, b in zip(got, want))
            out["gpu_vs_cpu_character_agreement"] = 1.0 - dist / max(1, sum(max(len(a), len(b)) for a, b in zip(got, want)))
    return out


def _edit_distance(a: str, b: str) -> int:
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_pages = 1
    v, sec_per_step, cores = cpu_reference_pages_per_sec(n_pages, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "C2: 2048x3072 pages -> 512x512 tiles -> DeepLabV3+/R34 forward -> mask threshold",
                   "pages_per_step": n_pages, "tiles_per_page": 24},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_pages} page(s) (24 tiles) per step through oracle/ (reference host code "
                                   "restated with cv2/numpy + torch-CPU fp32 network), all host threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# secondary workloads: C1 (line recognition) and C3 (page -> text), reported next to the C2 headline
# ------------------------------------------------------------------------------------------------
C1_LINES = 64


def _ocr_config():
    from ocr_inference_b200 import synthetic
    from ocr_inference_b200.data import CharsetEncoder, OCRModelConfig
    from ocr_inference_b200.synthetic_weights import easter2_weights_fitted
    from ocr_inference_b200.weights import write_blob

    return OCRModelConfig(write_blob(easter2_weights_fitted()), "Easter2", 2000, 80, "input", "output", True, False,
                          CharsetEncoder.WYLIE, synthetic.CHARSET_76, True, "fitted")


def cpu_reference_lines_per_sec(n_lines: int):
    """Oracle arm of C1: reference line prep (cv2) + torch-CPU fp32 Easter2 + greedy CTC."""
    import cv2
    import torch

    from ocr_inference_b200 import synthetic
    from ocr_inference_b200.synthetic_weights import easter2_weights_fitted
    from oracle import hostops as ho
    from oracle import nets

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cv2.setNumThreads(-1)
    w = easter2_weights_fitted()
    crops, _ = synthetic.glyph_lines(1001, n_lines)
    t0 = time.perf_counter()
    x = np.stack([ho.prepare_ocr_line(c, 80, 2000, pre_pad=False).reshape(80, 2000) for c in crops])
    logits = nets.easter2_forward(w, x)
    ids = [ho.ctc_greedy_ids(logits[i], 77) for i in range(n_lines)]
    dt = time.perf_counter() - t0
    return n_lines / dt, cores, ids


def bench_c1(eng, steps: int, warmup: int, with_cpu: bool):
    """64 ragged glyph-line crops -> a-7 prep -> Easter2 (tcgen05) -> greedy CTC -> text."""
    import torch

    from ocr_inference_b200 import synthetic
    from ocr_inference_b200.inference import OCRInference
    from ocr_inference_b200.synthetic_weights import easter2_flops_per_line

    ocr = OCRInference(_ocr_config(), None, engine=eng)
    crops, truth = synthetic.glyph_lines(1001, C1_LINES)
    dev, hw, off = eng.pack_crops(crops)
    blank = ocr.ctc_decoder.blank_index

    def step_device():
        lines, _ = eng.ocr_prepare_lines(dev, hw, off, C1_LINES, 80, 2000, False)
        return eng.ctc_greedy(eng.ocr_forward(ocr.ocr_session, lines), blank=blank)

    for _ in range(warmup):
        step_device()
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (eng.launch_count() - l0) // steps
    eng.profile(True)
    for _ in range(steps):
        step_device()
    torch.cuda.synchronize()
    prof = eng.profile_results()
    eng.profile(False)
    tc = [(n, t, f) for (n, t, f) in prof if f > 0]
    tc_ms, tc_fl = sum(x[1] for x in tc), sum(x[2] for x in tc)
    by = {}
    for n, t, f in prof:
        by[n] = by.get(n, 0.0) + t / steps
    # the same path on a page-batch-sized set of lines (8 x the C1 set): what a multi-page caller sees
    big = [eng.pack_crops(crops) for _ in range(8)]
    import torch as _t
    pad16 = lambda t: _t.nn.functional.pad(t, (0, (-int(t.numel())) % 16))   # crop offsets must stay 16-byte aligned
    dev8 = _t.cat([pad16(b[0]) for b in big])
    sizes = [int(pad16(b[0]).numel()) for b in big]
    hw8 = _t.cat([b[1] for b in big])
    off8 = _t.cat([b[2] + sum(sizes[:i]) for i, b in enumerate(big)])
    L8 = 8 * C1_LINES

    def step_big():
        lines, _ = eng.ocr_prepare_lines(dev8, hw8, off8, L8, 80, 2000, False)
        return eng.ctc_greedy(eng.ocr_forward(ocr.ocr_session, lines), blank=blank)

    for _ in range(2):
        step_big()
    torch.cuda.synchronize()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for _ in range(5):
        ids8, lens8 = step_big()
    b1.record()
    torch.cuda.synchronize()
    ms_big = b0.elapsed_time(b1) / 5
    eng.profile(True)
    for _ in range(3):
        step_big()
    torch.cuda.synchronize()
    prof_big = [(n_, t_, f_) for (n_, t_, f_) in eng.profile_results() if f_ > 0]
    eng.profile(False)
    big_tf = sum(x[2] for x in prof_big) / (sum(x[1] for x in prof_big) / 1e3) / 1e12 if prof_big else 0.0
    ids1, lens1 = step_device()
    valid = _t.arange(ids1.shape[1], device=ids1.device)[None, :] < lens1[:, None]     # ids beyond lens are unwritten
    same_big = bool(_t.equal(lens8[:C1_LINES], lens1) and _t.equal(ids8[:C1_LINES][valid], ids1[valid]))
    for _ in range(2):
        texts = ocr.run_batch(crops, pre_pad=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        texts = ocr.run_batch(crops, pre_pad=False)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    want = ["".join(synthetic.CHARSET_76[i] for i in t) for t in truth]
    peak_tf, _, peak_src = peaks()
    out = {
        "workload": f"C1: {C1_LINES} ragged glyph-line crops -> resize/pad/binarise -> Easter2 (CTC-trained on synthetic "
                    "glyphs) -> greedy CTC -> text",
        "lines_per_sec": C1_LINES / (ms / 1e3), "ms_per_step": ms, "launches_per_step": int(launches),
        "lines_per_sec_batch_of_512": L8 / (ms_big / 1e3), "batch_of_512_equals_batch_of_64": same_big,
        "e2e_lines_per_sec": C1_LINES / (e2e_ms / 1e3), "e2e_ms_per_step": e2e_ms,
        "e2e_api": "OCRInference.run_batch(host crops) -> list[str]",
        "h2d_bytes_per_step": int(sum(c.nbytes for c in crops)),
        "algorithmic_gflop_per_line": easter2_flops_per_line(77, 80, 2000) / 1e9,
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (Easter2 conv1d layers)",
                     "achieved": tc_fl / (tc_ms / 1e3) / 1e12 if tc_ms else 0.0, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (tc_fl / (tc_ms / 1e3) / 1e12) / peak_tf if tc_ms else 0.0, "peak_source": peak_src,
                     "share_of_step": tc_ms / steps / ms if ms else None},
        "roofline_batch_of_512": {"achieved": big_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": big_tf / peak_tf},
        "lines_equal_to_truth": sum(a == b for a, b in zip(texts, want)) / C1_LINES,
        "ms_by_kernel": {k: round(v, 4) for k, v in sorted(by.items(), key=lambda kv: -kv[1])},
    }
    if with_cpu:
        v, cores, ref_ids = cpu_reference_lines_per_sec(16)
        got_ids = ocr.run_batch(crops[:16], pre_pad=False, return_ids=True)
        out["cpu_baseline"] = {"value": v, "unit": "lines/s", "cores": cores, "kind": "port",
                               "sample": "16 of the 64 lines: cv2 prep + torch-CPU fp32 Easter2 + greedy CTC"}
        out["lines_identical_to_cpu_oracle"] = sum(a == b for a, b in zip(got_ids, ref_ids)) / 16
    return out


def bench_c3(eng, steps: int, pages_per_step: int = 96, workers: int | None = None, with_cpu: bool = False):
    """Full page -> text: OCRPipeline.run_ocr_batch on synthetic text pages (host numpy in, OCRLine text out)."""
    from ocr_inference_b200 import synthetic
    from ocr_inference_b200.data import LineDetectionConfig, OpStatus
    from ocr_inference_b200.inference import OCRPipeline
    from ocr_inference_b200.synthetic_weights import seg_weights_fitted
    from ocr_inference_b200.weights import write_blob

    pipe = OCRPipeline(_ocr_config(), LineDetectionConfig("", write_blob(seg_weights_fitted(11)), "deeplabv3plus-r34",
                                                          512, ["line"]), None, engine=eng)
    base = [synthetic.text_page(3003 + i) for i in range(4)]
    pages = [base[i % 4][0] for i in range(pages_per_step)]
    n_truth = sum(len(base[i % 4][2]) for i in range(pages_per_step))
    kw = dict(k_factor=2.0, bbox_tolerance=2.5)        # the reference's Demo-OCR.ipynb parameters
    pipe.run_ocr_batch(pages, workers=workers, **kw)      # warm-up: pinned mask buffers, allocator pools
    t0 = time.perf_counter()
    for _ in range(steps):
        res = pipe.run_ocr_batch(pages, workers=workers, **kw)
    dt = (time.perf_counter() - t0) / steps
    ok = [r for r in res if r[0] == OpStatus.SUCCESS]
    n_lines = sum(len(r[1][2]) for r in ok)
    want = [{"".join(synthetic.CHARSET_76[i] for i in t).strip().replace("§", " ") for t in base[i % 4][2]}
            for i in range(pages_per_step)]
    n_right = sum(sum(l.text in want[i] for l in r[1][2]) for i, r in enumerate(res) if r[0] == OpStatus.SUCCESS)
    cpu = None
    if with_cpu:
        cpu = cpu_reference_page_to_text(base[0][0], res[0])
    n_geo = len(getattr(pipe, "_gpool", None) or [])
    pipe.close()
    return {
        **({"cpu_baseline": cpu} if cpu else {}),
        "workload": f"C3: {pages_per_step} synthetic 2048x3072 text pages per step through OCRPipeline.run_ocr_batch "
                    "(staged: pages grouped 6 at a time through tile/binarise + DeepLabV3+ with fitted head + mask, cv2 "
                    "contours + reading order in geometry worker processes (no CUDA in them), GPU line crops, all lines of "
                    "the batch through line prep + Easter2 + greedy CTC in one call; ONE CUDA context)",
        "e2e_pages_per_sec": pages_per_step / dt, "e2e_lines_per_sec": n_lines / dt, "ms_per_page": dt * 1e3 / pages_per_step,
        "host_threads": workers, "geometry_processes": n_geo, "pages_ok": len(ok), "lines_found": n_lines, "lines_on_pages": n_truth,
        "lines_equal_to_truth": n_right,
        "note": "host-bound: cv2.findContours + reading-order Python (SURVEY 8(a)-10) take ~15-20 ms/page/core on the noisy "
                "masks of the fitted head (~750 contours per page); GPU stages ~3 ms/page",
    }


def bench_c4(eng, steps: int = 2):
    """Modern-layout path (torch-path semantics: float-bilinear resize_clamp, pad 255, RGB tiles, /255 in the
    stem, softmax over 5 classes, stitch, crop, resize back, > 0.8) on mixed page sizes."""
    import torch

    from ocr_inference_b200 import synthetic
    from ocr_inference_b200.data import LayoutDetectionConfig
    from ocr_inference_b200.inference import ModernBookFormatLayoutDetection
    from ocr_inference_b200.synthetic_weights import seg_weights
    from ocr_inference_b200.weights import write_blob

    classes = ["background", "image", "line", "caption", "margin"]
    det = ModernBookFormatLayoutDetection(LayoutDetectionConfig(write_blob(seg_weights(13, 5, 3)), "", "deeplabv3plus", 512,
                                                                classes), engine=eng)
    sizes = [(472, 2550), (1200, 1800), (2048, 3072), (3300, 2550), (700, 4600), (400, 1500)]
    pages = [synthetic.page(4004 + i, h, w, rgb=True)[0] for i, (h, w) in enumerate(sizes)]
    for p in pages:
        det.predict_masks(p, 0.8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for p in pages:
            m = det.predict_masks(p, 0.8)
    dt1 = (time.perf_counter() - t0) / steps
    batch = pages * 2                                   # 12 pages, 92 tiles: tiles of several pages per forward call
    det.predict_masks_batch(batch, 0.8, copy=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        ms = det.predict_masks_batch(batch, 0.8, copy=False)
    dt = (time.perf_counter() - t0) / steps
    same = all(a.shape == b.shape for a, b in zip(ms[:len(pages)], ms[len(pages):])) and bool((ms[len(pages) - 1] == m).all())
    return {"workload": "C4: ModernBookFormatLayoutDetection.predict_masks_batch on 6 page sizes "
                        f"{sizes} x 2 (5 classes, random-init, torch-path semantics), host pages in -> host [5,H,W] masks in pinned staging buffers (copy=False)",
            "e2e_pages_per_sec": len(batch) / dt, "ms_per_page": dt * 1e3 / len(batch),
            "megapixels_per_sec": 2 * sum(h * w for h, w in sizes) / dt / 1e6,
            "per_image_calls_pages_per_sec": len(pages) / dt1, "batch_equals_per_image": same,
            "last_mask_shape": list(m.shape)}


def _c3_worker(rank, n_workers, device, pages_per_worker, steps, threads, barrier, q):
    """One host process of the C3 page->text run: owns a shard of the pages and its own OCRPipeline (CUDA
    context, weights, workspaces) on `device`.  Pages are independent units: no exchange."""
    try:
        import torch

        from ocr_inference_b200 import synthetic
        from ocr_inference_b200.data import LineDetectionConfig, OpStatus
        from ocr_inference_b200.inference import OCRPipeline, get_engine
        from ocr_inference_b200.synthetic_weights import seg_weights_fitted
        from ocr_inference_b200.weights import write_blob

        torch.cuda.set_device(device)
        eng = get_engine(device)
        pipe = OCRPipeline(_ocr_config(), LineDetectionConfig("", write_blob(seg_weights_fitted(11)),
                                                              "deeplabv3plus-r34", 512, ["line"]), None, engine=eng)
        base = [synthetic.text_page(3003 + (rank * 2 + i) % 4) for i in range(2)]
        pages = [base[i % 2][0] for i in range(pages_per_worker)]
        kw = dict(k_factor=2.0, bbox_tolerance=2.5)
        pipe.run_ocr_batch(pages[:max(2, threads)], workers=threads, geometry_procs=0, **kw)
        torch.cuda.synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        n_lines = n_ok = 0
        for _ in range(steps):
            for st, res in pipe.run_ocr_batch(pages, workers=threads, geometry_procs=0, **kw):
                if st == OpStatus.SUCCESS:
                    n_ok += 1
                    n_lines += len(res[2])
        torch.cuda.synchronize()
        q.put((rank, t0, time.perf_counter(), n_ok, n_lines, None))
    except Exception as e:  # noqa: BLE001
        try:
            barrier.abort()
        except Exception:  # noqa: BLE001
            pass
        q.put((rank, 0.0, 0.0, 0, 0, repr(e)))


def bench_c3_procs(device: int, n_workers: int, pages_per_worker: int = 8, steps: int = 2, threads: int = 2):
    """C3 with several host processes per GPU: the cv2 / Python geometry of SURVEY 8(a)-10 is the per-page
    critical path (~25 ms per page per core, GIL-bound inside one process), the GPU stages take ~3-7 ms."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(n_workers), ctx.Queue()
    procs = [ctx.Process(target=_c3_worker, args=(r, n_workers, device, pages_per_worker, steps, threads, barrier, q))
             for r in range(n_workers)]
    for p in procs:
        p.start()
    out, deadline = [], time.time() + 240
    while len(out) < n_workers:
        try:
            out.append(q.get(timeout=2))
        except Exception:  # noqa: BLE001  (queue.Empty)
            dead = [p for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() > deadline:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                return {"error": f"{len(dead)} worker(s) died" if dead else "timeout waiting for the workers"}
    for p in procs:
        p.join(timeout=30)
        if p.is_alive():
            p.terminate()
    errs = [o[5] for o in out if o[5]]
    if errs:
        return {"error": errs[0]}
    t0, t1 = min(o[1] for o in out), max(o[2] for o in out)
    pages = sum(o[3] for o in out)
    lines = sum(o[4] for o in out)
    return {"host_processes": n_workers, "threads_per_process": threads, "pages": pages,
            "e2e_pages_per_sec": pages / (t1 - t0), "e2e_lines_per_sec": lines / (t1 - t0),
            "note": "perf_counter is CLOCK_MONOTONIC (system-wide): first worker start to last worker end"}


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from ocr_inference_b200.data import LineDetectionConfig
    from ocr_inference_b200.inference import LineDetection, get_engine
    from ocr_inference_b200.weights import write_blob

    sampler = ClockSampler(local_rank) if rank == 0 else None   # before the CUDA context exists
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL announces its version on stdout when the communicator comes up; stdout carries the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    eng = get_engine(local_rank)
    dev = eng.device

    from ocr_inference_b200.synthetic_weights import seg_flops_per_tile, seg_weights

    det = LineDetection(LineDetectionConfig("", write_blob(seg_weights(SEG_SEED, 1, 3)), "deeplabv3plus-r34", 512,
                                            ["line"]), engine=eng)
    pages = make_pages(PAGES_PER_STEP, DATA_SEED + 1000 * rank)
    pinned = [torch.from_numpy(p).pin_memory() for p in pages]
    pages_dev = [t.to(dev) for t in pinned]
    torch.cuda.synchronize()

    def step_device():
        return det._predict_device_batch(pages_dev, 0.9)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: HBM-resident inputs -------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    import gc

    gc.collect()
    gc.disable()            # no collector pause inside the timed region (one run showed a 60 ms hiccup over 10 steps)
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        masks = step_device()
    e1.record()
    barrier()
    gc.enable()
    if sampler:
        sampler.mark(w0, time.time())
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    checksum = int(sum(int(m.sum().item()) for m in masks) % (1 << 31))

    # ---- roofline: the same K steps again with a CUDA-event pair around every network launch ----
    # (recorded on the launching stream by bdrc_profile_*; kept out of `value` because ~1400 event
    # records per step cost a few per cent)
    eng.profile(True)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        step_device()
    p1.record()
    barrier()
    ms_profiled = p0.elapsed_time(p1)
    prof = eng.profile_results()
    eng.profile(False)

    # ---- e2e: host numpy pages in, host numpy masks out -----------------------------------
    for _ in range(max(1, args.warmup // 2)):
        det.predict_batch(pages, pinned=pinned, copy=False)
    barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = det.predict_batch(pages, pinned=pinned, copy=False)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if sampler:
        sampler.mark(w0, time.time())

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_pages = PAGES_PER_STEP * args.steps * world
    value = total_pages / (ms / 1e3)
    e2e = total_pages / (e2e_ms / 1e3)
    peak_tf, peak_gbs, peak_src = peaks()
    # roofline of the dominant kernel: every tcgen05 implicit-GEMM conv launch in the timed region
    tc = [(n, t_ms, f) for (n, t_ms, f) in prof if f > 0 and not n.startswith(("stem", "aspp.dw", "dec.dw", "aspp.pool"))]
    tc_ms = sum(x[1] for x in tc)
    tc_flops = sum(x[2] for x in tc)
    groups: dict[str, list[float]] = {}
    for n, t_ms, f in prof:
        g = groups.setdefault(n, [0.0, 0.0, 0])
        g[0] += t_ms
        g[1] += f
        g[2] += 1
    all_ms = sum(x[1] for x in prof)
    achieved = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
    roofline = {
        "bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, all layers)",
        "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "peak_source": peak_src,
        # dram__bytes_read.sum + dram__bytes_write.sum of the five layer-4 launches (512->512 3x3, dilation 2,
        # 148 tiles) in profiles/r01c_conv_full_raw.csv, mean: 378 MB per launch against 315 MB of algorithmic
        # input + output + weight bytes (conv2 of a block also reads the residual: +155 MB on 2 of the 5)
        "traffic": 378.4e6, "traffic_launch": "r34.conv3x3 512->512 @32x32 d2, 148 tiles", "traffic_algorithmic": 315.1e6,
        "launches": len(tc), "avg_launch_ms": tc_ms / max(1, len(tc)),
        "share_of_network_time": tc_ms / all_ms if all_ms else None,
        "share_of_step": tc_ms / ms_profiled if ms_profiled else None,
        "profiled_ms_per_step": ms_profiled / args.steps,
        "by_kernel": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[2] / args.steps,
                          "tflops": (v[1] / (v[0] / 1e3) / 1e12 if v[0] > 0 and v[1] > 0 else None)}
                      for k, v in sorted(groups.items(), key=lambda kv: -kv[1][0])},
    }
    cpu_v, cpu_s, cores = cpu_reference_pages_per_sec(1) if world >= 1 and args.cpu_baseline else (None, None, None)
    extras = {}
    if world == 1 and args.extras:
        try:
            extras["c1"] = bench_c1(eng, max(5, args.steps), 3, args.cpu_baseline)
            extras["c3"] = bench_c3(eng, 2, with_cpu=args.cpu_baseline)
            extras["c4"] = bench_c4(eng)
            nproc = max(2, min(8, len(os.sched_getaffinity(0)) // 2))
            extras["c3"]["multi_process"] = bench_c3_procs(local_rank, nproc, threads=1)
        except Exception as e:  # noqa: BLE001  (the headline must still print)
            extras["error"] = repr(e)
    h2d = sum(p.nbytes for p in pages)
    d2h = sum(o.nbytes for o in out)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "C2: 16 pages 2048x3072 per GPU per step -> 384 tiles 512x512 -> binarise/tile -> "
                               "DeepLabV3+/R34 forward (random-init) -> sigmoid>0.9 mask",
                   "pages_per_step_per_gpu": PAGES_PER_STEP, "tiles_per_page": 24,
                   "algorithmic_gflop_per_page": seg_flops_per_tile(1) * 24 / 1e9,
                   "l2": "inputs larger than L2: 302 MB of pages + >2 GB of activations per step (L2 = 126 MB)",
                   "sharding": "pages sharded across ranks, no collective on the data path"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "api": "LineDetection.predict_batch(host pages, copy=False) -> host masks in pinned staging buffers"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": ({"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                          "sample": "1 page (24 tiles), one pass through oracle/ (cv2 host code + torch-CPU fp32 "
                                    "network), all host threads"} if cpu_v else None),
        "clocks": sampler.summary(),
        "mask_checksum": checksum,
        "other_workloads": extras,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="skip the C1 (lines/s) and C3 (page->text) side measurements of the default N=1 run")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
