#!/usr/bin/env python
"""bench.py -- benchmark of the B200 page->text hot path.

Headline workload (BASELINE.json configs[1], "C2"): a batch of 16 synthetic 2048x3072 pages per GPU per
step -> 384 tiles of 512x512 -> binarise/tile (a-1) -> DeepLabV3+/R34 forward (a-2) -> mask
threshold (a-3).  Metric: pages/s (whole job, all ranks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`value`   : inputs (u8 pages) resident in HBM, masks left in HBM, timed with CUDA events.
`e2e`     : LineDetection.predict_batch(pageable numpy pages) -> fresh numpy masks, the reference's calling
            convention (H2D of the pages and D2H of the masks inside the timed region).
`e2e_io_binding`: the same call with caller-pinned inputs and masks left in the object's pinned staging buffers.
`roofline`: tcgen05 implicit-GEMM conv kernel, algorithmic FLOPs / CUDA-event time per launch,
            recorded live during the timed steps, vs MEASURED_PEAKS.json.
`roofline_stages`: the HBM-class stages (a-1, a-3, a-5, a-7, a-9) at step granularity: algorithmic bytes
            / CUDA-event time, vs the measured HBM copy bandwidth.
`cpu_baseline`: the oracle (reference host code restated + torch-CPU fp32 network) on the host
            cores, rank 0, on a bounded sample.
`c1` / `c3` (N = 1) and `c5` (every N): the other BASELINE.json configurations as first-class blocks -- line
            recognition (lines/s), page -> text (pages/s, lines/s, identity with the CPU oracle on 8 pages) and
            the page-sharded page -> text run (sharding.run_sharded: every rank drives its share of the pages
            through OCRPipeline.run_ocr_batch, rank 0 gathers the per-page results on the host).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
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
        try:
            ut = float(nv.nvmlDeviceGetUtilizationRates(h).gpu)
        except Exception:
            ut = -1.0
        print("s", time.time(), mhz, mask, pw, ut, flush=True)
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
        self.windows: list[tuple[float, float, str]] = []
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(phys)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception as e:  # noqa: BLE001
            self.proc, self.err = None, repr(e)

    def mark(self, t0: float, t1: float, tag: str = "headline"):
        self.windows.append((t0, t1, tag))

    def _finish(self):
        if getattr(self, "_rows", None) is not None:
            return
        self._rows, self._max, self._err = [], None, None
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
            out, _ = self.proc.communicate()
        for ln in out.splitlines():
            f = ln.split()
            if f and f[0] == "max":
                self._max = float(f[1])
            elif f and f[0] == "s" and len(f) >= 5:
                self._rows.append((float(f[1]), float(f[2]), int(f[3]), float(f[4]), float(f[5]) if len(f) > 5 else -1.0))
            elif f and f[0] == "err":
                self._err = ln

    def gpu_util(self, tag: str):
        """Median NVML GPU utilisation (% of time a kernel was running) inside the windows tagged `tag`."""
        self._finish()
        u = sorted(r[4] for r in self._rows if r[4] >= 0 and any(a <= r[0] <= b and t == tag for a, b, t in self.windows))
        return u[len(u) // 2] if u else None

    def summary(self, tag: str = "headline"):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"sampler unavailable: {self.err}"]}
        self._finish()
        inside = [r for r in self._rows if any(a <= r[0] <= b and t == tag for a, b, t in self.windows)]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self._max,
                    "reasons": [f"no samples inside the timed region ({self._err})"], "samples": 0}
        sm = sorted(r[1] for r in inside)
        seen = 0
        for r in inside:
            seen |= r[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self._max,
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


PT_KW = dict(k_factor=2.0, bbox_tolerance=2.5)        # the reference's Demo-OCR.ipynb parameters


def text_pages(n: int, seed0: int = 3003, n_base: int = 8):
    """n synthetic text pages + their transcriptions: `n_base` generated pages, the rest are
    augmentations (vertical shift by a few rows, brightness) -- the text of a page stays what it was."""
    from ocr_inference_b200 import synthetic

    base = [synthetic.text_page(seed0 + i) for i in range(min(n, n_base))]
    pages, truth = [], []
    for i in range(n):
        img, _, texts = base[i % len(base)]
        k = i // len(base)
        if k:
            img = np.roll(img, ((k * 7) % 41) - 20, axis=0)
            img = np.clip(img.astype(np.int16) + ((k * 3) % 9 - 4), 0, 255).astype(np.uint8)
        pages.append(np.ascontiguousarray(img))
        truth.append(["".join(synthetic.CHARSET_76[j] for j in t).strip().replace("\u00a7", " ") for t in texts])
    return pages, truth


def cpu_reference_page_to_text(pages, gpu_results):
    """The oracle's whole page -> text pipeline (tile/binarise, torch-CPU fp32 DeepLabV3+ with the trained decoder,
    mask, the reference's cv2/numpy line extraction, line prep, torch-CPU Easter2, greedy CTC) on `pages`, on all
    host cores, beside the GPU pipeline's results for the same pages."""
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
    n_lines = same = same_box = gpu_lines = pages_same_count = 0
    max_box = 0
    dist = chars = 0
    t0 = time.perf_counter()
    per_page = []
    for page in pages:
        tiles, meta = ho.seg_preprocess(page, closed_form=False)
        x = np.repeat(tiles[:, None].astype(np.float32), 3, axis=1)
        logits = np.concatenate([nets.seg_forward(ws, x[i:i + 8]) for i in range(0, len(x), 8)])
        mask = ho.line_postprocess(logits, meta, 0.9, closed_form=False)
        r = og.page_to_lines(page, mask, PT_KW["k_factor"], PT_KW["bbox_tolerance"])
        texts = [] if isinstance(r, str) else og.recognise(r[2], we, synthetic.CHARSET_76)[0]
        per_page.append((r, texts))
    dt = time.perf_counter() - t0
    for (r, texts), g in zip(per_page, gpu_results):
        want = [t.strip().replace("\u00a7", " ") for t in texts]
        n_lines += len(want)
        if g[0] != OpStatus.SUCCESS:
            continue
        got = [l.text for l in g[1][2]]
        gpu_lines += len(got)
        if len(got) != len(want):
            continue
        pages_same_count += 1
        same += sum(a == b for a, b in zip(got, want))
        dist += sum(_edit_distance(a, b) for a, b in zip(got, want))
        chars += sum(max(len(a), len(b)) for a, b in zip(got, want))
        if not isinstance(r, str):
            gb = [(l.bbox.x, l.bbox.y, l.bbox.w, l.bbox.h) for l in g[1][1]]
            ob = [tuple(l.bbox) for l in r[1]]
            if len(gb) == len(ob):
                same_box += sum(a == b for a, b in zip(gb, ob))
                max_box = max([max_box] + [max(abs(p_ - q_) for p_, q_ in zip(a, b)) for a, b in zip(gb, ob)])
    return {"value": len(pages) / dt, "unit": "pages/s", "lines_per_sec": n_lines / dt, "cores": cores, "kind": "port",
            "sample": f"{len(pages)} of the pages: oracle/ page -> text (cv2/numpy host code of the reference restated, "
                      "torch-CPU fp32 networks), all host threads", "seconds_per_page": dt / len(pages), "lines": n_lines,
            "gpu_lines": gpu_lines, "pages_with_equal_line_count": pages_same_count,
            "gpu_lines_identical_to_cpu_oracle": same,
            "identical_line_fraction": same / max(1, n_lines),
            "gpu_line_boxes_identical": same_box, "gpu_line_boxes_max_abs_diff_px": max_box,
            "gpu_vs_cpu_character_agreement": (1.0 - dist / chars) if chars else None,
            "note": "Easter2 internals are our restatement of arXiv 2205.14879 (parity unpinned against BDRC's ONNX "
                    "files); greedy CTC on both sides"}


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
    # a step lasts under a millisecond: three timed loops of >= 30 steps each, the median is reported (one loop of
    # 10 steps was at the mercy of a single host hiccup: 0.92 ms against 0.74 ms for the same build on one box)
    steps = max(30, steps)
    loops = []
    for _ in range(3):
        l0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_device()
        e1.record()
        torch.cuda.synchronize()
        loops.append(e0.elapsed_time(e1) / steps)
        launches = (eng.launch_count() - l0) // steps
    ms = sorted(loops)[1]
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
    # SURVEY 8(d) C1, fixed-size variant: 64 crops that already are 80 x 2000 (the resize of a-7 is the identity)
    fx_dev, fx_hw, fx_off = eng.pack_crops(synthetic.line_crops(1001, C1_LINES, fixed=(80, 2000)))

    def step_fixed():
        lines, _ = eng.ocr_prepare_lines(fx_dev, fx_hw, fx_off, C1_LINES, 80, 2000, False)
        return eng.ctc_greedy(eng.ocr_forward(ocr.ocr_session, lines), blank=blank)

    for _ in range(3):
        step_fixed()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(steps):
        step_fixed()
    f1.record()
    torch.cuda.synchronize()
    ms_fixed = f0.elapsed_time(f1) / steps
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
        "timed": f"median of 3 loops of {steps} steps, CUDA events", "ms_per_step_loops": [round(x, 4) for x in loops],
        "lines_per_sec_batch_of_512": L8 / (ms_big / 1e3), "batch_of_512_equals_batch_of_64": same_big,
        "lines_per_sec_fixed_size_80x2000": C1_LINES / (ms_fixed / 1e3), "ms_per_step_fixed_size_80x2000": ms_fixed,
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


def _pipeline(eng):
    from ocr_inference_b200.data import LineDetectionConfig
    from ocr_inference_b200.inference import OCRPipeline
    from ocr_inference_b200.synthetic_weights import seg_weights_fitted
    from ocr_inference_b200.weights import write_blob

    return OCRPipeline(_ocr_config(), LineDetectionConfig("", write_blob(seg_weights_fitted(11)), "deeplabv3plus-r34",
                                                          512, ["line"]), None, engine=eng)


def _score(results, truth):
    """(pages ok, lines found, lines equal to the generator's transcription)."""
    from ocr_inference_b200.data import OpStatus

    ok = n_lines = right = 0
    for r, want in zip(results, truth):
        if r[0] != OpStatus.SUCCESS:
            continue
        ok += 1
        got = [l.text for l in r[1][2]]
        n_lines += len(got)
        pool = list(want)
        for g in got:
            if g in pool:
                pool.remove(g)
                right += 1
    return ok, n_lines, right


def bench_c3(eng, steps: int, pages_per_step: int = 96, workers: int | None = None, with_cpu: bool = False):
    """Full page -> text: OCRPipeline.run_ocr_batch on synthetic text pages (pageable numpy in, OCRLine text out)."""
    import torch

    pipe = _pipeline(eng)
    pages, truth = text_pages(pages_per_step)
    n_truth = sum(len(t) for t in truth)
    for _ in range(2):                                       # warm-up: pinned mask buffers, allocator pools, workers
        pipe.run_ocr_batch(pages, workers=workers, **PT_KW)
    l0 = eng.launch_count()
    step_s = []
    for _ in range(steps):
        t0 = time.perf_counter()
        res = pipe.run_ocr_batch(pages, workers=workers, **PT_KW)
        step_s.append(time.perf_counter() - t0)
    dt = sorted(step_s)[len(step_s) // 2]                   # median step (host-side jitter: worker processes, GIL hand-offs)
    launches = (eng.launch_count() - l0) // steps
    ok, n_lines, n_right = _score(res, truth)
    # the conv kernels' share of the page -> text wall time and their rate, from one more (profiled) pass
    eng.profile(True)
    tp = time.perf_counter()
    pipe.run_ocr_batch(pages, workers=workers, **PT_KW)
    torch.cuda.synchronize()
    dtp = time.perf_counter() - tp
    prof = eng.profile_results()
    eng.profile(False)
    tc = [(n, t, f) for (n, t, f) in prof if f > 0 and not n.startswith(("stem", "aspp.dw", "dec.dw", "aspp.pool"))]
    tc_ms, tc_fl = sum(x[1] for x in tc), sum(x[2] for x in tc)
    gpu_ms = sum(x[1] for x in prof)
    peak_tf, _, peak_src = peaks()
    cpu = cpu_reference_page_to_text(pages[:8], res[:8]) if with_cpu else None
    n_geo = len(getattr(pipe, "_gpool", None) or [])
    pipe.close()
    ach = tc_fl / (tc_ms / 1e3) / 1e12 if tc_ms else 0.0
    return {
        "workload": f"C3: {pages_per_step} synthetic 2048x3072 text pages per step through OCRPipeline.run_ocr_batch "
                    "(pageable numpy in; staged: pages grouped 6 at a time through tile/binarise + DeepLabV3+ (trained "
                    "decoder) + mask, cv2 contours + reading order in geometry worker processes (no CUDA in them), GPU "
                    "line crops, all lines of the batch through line prep + Easter2 + greedy CTC in one call; ONE CUDA "
                    "context) -> OCRLine text",
        "metric": "pages_per_sec", "value": pages_per_step / dt, "unit": "pages/s",
        "e2e": {"value": pages_per_step / dt, "unit": "pages/s", "lines_per_sec": n_lines / dt,
                "h2d_bytes_per_step": int(sum(p.nbytes for p in pages)), "d2h_bytes_per_step": int(sum(p.shape[0] * p.shape[1] for p in pages)),
                "api": "OCRPipeline.run_ocr_batch(list of pageable HWC u8 pages) -> [(OpStatus, [mask, lines, OCRLine, angle])]"},
        "ms_per_page": dt * 1e3 / pages_per_step, "gpu_launches_per_step": int(launches),
        "steps": steps, "step_seconds": [round(x, 4) for x in step_s], "statistic": "median step",
        "best_step_pages_per_sec": pages_per_step / min(step_s),
        "host_threads": workers, "geometry_processes": n_geo, "host_cores": len(os.sched_getaffinity(0)),
        "pages_ok": ok, "lines_found": n_lines, "lines_on_pages": n_truth, "lines_equal_to_truth": n_right,
        "lines_equal_to_truth_fraction": n_right / max(1, n_truth),
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (both networks) inside the page -> text run",
                     "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "peak_source": peak_src,
                     "conv_ms_per_page": tc_ms / pages_per_step, "network_kernel_ms_per_page": gpu_ms / pages_per_step,
                     "share_of_wall": tc_ms / 1e3 / dtp if dtp else None},
        **({"cpu_baseline": cpu} if cpu else {}),
    }


def _proc_tree_cpu_seconds():
    """user + system CPU seconds of this process and its live children (geometry workers)."""
    try:
        import psutil

        me = psutil.Process()
        tot = sum(me.cpu_times()[:2])
        for c in me.children(recursive=True):
            try:
                tot += sum(c.cpu_times()[:2])
            except Exception:  # noqa: BLE001
                pass
        return tot
    except Exception:  # noqa: BLE001
        t = os.times()
        return t.user + t.system


def bench_c5(eng, rank: int, world: int, sampler, steps: int = 3, pages_per_rank: int = 192):
    """Config C5, time-bounded: the page-sharded page -> text run.  `pages_per_rank` x world synthetic text pages
    (the C3 generator) are split into contiguous per-rank ranges by sharding.run_sharded; every rank pushes its
    range through its own OCRPipeline.run_ocr_batch (pageable numpy in) and rank 0 gathers the per-page line
    boxes + texts on the host.  No collective on the data path.  Timed with a barrier on both sides, max over
    ranks; CPU seconds of every rank's process tree and the H2D byte rate say what limits it."""
    import torch
    import torch.distributed as dist

    from ocr_inference_b200 import sharding
    from ocr_inference_b200.data import OpStatus

    cores = len(os.sched_getaffinity(0))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    geo = max(1, min(8, cores // max(1, local_world) - 2))
    pipe = _pipeline(eng)
    n_pages = pages_per_rank * world
    mine = sharding.shard_range(n_pages, world, rank)
    # every rank materialises only its own shard (generated on the fly per shard)
    shard_pages, shard_truth = text_pages(len(mine), seed0=3003 + 16 * rank)
    offset = mine.start

    def load(i):
        return shard_pages[i - offset]

    local_out = []

    def process(pages):
        out = []
        for st, res in pipe.run_ocr_batch(pages, geometry_procs=geo, **PT_KW):
            if st == OpStatus.SUCCESS:
                out.append(("ok", [((l.bbox.x, l.bbox.y, l.bbox.w, l.bbox.h), o.text) for l, o in zip(res[1], res[2])]))
            else:
                out.append(("failed", str(res)))
        local_out.extend(out)
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sharding.run_sharded(n_pages, load, process, chunk=96)      # warm-up at full size: pinned buffers, worker processes
    barrier()
    if rank == C5_FAULT_RANK:
        pipe.close()
        raise RuntimeError("injected C5 failure (BDRC_BENCH_C5_FAULT)")
    cpu0, w0 = _proc_tree_cpu_seconds(), time.time()
    t0 = time.perf_counter()
    for _ in range(steps):
        results = sharding.run_sharded(n_pages, load, process, chunk=96)
    barrier()
    dt = time.perf_counter() - t0
    cpu_s = _proc_tree_cpu_seconds() - cpu0
    if sampler:
        sampler.mark(w0, time.time(), "c5")
    # every rank scores its own shard of the last step against the generator's transcriptions
    ok = n_lines = right = 0
    for res, want in zip(local_out[-len(mine):], shard_truth):
        if res[0] != "ok":
            continue
        ok += 1
        n_lines += len(res[1])
        pool = list(want)
        for _, txt in res[1]:
            if txt in pool:
                pool.remove(txt)
                right += 1
    t = torch.tensor([dt, cpu_s, ok, n_lines, right, sum(len(t_) for t_ in shard_truth)], dtype=torch.float64,
                     device=eng.device)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dt_max, cpu_total = float(tmax[0]), float(t[1])
    local_ok, local_lines, local_right, n_truth = int(t[2]), int(t[3]), int(t[4]), int(t[5])
    n_geo = len(getattr(pipe, "_gpool", None) or [])
    pipe.close()
    if rank != 0:
        return None
    assert results is not None and len(results) == n_pages
    pages_done = n_pages * steps
    busy = cpu_total / dt_max
    h2d = pages_done * 2048 * 3072 * 3 / dt_max / 1e9
    util = sampler.gpu_util("c5") if sampler else None
    if busy >= 0.75 * cores:
        limiter = f"host cores: {busy:.1f} of {cores} busy (cv2 geometry, page copies into pinned memory, Python dispatch)"
    elif util is not None and util >= 85:
        limiter = f"GPU: a kernel was running {util:.0f} % of the time on rank 0's device"
    else:
        limiter = (f"the GPU-owning Python process of each rank (stage dispatch under the GIL): {busy:.1f} of {cores} "
                   f"cores busy, GPU kernel-active {util if util is not None else float('nan'):.0f} %")
    return {
        "workload": f"C5 (time-bounded): {n_pages} synthetic 2048x3072 text pages per step ({pages_per_rank} per GPU), "
                    f"sharded over {world} rank(s) by sharding.run_sharded -> OCRPipeline.run_ocr_batch per rank, per-page "
                    "(line boxes, texts) gathered on rank 0 by gather_object; pageable numpy pages in",
        "pages_per_sec": pages_done / dt_max, "lines_per_sec": local_lines * steps / dt_max,
        "n_gpus": world, "pages_per_step": n_pages, "steps": steps, "seconds": dt_max,
        "host_cores": cores, "host_cores_busy": busy, "host_cores_busy_per_gpu": busy / world,
        "geometry_procs_per_gpu": n_geo, "h2d_gbytes_per_sec": h2d, "gpu_kernel_active_pct_rank0": util,
        "pages_ok": local_ok, "lines_found": local_lines, "lines_on_pages": n_truth,
        "lines_equal_to_truth_fraction": local_right / max(1, n_truth),
        "limiter": limiter, "scaling": "weak", "collectives_on_data_path": 0,
    }


def roofline_stages(eng):
    """The HBM-class stages at step granularity (what one C2 / C1 step launches), CUDA events around each call,
    inputs rotated through sets larger than the 126 MB L2.  bytes = SURVEY 8(d)'s algorithmic bytes."""
    import cv2
    import torch

    from ocr_inference_b200 import line_detection as ld
    from ocr_inference_b200 import synthetic
    from ocr_inference_b200.engine import LAYOUT_TV, POST_SIGMOID

    _, hbm, src = peaks()
    dev = eng.device
    H, W = PAGE_H, PAGE_W
    out = {}

    def timeit(fn, sets, reps=12, warm=3, kernels=None):
        """(ms per call between two CUDA events around `reps` calls -- host call overhead included --,
        ms per call of the launches whose profile name starts with `kernels`, from the ABI's own event pairs around
        each launch on the launching stream; None when the stage has no profiled launch)"""
        for i in range(warm):
            fn(sets[i % len(sets)])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(sets[i % len(sets)])
        e1.record()
        torch.cuda.synchronize()
        call_ms = e0.elapsed_time(e1) / reps
        kern_ms = None
        if kernels:
            eng.profile(True)
            for i in range(reps):
                fn(sets[i % len(sets)])
            torch.cuda.synchronize()
            res = [ms for name, ms, _ in eng.profile_results() if name.startswith(kernels)]
            eng.profile(False)
            if res:
                kern_ms = sum(res) / reps
        return call_ms, kern_ms

    def add(key, what, t, nbytes, launches):
        # the roofline figure is the KERNEL time where the ABI profiles the launch (a one-launch stage called from
        # Python is otherwise timed at the wrapper's speed: ~90 us of argument marshalling for a 16-page table)
        call_ms, kern_ms = t
        ms = kern_ms if kern_ms else call_ms
        gbs = nbytes / (ms / 1e3) / 1e9
        out[key] = {"what": what, "bytes": int(nbytes), "us": ms * 1e3, "us_call": call_ms * 1e3,
                    "timed": "kernel (CUDA events around the launches)" if kern_ms else "call (host overhead included)",
                    "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "launches": launches}

    base = make_pages(4, DATA_SEED)
    psets = [[torch.from_numpy(np.roll(base[(i + j) % 4], 31 * (i + 4 * j), axis=1).copy()).to(dev) for i in range(16)]
             for j in range(2)]                                     # 2 x 302 MB
    ms = timeit(lambda ps: eng.seg_preprocess_batch(ps), psets, kernels="a-1")
    add("a-1", "gray + 51-tap Gaussian adaptive threshold + tile, 16 pages 2048x3072x3 -> 384 u8 tiles, ONE launch "
               "(fp32-FMA bound by construction: 102 flop per pixel)", ms, 16 * 5 * H * W, 1)
    metas = [dict(H0=H, W0=W, Hc=H, Wc=W, nty=4, ntx=6, pad_x=0, pad_y=0) for _ in range(16)]
    # text-band-like logits: wide bands far above / below the threshold with soft edges (what a line model emits)
    yy = torch.arange(128 * 4, device=dev, dtype=torch.float32)[:, None]
    band = (torch.sin(yy / 9.0) * 14.0).expand(-1, 128 * 6)
    lsets = []
    for j in range(3):
        page_low = band + torch.randn((512, 768), device=dev) * 0.5 + j
        tiles = page_low.reshape(4, 128, 6, 128).permute(0, 2, 1, 3).reshape(24, 128, 128, 1)
        lsets.append(tiles.repeat(16, 1, 1, 1).contiguous())         # [384,128,128,1]
    ms = timeit(lambda lg: eng.seg_postprocess_lowres_batch(lg, metas, POST_SIGMOID, 0.9), lsets, kernels="a-3")
    add("a-3", "line masks of 16 pages from low-res logits (x4 upsample + sigmoid>0.9 + stitch fused), ONE launch",
        ms, 16 * (24 * 128 * 128 * 4 + H * W), 1)
    page_h, boxes, _ = synthetic.text_page(3003)
    bandm = synthetic.band_mask(boxes, H, W, grow=2, thin=0.45)
    contours, _ = cv2.findContours(bandm, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    rois, masks, hs = [], [], []
    for c in contours:
        roi, m, h = ld._roi_mask(c, H, W)
        rois.append(roi), masks.append(m), hs.append(h)
    pack = eng.pack_masks(masks, rois)
    ses = [ld._se_for(h, 2.0) for h in hs]
    idx = list(range(len(rois)))
    tp = [torch.from_numpy(np.roll(page_h, 13 * i, axis=1).copy()).to(dev) for i in range(8)]
    crops, hw, off = eng.extract_lines(tp[0], pack, idx, ses)
    hw_h = hw.cpu().numpy()
    crop_px = int((hw_h[:, 0].astype(np.int64) * hw_h[:, 1]).sum())
    l0 = eng.launch_count()
    eng.extract_lines(tp[0], pack, idx, ses)
    n_l = eng.launch_count() - l0
    ms = timeit(lambda p_: eng.extract_lines(p_, pack, idx, ses), tp)
    add("a-5", f"masked crops of the {len(idx)} lines of one page ({crop_px / 1e6:.1f} Mpx): SAT dilation, AND, empty "
               "row/column deletion, gather", ms, 6 * crop_px, int(n_l))
    L = C1_LINES
    csets = []
    for j in range(8):
        cs, _ = synthetic.glyph_lines(1001 + j, L)
        csets.append(eng.pack_crops(cs, key=f"rs{j}") + (sum(c.shape[0] * c.shape[1] for c in cs),))
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    eng.ocr_prepare_lines(csets[0][0], csets[0][1], csets[0][2], L, 80, 2000, False)
    n_l = eng.launch_count() - l0
    ms = timeit(lambda c: eng.ocr_prepare_lines(c[0], c[1], c[2], L, 80, 2000, False), csets, kernels="a-7")
    add("a-7", "64 ragged crops -> cv2-exact resize, pad, adaptive threshold -> bf16 [64,2000,128]", ms,
        3 * csets[0][3] + L * 2 * 80 * 2000 * 128 // 80, int(n_l))
    lg = [torch.randn((L, 500, 77), device=dev) for _ in range(16)]
    ms = timeit(lambda x: eng.ctc_greedy(x, layout=LAYOUT_TV), lg, kernels="a-9")
    add("a-9", "argmax + CTC collapse, 64 lines [500,77] fp32 (C1 size)", ms, L * 500 * 77 * 4, 1)
    big = [torch.randn((1024, 500, 77), device=dev) for _ in range(2)]
    ms = timeit(lambda x: eng.ctc_greedy(x, layout=LAYOUT_TV), big, reps=8, kernels="a-9")
    add("a-9@1024", "argmax + CTC collapse, 1024 lines (a page batch)", ms, 1024 * 500 * 77 * 4, 1)
    out["peak_source"] = src
    return out


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


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
C5_DEADLINE_S = int(os.environ.get("BDRC_C5_DEADLINE_S", "240"))
C5_FAULT_RANK = int(os.environ.get("BDRC_BENCH_C5_FAULT", "-1"))      # test hook: this rank fails inside C5 (tools/c5_fault.sh)


def _kill_children():
    """os._exit skips multiprocessing's atexit hook: end the geometry workers by hand so that no orphan keeps the
    job's stdout / stderr pipes open."""
    try:
        import multiprocessing as mp

        for c in mp.active_children():
            c.kill()
    except Exception:  # noqa: BLE001
        pass


class Deadline:
    """Arms a timer: unless cancel() comes first, `on_timeout` runs (if given) and the PROCESS exits with code 0.
    seconds=None arms nothing."""

    def __init__(self, seconds, on_timeout=None, _exit=os._exit):
        self._timer = None
        self._lock = threading.Lock()
        self._state = "armed"
        if seconds is None:
            return

        def fire():
            with self._lock:
                if self._state != "armed":
                    return
                self._state = "fired"
            try:
                if on_timeout:
                    on_timeout()
            except Exception as e:  # noqa: BLE001
                print(f"[bench] deadline callback failed: {e!r}", file=sys.stderr)
            finally:
                sys.stdout.flush()
                sys.stderr.flush()
                _kill_children()
                _exit(0)

        self._timer = threading.Timer(seconds, fire)
        self._timer.daemon = True
        self._timer.start()

    def cancel(self):
        """Disarm.  If the deadline is already firing, the caller parks here until the process exit it ends with: the
        guarded section must not go on to print a second line."""
        with self._lock:
            fired = self._state == "fired"
            if not fired:
                self._state = "cancelled"
        if self._timer is not None:
            self._timer.cancel()
        if fired:
            threading.Event().wait(600.0)


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

    # ---- e2e: the reference's calling convention -- pageable numpy pages in, fresh numpy masks out ------------
    for _ in range(max(1, args.warmup // 2)):
        det.predict_batch(pages)
    barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = det.predict_batch(pages)                       # pinned=None (pageable route), copy=True
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if sampler:
        sampler.mark(w0, time.time())
    assert all(isinstance(o, np.ndarray) and o.flags["OWNDATA"] for o in out)
    # ---- the io-binding style of use: caller-pinned inputs, masks left in the object's pinned staging buffers ----
    det.predict_batch(pages, pinned=pinned, copy=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        det.predict_batch(pages, pinned=pinned, copy=False)
    torch.cuda.synchronize()
    iob_s = time.perf_counter() - t0

    t = torch.tensor([ms, e2e_s * 1e3, iob_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, iob_ms = float(t[0]), float(t[1]), float(t[2])

    def finish(c5_block, side: bool):
        _emit_line(args, world, eng, sampler, pages, out, prof, ms, ms_profiled, e2e_ms, iob_ms, launches, checksum,
                   c5_block, side)

    # ---- C5: the page-sharded page -> text run, at every N ----------------------------------------------
    # Under torchrun a rank that fails or stalls inside C5 would leave the others waiting in a collective and the job
    # without its JSON line; every rank therefore arms a deadline: when it passes, rank 0 prints the headline with
    # c5 = {"error": ...} and every rank leaves with exit code 0.
    c5 = None
    c5_failed = False
    if args.extras:
        guard = Deadline(C5_DEADLINE_S if world > 1 else None,
                         (lambda: finish({"error": f"C5 did not finish within {C5_DEADLINE_S} s on {world} ranks"},
                                         side=False)) if rank == 0 else None)
        try:
            c5 = bench_c5(eng, rank, world, sampler)
        except Exception as e:  # noqa: BLE001  (the headline must still print)
            c5 = {"error": repr(e)}
            c5_failed = True
            print(f"[bench rank {rank}] C5 failed: {e!r}", file=sys.stderr, flush=True)
        finally:
            guard.cancel()
    if rank != 0:
        if world > 1:
            if c5_failed:
                _kill_children()
                os._exit(0)          # peers may be stuck in a collective of C5: no orderly shutdown to wait for
            dist.destroy_process_group()
        return
    finish(c5, side=True)
    if world > 1:
        if c5_failed:
            sys.stdout.flush()
            _kill_children()
            os._exit(0)
        dist.destroy_process_group()


def _emit_line(args, world, eng, sampler, pages, out, prof, ms, ms_profiled, e2e_ms, iob_ms, launches, checksum, c5,
               side: bool):
    """Rank 0: assemble and print the ONE JSON line.  side=False (C5 deadline path) skips every further measurement."""
    from ocr_inference_b200.synthetic_weights import seg_flops_per_tile

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
        # 148 tiles) in profiles/r02c_seg_convs_full_raw.csv (IDs 31-35), mean: 381 MB per launch against 315 MB of
        # algorithmic input + output + weight bytes (conv2 of a block also reads the residual: +155 MB on 3 of the 5)
        "traffic": 381.0e6, "traffic_launch": "r34.conv3x3 512->512 @32x32 d2, 148 tiles", "traffic_algorithmic": 315.1e6,
        "launches": len(tc), "avg_launch_ms": tc_ms / max(1, len(tc)),
        "share_of_network_time": tc_ms / all_ms if all_ms else None,
        "share_of_step": tc_ms / ms_profiled if ms_profiled else None,
        "profiled_ms_per_step": ms_profiled / args.steps,
        "whole_step_frac": (seg_flops_per_tile(1) * 24 * PAGES_PER_STEP / ((ms / args.steps) / 1e3) / 1e12) / peak_tf,
        "by_kernel": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[2] / args.steps,
                          "tflops": (v[1] / (v[0] / 1e3) / 1e12 if v[0] > 0 and v[1] > 0 else None)}
                      for k, v in sorted(groups.items(), key=lambda kv: -kv[1][0])},
    }
    cpu_v, cpu_s, cores = cpu_reference_pages_per_sec(1) if side and args.cpu_baseline else (None, None, None)
    c1 = c3 = c4 = stages = None
    errors = {}
    if args.extras and side:
        for name, fn in (("roofline_stages", lambda: roofline_stages(eng)),
                         ("c1", lambda: bench_c1(eng, max(5, args.steps), 3, args.cpu_baseline) if world == 1 else None),
                         ("c3", lambda: bench_c3(eng, 5, with_cpu=args.cpu_baseline) if world == 1 else None),
                         ("c4", lambda: bench_c4(eng) if world == 1 else None)):
            try:
                v = fn()
            except Exception as e:  # noqa: BLE001  (the headline must still print)
                v = None
                errors[name] = repr(e)
            if name == "roofline_stages":
                stages = v
            elif name == "c1":
                c1 = v
            elif name == "c3":
                c3 = v
            else:
                c4 = v
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
                "ms_per_step": e2e_ms / args.steps,
                "api": "LineDetection.predict_batch(list of pageable numpy pages) -> list of fresh numpy masks (the "
                       "reference's ownership: pageable in, caller-owned arrays out)"},
        "e2e_io_binding": {"value": total_pages / (iob_ms / 1e3), "unit": UNIT, "ms_per_step": iob_ms / args.steps,
                           "api": "predict_batch(pages, pinned=caller-pinned tensors, copy=False): masks stay in the "
                                  "object's pinned staging buffers (onnxruntime io-binding style of use)"},
        "gpu_launches": int(launches),
        "simt_conv_launches": int(eng.simt_launch_count()),
        "conv_plans_built": int(eng.conv_plans_built()),
        "roofline": roofline,
        "roofline_stages": stages,
        "cpu_baseline": ({"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                          "sample": "1 page (24 tiles), one pass through oracle/ (cv2 host code + torch-CPU fp32 "
                                    "network), all host threads"} if cpu_v else None),
        "clocks": sampler.summary(),
        "mask_checksum": checksum,
        "c1": c1, "c3": c3, "c5": c5,
        "other_workloads": {"c4": c4, "c5": c5, "c1": "top-level key c1", "c3": "top-level key c3",
                            **({"errors": errors} if errors else {})},
    }
    print(json.dumps(line), flush=True)


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
