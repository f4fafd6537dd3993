"""Per-stage wall time of OCRPipeline.run_ocr on synthetic pages (B200): python tools/pipeline_profile.py [pages]"""
import os
import sys
import time
from collections import defaultdict

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ocr_inference_b200 import synthetic  # noqa: E402
from ocr_inference_b200.data import CharsetEncoder, LineDetectionConfig, OCRModelConfig, OpStatus  # noqa: E402
from ocr_inference_b200.inference import OCRPipeline, get_engine  # noqa: E402
from ocr_inference_b200.synthetic_weights import easter2_weights_fitted, seg_weights_fitted  # noqa: E402


def build_pipeline(eng):
    ocr_cfg = OCRModelConfig(easter2_weights_fitted(), "Easter2", 2000, 80, "input", "output", True, False,
                             CharsetEncoder.WYLIE, synthetic.CHARSET_76, True, "fitted")
    line_cfg = LineDetectionConfig("", seg_weights_fitted(11), "deeplabv3plus-r34", 512, ["line"])
    return OCRPipeline(ocr_cfg, line_cfg, None, engine=eng)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    if os.environ.get("CVT"):
        import cv2

        cv2.setNumThreads(int(os.environ["CVT"]))
    if os.environ.get("SWI"):
        sys.setswitchinterval(float(os.environ["SWI"]))
    eng = get_engine(0)
    pipe = build_pipeline(eng)
    pages = [synthetic.text_page(3003 + i)[0] for i in range(n)]
    times = defaultdict(float)
    for name in ("detect_lines", "build_lines", "apply_dewarping", "extract_lines", "run_text_recognition"):
        fn = getattr(pipe, name)

        def wrap(*a, _fn=fn, _name=name, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = _fn(*a, **k)
            torch.cuda.synchronize()
            times[_name] += time.perf_counter() - t0
            return r

        setattr(pipe, name, wrap)
    pipe.run_ocr(pages[0])   # warm-up
    times.clear()
    t0 = time.perf_counter()
    nlines = 0
    for p in pages:
        st, res = pipe.run_ocr(p)
        assert st == OpStatus.SUCCESS, res
        nlines += len(res[2])
    dt = time.perf_counter() - t0
    print(f"{n} pages, {nlines} lines in {dt:.3f}s -> {n / dt:.2f} pages/s, {nlines / dt:.1f} lines/s")
    for k, v in times.items():
        print(f"  {k:22s} {1e3 * v / n:8.1f} ms/page")
    print("  sample text:", res[2][0].text[:60] if res[2] else None)
    for name in ("detect_lines", "build_lines", "apply_dewarping", "extract_lines", "run_text_recognition"):
        delattr(pipe, name)   # drop the timing wrappers (they synchronise the device)
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    batch = [pages[i % n] for i in range(nb)]
    for staged, gp, ws in ((False, 0, (1, 4)), (True, 0, (1, 2, 4)), (True, 4, (4, 8)), (True, 8, (8, 12)), (True, 12, (12,))):
        for workers in ws:
            pipe.run_ocr_batch(batch, workers=workers, staged=staged, geometry_procs=gp)
            t0 = time.perf_counter()
            out = pipe.run_ocr_batch(batch, workers=workers, staged=staged, geometry_procs=gp)
            dt = time.perf_counter() - t0
            nl = sum(len(r[1][2]) for r in out if r[0] == OpStatus.SUCCESS)
            print(f"run_ocr_batch staged={staged} geometry_procs={gp} workers={workers}: {nb / dt:.1f} pages/s, "
                  f"{nl / dt:.0f} lines/s", flush=True)
    # chunk size of the staged schedule (pages resident on the GPU / pinned mask buffers per chunk)
    big = [pages[i % n] for i in range(max(nb, 192))]
    for chunk in (48, 96, 192):
        pipe.run_ocr_batch(big, staged=True, geometry_procs=8, chunk_pages=chunk)
        t0 = time.perf_counter()
        out = pipe.run_ocr_batch(big, staged=True, geometry_procs=8, chunk_pages=chunk)
        dt = time.perf_counter() - t0
        print(f"run_ocr_batch {len(big)} pages, geometry_procs=8, chunk_pages={chunk}: {len(big) / dt:.1f} pages/s", flush=True)
    pipe.close()


if __name__ == "__main__":
    main()
