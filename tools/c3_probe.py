"""C3 under variants (B200): python tools/c3_probe.py   -- census on/off; BDRC_SPIN_EVENTS=1 in the environment for spinning waits"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from ocr_inference_b200.inference import get_engine  # noqa: E402


def main():
    eng = get_engine(0)
    pages, truth = bench.text_pages(96)
    for census in (True, False, True, False):
        pipe = bench._pipeline(eng)
        pipe.line_inference.with_census = census
        pipe.run_ocr_batch(pages, **bench.PT_KW)
        t0 = time.perf_counter()
        for _ in range(3):
            res = pipe.run_ocr_batch(pages, **bench.PT_KW)
        dt = (time.perf_counter() - t0) / 3
        ok, n_lines, right = bench._score(res, truth)
        print(f"census={census} spin={os.environ.get('BDRC_SPIN_EVENTS', '0')}: {96 / dt:.1f} pages/s, {right}/{n_lines} lines right", flush=True)
        pipe.close()


if __name__ == "__main__":
    main()
