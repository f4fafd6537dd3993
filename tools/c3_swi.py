"""C3 under GIL switch intervals and chunk sizes (B200): python tools/c3_swi.py"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from ocr_inference_b200.inference import get_engine  # noqa: E402


def main():
    eng = get_engine(0)
    pages, truth = bench.text_pages(192)
    pipe = bench._pipeline(eng)
    pipe.run_ocr_batch(pages[:96], **bench.PT_KW)
    pipe.run_ocr_batch(pages[:96], **bench.PT_KW)
    for swi, chunk, n in ((0.005, 48, 96), (0.005, 96, 96), (0.005, 32, 96), (0.005, 64, 96), (0.005, 48, 192), (0.005, 96, 192),
                          (0.005, 192, 192), (0.005, 64, 192), (0.005, 48, 96)):
        sys.setswitchinterval(swi)
        kw = dict(bench.PT_KW)
        kw["chunk_pages"] = chunk
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            res = pipe.run_ocr_batch(pages[:n], **kw)
            ts.append(time.perf_counter() - t0)
        ok, n_lines, right = bench._score(res, truth[:n])
        print(f"pages {n} chunk {chunk}: {n / min(ts):.1f} best / {n / sorted(ts)[1]:.1f} median pages/s, "
              f"{right}/{n_lines} lines right", flush=True)
    pipe.close()


if __name__ == "__main__":
    main()
