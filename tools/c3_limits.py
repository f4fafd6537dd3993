"""Where does run_ocr_batch spend a page?  Upper bounds by switching parts off: python tools/c3_limits.py  (B200)
  full              : the real thing (8 geometry processes)
  cached geometry   : page_geometry answered from a cache -> what the GPU-owning process alone can do
"""
import os
import sys
import time
from concurrent.futures import Future

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ocr_inference_b200 import synthetic  # noqa: E402
from ocr_inference_b200.inference import get_engine  # noqa: E402
from tools.pipeline_profile import build_pipeline  # noqa: E402


def main():
    pipe = build_pipeline(get_engine(0))
    pages = [synthetic.text_page(3003 + i)[0] for i in range(8)]
    batch = [pages[i % 8] for i in range(192)]

    def run(label, **kw):
        pipe.run_ocr_batch(batch, **kw)
        t0 = time.perf_counter()
        pipe.run_ocr_batch(batch, **kw)
        dt = time.perf_counter() - t0
        print(f"{label}: {len(batch) / dt:.1f} pages/s ({1e3 * dt / len(batch):.2f} ms per page)", flush=True)

    run("full, 8 geometry processes", geometry_procs=8)
    gp = pipe._gpool
    real_submit = gp.submit
    cache = {}

    def cached_submit(mask, merge_lines=True):
        key = (mask.shape, int(mask[::97, ::89].sum()))
        if key not in cache:
            cache[key] = real_submit(mask, merge_lines).result()
        f = Future()
        f.set_result(cache[key])
        return f

    gp.submit = cached_submit
    run("cached geometry (GPU-owning process alone), 8 page threads", geometry_procs=8)
    run("cached geometry, 4 page threads", geometry_procs=8, workers=4)
    run("cached geometry, 2 page threads", geometry_procs=8, workers=2)
    gp.submit = real_submit
    pipe.close()


if __name__ == "__main__":
    main()
