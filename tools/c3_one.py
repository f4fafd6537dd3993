"""One C3 measurement (B200): python tools/c3_one.py [tag] [geometry_procs]  -- pages/s of run_ocr_batch(96 pages), CPU cores busy"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from ocr_inference_b200.inference import get_engine  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else ""
    geo = int(sys.argv[2]) if len(sys.argv) > 2 else None
    if os.environ.get("CVT"):
        import cv2

        cv2.setNumThreads(int(os.environ["CVT"]))
    eng = get_engine(0)
    pages, truth = bench.text_pages(96)
    pipe = bench._pipeline(eng)
    for _ in range(2):
        pipe.run_ocr_batch(pages, geometry_procs=geo, **bench.PT_KW)
    import threading

    import psutil

    me = psutil.Process()

    def snap():
        names = {t.native_id: t.name for t in threading.enumerate()}
        th = {t.id: (t.user_time + t.system_time, names.get(t.id, "native")) for t in me.threads()}
        ch = {c.pid: sum(c.cpu_times()[:2]) for c in me.children(recursive=True)}
        return th, ch

    th0, ch0 = snap()
    c0 = bench._proc_tree_cpu_seconds()
    t0 = time.perf_counter()
    steps = []
    for _ in range(5):
        t1 = time.perf_counter()
        res = pipe.run_ocr_batch(pages, geometry_procs=geo, **bench.PT_KW)
        steps.append(time.perf_counter() - t1)
    dt = time.perf_counter() - t0
    busy = (bench._proc_tree_cpu_seconds() - c0) / dt
    th1, ch1 = snap()
    n_pages = 5 * 96
    rows = sorted(((th1[k][0] - th0.get(k, (0, ""))[0]) * 1e3 / n_pages, th1[k][1]) for k in th1)
    agg = {}
    for ms, name in rows:
        key = name.rstrip("0123456789_-")
        agg[key] = agg.get(key, 0.0) + ms
    print("  main-process CPU ms per page by thread group:", {k: round(v, 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01})
    print("  child processes CPU ms per page:", round(sum(ch1[k] - ch0.get(k, 0) for k in ch1) * 1e3 / n_pages, 2), "over", len(ch1))
    ok, n_lines, right = bench._score(res, truth)
    print(f"[{tag}] median {96 / sorted(steps)[2]:.1f} pages/s, best {96 / min(steps):.1f}, cores busy {busy:.1f}, "
          f"{right}/{n_lines} right, OMP={os.environ.get('OMP_NUM_THREADS')}, geo={len(pipe._gpool or [])}", flush=True)
    pipe.close()


if __name__ == "__main__":
    main()
