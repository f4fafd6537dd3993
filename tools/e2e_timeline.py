"""Where the end-to-end C2 step (LineDetection.predict_batch, pageable numpy in -> fresh numpy out) spends its
time on a B200 box: host copy rate into pinned memory by thread count, H2D / D2H rates, the device-resident step,
and the e2e step for several group plans / upload thread counts.   python tools/e2e_timeline.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ocr_inference_b200 import synthetic  # noqa: E402
from ocr_inference_b200.data import LineDetectionConfig  # noqa: E402
from ocr_inference_b200.inference import LineDetection, get_engine  # noqa: E402
from ocr_inference_b200.synthetic_weights import seg_weights  # noqa: E402


def med(f, reps=7):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


def main():
    eng = get_engine(0)
    det = LineDetection(LineDetectionConfig("", seg_weights(11, 1, 3), "deeplabv3plus-r34", 512, ["line"]), engine=eng)
    pages = [synthetic.page(100 + i, 2048, 3072)[0] for i in range(16)]
    nb = pages[0].nbytes
    print(f"cores {os.cpu_count()}, page {nb / 1e6:.1f} MB")
    pin = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    dev = torch.empty(nb, dtype=torch.uint8, device="cuda")
    ptr = np.array([pages[0].ctypes.data], dtype=np.uint64)
    nbs = np.array([nb], dtype=np.int64)
    zero = np.zeros(1, dtype=np.int64)
    for T in (1, 2, 4, 6, 8):
        ms = med(lambda: eng.lib.bdrc_host_gather(ptr.ctypes.data, nbs.ctypes.data, zero.ctypes.data, 1, pin.data_ptr(), T))
        print(f"host copy pageable -> pinned, {T} threads: {ms:.2f} ms = {nb / ms / 1e6:.1f} GB/s")

    def h2d():
        dev.copy_(pin, non_blocking=True)
        torch.cuda.synchronize()

    ms = med(h2d)
    print(f"H2D pinned -> device: {ms:.2f} ms = {nb / ms / 1e6:.1f} GB/s")

    def pageable():
        torch.from_numpy(pages[1]).to("cuda")
        torch.cuda.synchronize()

    ms = med(pageable)
    print(f"pageable .to(cuda): {ms:.2f} ms = {nb / ms / 1e6:.1f} GB/s")
    m = np.empty((2048, 3072), dtype=np.uint8)

    def fresh():
        f = np.empty((2048, 3072), dtype=np.uint8)
        np.copyto(f, m)

    print(f"fresh 6.3 MB numpy array + copy (1 thread): {med(fresh):.2f} ms")
    # device-resident groups
    pd = [torch.from_numpy(p).cuda() for p in pages]
    for g in (1, 2, 3, 6, 8, 16):
        def run():
            for g0 in range(0, 16, g):
                det._predict_device_batch(pd[g0:g0 + g], 0.9)
            torch.cuda.synchronize()
        run()
        print(f"device-resident 16 pages in groups of {g}: {med(run, 5):.2f} ms")
    for plan in ([2], [2, 6, 6, 2], [2, 6, 8], [3, 6, 7], [2, 5, 9], [2, 4, 10], [3, 13], [2, 4, 4, 6], [2, 6, 7, 1]):
        for T in (4, 8):
            eng.upload_threads = T
            det.predict_batch(pages, group_sizes=plan)
            ms = med(lambda: det.predict_batch(pages, group_sizes=plan), 7)
            print(f"e2e predict_batch(16 pageable pages) plan {plan} upload threads {T}: {ms:.2f} ms = {16e3 / ms:.1f} pages/s", flush=True)
    pinned = [torch.from_numpy(p).pin_memory() for p in pages]
    ms = med(lambda: det.predict_batch(pages, pinned=pinned, copy=False), 7)
    print(f"io-binding style: {ms:.2f} ms = {16e3 / ms:.1f} pages/s")


if __name__ == "__main__":
    main()
