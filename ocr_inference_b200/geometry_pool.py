"""Worker processes for the CPU geometry between the two networks.

cv2.findContours, approxPolyDP and the reading-order Python of BDRC/line_detection.py:107-264, 362-634 take
15-25 ms per page per core and hold the GIL between many short cv2 calls, so threads of ONE process cannot feed a
B200 (more than two geometry threads lose throughput).  `GeometryPool` keeps the GPU in one process -- one CUDA
context, one copy of the weights, pages batched through the networks -- and moves only `page_geometry(mask)` to
worker processes that never touch CUDA: the line mask travels through POSIX shared memory, the answer (a few
contours and ROI masks) comes back pickled.  Pages are independent units: no ordering between them.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import threading
import traceback
from concurrent.futures import Future
from multiprocessing import shared_memory

import numpy as np


def _worker(tasks, results):
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")       # geometry only: no CUDA context in this process
    try:
        import cv2

        from .line_detection import page_geometry

        cv2.setNumThreads(1)
        results.put(("ready", os.getpid()))
    except Exception:  # noqa: BLE001
        results.put(("init_error", traceback.format_exc()))
        return
    shms: dict[str, shared_memory.SharedMemory] = {}
    import queue

    parent = os.getppid()
    while True:
        try:
            job = tasks.get(timeout=2.0)
        except queue.Empty:
            if os.getppid() != parent:      # the owning process is gone without close() (killed, os._exit): no orphans
                break
            continue
        if job is None:
            break
        ticket, name, shape, merge_lines, census = job
        try:
            shm = shms.get(name)
            if shm is None:
                shm = shms[name] = shared_memory.SharedMemory(name=name)
            mask = np.ndarray(shape, dtype=np.uint8, buffer=shm.buf)
            results.put(("done", ticket, page_geometry(mask, merge_lines, census=census)))
        except Exception:  # noqa: BLE001
            results.put(("error", ticket, traceback.format_exc()))
    for shm in shms.values():
        shm.close()


class GeometryPool:
    """`submit(mask, merge_lines)` -> Future of line_detection.page_geometry(mask, merge_lines)."""

    def __init__(self, procs: int, slots: int | None = None, slot_bytes: int = 2048 * 3072, start_timeout: float = 300.0):
        self._ctx = mp.get_context("spawn")
        self._tasks, self._results = self._ctx.Queue(), self._ctx.Queue()
        self._procs = [self._ctx.Process(target=_worker, args=(self._tasks, self._results), daemon=True)
                       for _ in range(procs)]
        for p in self._procs:
            p.start()
        self._slot_bytes = slot_bytes
        self._free: list[shared_memory.SharedMemory] = []
        self._all: list[shared_memory.SharedMemory] = []
        self._slots_left = slots if slots is not None else 3 * procs
        self._cv = threading.Condition()
        self._pending: dict[int, tuple[Future, shared_memory.SharedMemory]] = {}
        self._ticket = 0
        self._closed = False
        self._broken = False
        import queue
        import time

        ready, deadline = 0, time.time() + start_timeout
        while ready < procs:
            try:
                msg = self._results.get(timeout=1.0)
            except queue.Empty:
                dead = [p for p in self._procs if p.exitcode not in (None, 0)]
                if dead or time.time() > deadline:
                    self.close()
                    raise RuntimeError("GeometryPool: workers did not come up"
                                       + (f" ({len(dead)} exited during start-up)" if dead else "")) from None
                continue
            if msg[0] == "init_error":
                self.close()
                raise RuntimeError("GeometryPool: a worker failed to start:\n" + msg[1])
            ready += 1
        self._reader = threading.Thread(target=self._read, name="bdrc-geometry-results", daemon=True)
        self._reader.start()

    def __len__(self):
        return len(self._procs)

    def _slot(self, nbytes: int) -> shared_memory.SharedMemory:
        with self._cv:
            while True:
                for i, s in enumerate(self._free):
                    if s.size >= nbytes:
                        return self._free.pop(i)
                if self._slots_left > 0:
                    self._slots_left -= 1
                    s = shared_memory.SharedMemory(create=True, size=max(nbytes, self._slot_bytes))
                    self._all.append(s)
                    return s
                if self._closed or self._broken:
                    raise RuntimeError("GeometryPool: closed or a worker process died")
                self._cv.wait(timeout=1.0)

    def submit(self, mask: np.ndarray, merge_lines: bool = True, census: dict | None = None) -> Future:
        """`census` (optional): line_detection.parse_census() of the GPU's connected-component census of `mask`."""
        if self._closed or self._broken:
            raise RuntimeError("GeometryPool is closed or a worker process died")
        if mask.ndim != 2 or mask.dtype != np.uint8:
            raise ValueError("GeometryPool.submit: the line mask must be a 2-D u8 array")
        slot = self._slot(mask.nbytes)
        np.copyto(np.ndarray(mask.shape, dtype=np.uint8, buffer=slot.buf), mask)
        fut: Future = Future()
        with self._cv:
            self._ticket += 1
            t = self._ticket
            self._pending[t] = (fut, slot)
        self._tasks.put((t, slot.name, tuple(mask.shape), bool(merge_lines), census))
        return fut

    def _read(self):
        import queue

        while not self._closed:
            try:
                msg = self._results.get(timeout=1.0)
            except queue.Empty:
                if any(p.exitcode not in (None, 0) for p in self._procs):
                    # fail everything in flight AND hand the slots back: callers blocked in _slot() must wake up
                    with self._cv:
                        self._broken = True
                        pend = list(self._pending.values())
                        self._pending.clear()
                        self._free.extend(slot for _, slot in pend)
                        self._cv.notify_all()
                    for fut, _ in pend:
                        fut.set_exception(RuntimeError("GeometryPool: a worker process died"))
                continue
            except (EOFError, OSError):
                return
            kind, ticket, payload = msg
            with self._cv:
                fut, slot = self._pending.pop(ticket, (None, None))
                if slot is not None:
                    self._free.append(slot)
                    self._cv.notify()
            if fut is None:
                continue
            if kind == "done":
                fut.set_result(payload)
            else:
                fut.set_exception(RuntimeError("page_geometry failed in a worker:\n" + payload))

    def close(self):
        if self._closed:
            return
        self._closed = True
        for _ in self._procs:
            try:
                self._tasks.put(None)
            except Exception:  # noqa: BLE001
                pass
        for p in self._procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()
        for s in self._all:
            try:
                s.close()
                s.unlink()
            except Exception:  # noqa: BLE001
                pass
        self._all.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
