"""Several host processes per GPU for page -> text throughput.

The two networks take a few milliseconds per page on a B200; the cv2 contour geometry and reading-order
Python between them (BDRC/line_detection.py:107-264, 362-634) take 12-25 ms per page per
core and hold the GIL, so one process cannot feed one GPU (75 pages/s against 240 with eight processes).
Pages are independent units: `PipelinePool` starts `procs_per_device` workers per GPU, each
with its own `OCRPipeline` (CUDA context, weights, workspaces), hands pages to them through POSIX shared
memory and returns exactly what `OCRPipeline.run_ocr` returns, in input order.  No collective, no GPU-side
exchange: a host-side gather, like `sharding.run_sharded`.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import queue
import traceback
from multiprocessing import shared_memory
from typing import Callable, Sequence

import numpy as np


def _default_factory(ocr_config, line_config, use_line_prepadding, device):
    import torch

    from .inference import OCRPipeline, get_engine

    torch.cuda.set_device(device)
    return OCRPipeline(ocr_config, line_config, None, use_line_prepadding=use_line_prepadding, engine=get_engine(device))


class _EchoPipeline:
    """Stand-in pipeline for testing the pool's plumbing without a GPU (tests/test_cpu_host_logic.py)."""

    def __init__(self, device):
        self.device = device

    def run_ocr(self, image, **kw):
        from .data import OpStatus

        if image.shape[0] == 1:
            return OpStatus.FAILED, "No lines detected"
        return OpStatus.SUCCESS, [image[..., 0].copy(), [], [f"sum={int(image.sum())} k={kw.get('k_factor')}"], float(self.device)]


def _echo_factory(ocr_config, line_config, use_line_prepadding, device):
    return _EchoPipeline(device)


def _worker(device, factory, factory_args, tasks, results, keep_mask):
    try:
        pipe = factory(*factory_args, device)
    except Exception:  # noqa: BLE001
        results.put(("init_error", traceback.format_exc()))
        return
    results.put(("ready", os.getpid()))
    while True:
        job = tasks.get()
        if job is None:
            return
        call, idx, name, shape, kw = job
        shm = None
        try:
            shm = shared_memory.SharedMemory(name=name)
            page = np.ndarray(shape, dtype=np.uint8, buffer=shm.buf)
            status, payload = pipe.run_ocr(page, **kw)
            if isinstance(payload, list) and not keep_mask:
                payload = [None] + list(payload[1:])     # rot_mask is page-sized: only sent back on request
            elif isinstance(payload, list):
                payload = [np.array(payload[0])] + list(payload[1:])
            results.put(("done", call, idx, status, payload))
        except Exception:  # noqa: BLE001
            results.put(("error", call, idx, traceback.format_exc()))
        finally:
            if shm is not None:
                shm.close()


class PipelinePool:
    def __init__(self, ocr_config=None, line_config=None, devices: Sequence[int] = (0,), procs_per_device: int = 4,
                 use_line_prepadding: bool = False, return_masks: bool = False,
                 factory: Callable | None = None, start_timeout: float = 300.0):
        """`factory(ocr_config, line_config, use_line_prepadding, device)` builds the per-process pipeline
        (default: OCRPipeline on that GPU); anything with a `run_ocr(image, **kw)` method works."""
        self._ctx = mp.get_context("spawn")
        self._tasks = self._ctx.Queue()
        self._results = self._ctx.Queue()
        self._procs = []
        args = (ocr_config, line_config, use_line_prepadding)
        for dev in devices:
            for _ in range(procs_per_device):
                p = self._ctx.Process(target=_worker, args=(dev, factory or _default_factory, args, self._tasks,
                                                            self._results, return_masks), daemon=True)
                p.start()
                self._procs.append(p)
        ready = 0
        while ready < len(self._procs):
            try:
                msg = self._results.get(timeout=start_timeout)
            except queue.Empty:
                self.close()
                raise RuntimeError("PipelinePool: workers did not come up") from None
            if msg[0] == "init_error":
                self.close()
                raise RuntimeError("PipelinePool: a worker failed to build its pipeline:\n" + msg[1])
            ready += 1

    def __len__(self):
        return len(self._procs)

    def run_ocr_many(self, images: Sequence[np.ndarray], **run_ocr_kwargs):
        """[(OpStatus, payload)] like [OCRPipeline.run_ocr(i, **kw) for i in images] (payload[0], the rotated
        mask, is None unless the pool was built with return_masks=True)."""
        from .data import OpStatus

        # every call tags its jobs: results of an earlier call that was abandoned half-way (a dead worker)
        # can never be mistaken for this call's pages
        self._call = getattr(self, "_call", 0) + 1
        call = self._call
        shms = []
        try:
            for idx, img in enumerate(images):
                img = np.ascontiguousarray(img, dtype=np.uint8)
                shm = shared_memory.SharedMemory(create=True, size=max(1, img.nbytes))
                np.ndarray(img.shape, dtype=np.uint8, buffer=shm.buf)[...] = img
                shms.append(shm)
                self._tasks.put((call, idx, shm.name, img.shape, run_ocr_kwargs))
            out: list = [None] * len(images)
            got = 0
            while got < len(images):
                try:
                    msg = self._results.get(timeout=5)
                except queue.Empty:
                    if any(p.exitcode not in (None, 0) for p in self._procs):
                        # pages still in flight are lost with the worker; report them, keep the finished ones
                        for i in range(len(images)):
                            if out[i] is None:
                                out[i] = (OpStatus.FAILED, "OCR pipeline failed: a PipelinePool worker process died")
                        self._broken = True
                        return out
                    continue
                if msg[1] != call:
                    continue                      # stale result of an abandoned call
                if msg[0] == "error":
                    _, _, idx, tb = msg
                    out[idx] = (OpStatus.FAILED, f"OCR pipeline failed: {tb.strip().splitlines()[-1]}")
                else:
                    _, _, idx, status, payload = msg
                    out[idx] = (status, payload)
                got += 1
            return out
        finally:
            if getattr(self, "_broken", False):
                self.close()                      # workers may still read the shared pages: stop them first
            for shm in shms:
                shm.close()
                shm.unlink()

    def close(self):
        for _ in self._procs:
            try:
                self._tasks.put(None)
            except Exception:  # noqa: BLE001
                pass
        for p in self._procs:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()
        self._procs = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
