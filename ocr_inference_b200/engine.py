"""Thin device-side engine over the C ABI: torch owns memory and streams, libbdrc_b200.so does the work.

Every method takes / returns torch CUDA tensors.  No method has a CPU fallback; constructing an
Engine without a B200 raises (ocr_inference_b200._lib.Context).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading

import numpy as np
import torch

from . import _lib
from .weights import write_blob

SEG, OCR = 1, 2
POST_SIGMOID, POST_SOFTMAX = 0, 1
LAYOUT_TV, LAYOUT_VT = 0, 1
# events that host threads wait on are created with cudaEventBlockingSync: the waiting thread sleeps instead of
# spinning on a core (the page -> text path keeps ~10 such threads per GPU); BDRC_SPIN_EVENTS=1 restores spinning
BLOCKING_EVENTS = os.environ.get("BDRC_SPIN_EVENTS", "0") != "1"
OCR_CPAD = 128
PATCH = 512


def clamp_size(H: int, W: int, patch: int = PATCH, clamp_w: int = 4096, clamp_h: int = 2048):
    """Target size of the clamp / upsize branch of utils.preprocess_image (BDRC/utils.py:764-773);
    resize_to_width/height truncate the free side with int() (:473, :491)."""
    if W > H and W > clamp_w:
        return int(H * (clamp_w / W)), clamp_w
    if H > W and H > clamp_h:
        return clamp_h, int(W * (clamp_h / H))
    if H < patch:
        return patch, int(W * (patch / H))
    return H, W


def clamp_size_torch(H: int, W: int, patch: int = PATCH, max_w: int = 4096, max_h: int = 2048):
    """utils.resize_clamp (BDRC/utils.py:967-987): same branches, but round() instead of int()."""
    if W > H and W > max_w:
        sc = max_w / W
    elif H > W and H > max_h:
        sc = max_h / H
    elif H < patch:
        sc = patch / H
    else:
        return H, W
    return int(round(H * sc)), int(round(W * sc))


class Model:
    def __init__(self, eng: "Engine", kind: int, handle, n_out: int):
        self.eng, self.kind, self.h, self.n_out = eng, kind, handle, n_out

    def close(self):
        if self.h is not None and self.h.value:
            self.eng.lib.bdrc_free_model(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class Engine:
    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise _lib.BdrcError("no CUDA device: ocr_inference_b200 has no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.ctx = _lib.Context(device)
        self.lib = self.ctx.lib
        # workspaces and pinned staging buffers are per host thread: run_ocr_batch drives one page per
        # worker thread, each on its own CUDA stream
        self._ws: dict[tuple, torch.Tensor] = {}
        self._pin: dict[tuple, list] = {}
        # bdrc_host_gather runs on persistent native threads; tools/pack_bench.py on the B200 host, 19 MB of crops:
        # numpy 1.53 ms, 2 threads 1.21, 4 threads 1.29, 8 threads 1.80 (host copy bandwidth, not thread count)
        self._gather_threads = max(1, min(2, len(os.sched_getaffinity(0)) // 2))

    # ---- plumbing ------------------------------------------------------------------------
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc, what):
        self.ctx.check(rc, what)

    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        k = (key, threading.get_ident())
        t = self._ws.get(k)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
            self._ws[k] = t
        return t

    def _pinned(self, key: str, nbytes: int) -> torch.Tensor:
        """Grow-only pinned staging buffer of this thread (cudaHostAlloc per call costs milliseconds).
        The previous asynchronous copy out of it is waited for before it is handed out again."""
        k = (key, threading.get_ident())
        ent = self._pin.get(k)
        if ent is not None and ent[1] is not None:
            ent[1].synchronize()
        if ent is None or ent[0].numel() < nbytes:
            ent = [torch.empty(max(int(nbytes * 1.25), 1 << 16), dtype=torch.uint8, pin_memory=True), None]
            self._pin[k] = ent
        return ent[0][:max(nbytes, 1)]

    def _upload(self, key: str, host: torch.Tensor) -> torch.Tensor:
        """Asynchronous H2D of a buffer obtained from _pinned(key, ...)."""
        dev = host.to(self.device, non_blocking=True)
        ev = torch.cuda.Event(blocking=BLOCKING_EVENTS)          # host threads wait on it (_pinned): sleep, do not spin
        ev.record(torch.cuda.current_stream(self.device))
        self._pin[(key, threading.get_ident())][1] = ev
        return dev

    def upload_page(self, img: np.ndarray, ring: list, k: int, threads: int | None = None):
        """Asynchronous H2D of a pageable u8 array on the current stream through slot k of a ring of pinned
        buffers (`ring` = [[tensor | None, event | None], ...], owned by the caller): a few native threads copy
        the array into the slot (bdrc_host_gather, GIL released), the DMA out of it does not block the host, so
        the copy of the next page overlaps it.  A plain pageable .to() moves ~7 GB/s from one thread (2.7 ms per
        19 MB page; the native copy: 16 / 32 / 65 GB/s with 1 / 4 / 8 threads on a 16-core box).  Returns (device tensor,
        event recorded after the copy)."""
        img = np.ascontiguousarray(img)
        nbytes = int(img.nbytes)
        slot = ring[k % len(ring)]
        if slot[1] is not None:
            slot[1].synchronize()
        if slot[0] is None or slot[0].numel() < nbytes:
            slot[0] = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
        if nbytes >= (1 << 20):
            ptr = np.array([img.ctypes.data], dtype=np.uint64)
            nb = np.array([nbytes], dtype=np.int64)
            zero = np.zeros(1, dtype=np.int64)
            self._check(self.lib.bdrc_host_gather(ptr.ctypes.data, nb.ctypes.data, zero.ctypes.data, 1,
                                                  slot[0].data_ptr(), int(threads or getattr(self, "upload_threads", 4))),
                        "host_gather")
        elif nbytes:
            np.copyto(slot[0][:nbytes].numpy().reshape(img.shape), img)
        d = torch.empty(img.shape, dtype=torch.uint8, device=self.device)
        if nbytes:
            d.copy_(slot[0][:nbytes].view(img.shape), non_blocking=True)
        ev = torch.cuda.Event(blocking=BLOCKING_EVENTS)          # the upload thread waits on it before it refills the slot
        ev.record(torch.cuda.current_stream(self.device))
        slot[1] = ev
        return d, ev

    def host_copy_many(self, src_ptrs: list, dst_ptrs: list, nbytes: list, threads: int = 4):
        """Host memory -> host memory, n ranges, by a few native threads with the GIL released (bdrc_host_copy)."""
        n = len(src_ptrs)
        if n == 0:
            return
        sp = np.array(src_ptrs, dtype=np.uint64)
        dp = np.array(dst_ptrs, dtype=np.uint64)
        nb = np.array(nbytes, dtype=np.int64)
        self._check(self.lib.bdrc_host_copy(sp.ctypes.data, dp.ctypes.data, nb.ctypes.data, n, threads), "host_copy")

    def set_dec_tail(self, fused: bool = True):
        """Decoder tail of the segmentation network as one fused kernel (default) or three (cross-check)."""
        self._check(self.lib.bdrc_set_dec_tail(self.ctx.h, int(bool(fused))), "set_dec_tail")

    def set_wide_conv(self, on: bool = True):
        """64-channel 3x3 layers on the wide-N kernel (default) or the slab kernel (cross-check)."""
        self._check(self.lib.bdrc_set_wide_conv(self.ctx.h, int(bool(on))), "set_wide_conv")

    def set_blocking_sync(self, on: bool = True):
        """Host threads waiting for this GPU sleep instead of spinning (see bdrc_set_blocking_sync)."""
        self._check(self.lib.bdrc_set_blocking_sync(self.ctx.h, int(bool(on))), "set_blocking_sync")
        self.blocking_sync = bool(on)

    def simt_launch_count(self) -> int:
        return self.ctx.simt_launch_count()

    def conv_plans_built(self) -> int:
        return self.ctx.conv_plans_built()

    def launch_count(self) -> int:
        return self.ctx.launch_count()

    def set_conv_engine(self, engine: int):
        self._check(self.lib.bdrc_set_conv_engine(self.ctx.h, engine), "set_conv_engine")

    def profile(self, on: bool):
        self._check(self.lib.bdrc_profile_enable(self.ctx.h, int(on)), "profile_enable")
        self._check(self.lib.bdrc_profile_reset(self.ctx.h), "profile_reset")

    def profile_results(self):
        out = []
        for i in range(self.lib.bdrc_profile_count(self.ctx.h)):
            name, ms, fl = C.c_char_p(), C.c_float(), C.c_double()
            self._check(self.lib.bdrc_profile_get(self.ctx.h, i, C.byref(name), C.byref(ms), C.byref(fl)),
                        "profile_get")
            out.append((name.value.decode(), ms.value, fl.value))
        self._check(self.lib.bdrc_profile_reset(self.ctx.h), "profile_reset")
        return out

    # ---- models --------------------------------------------------------------------------
    def load_model(self, kind: int, weights: dict[str, np.ndarray] | bytes, n_out: int) -> Model:
        blob = weights if isinstance(weights, (bytes, bytearray)) else write_blob(weights)
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        h = C.c_void_p()
        self._check(self.lib.bdrc_load_weights(self.ctx.h, kind, buf, len(blob), n_out, C.byref(h)),
                    "load_weights")
        return Model(self, kind, h, n_out)

    # ---- a-1 -----------------------------------------------------------------------------
    def resize_u8(self, img: torch.Tensor, Hd: int, Wd: int) -> torch.Tensor:
        """cv2.resize(INTER_LINEAR) on an HWC (or HW) u8 tensor."""
        img = img.contiguous()
        Hs, Ws = img.shape[:2]
        Cn = img.shape[2] if img.dim() == 3 else 1
        out = torch.empty((Hd, Wd) + ((Cn,) if img.dim() == 3 else ()), dtype=torch.uint8, device=self.device)
        self._check(self.lib.bdrc_resize_bilinear_u8(self.ctx.h, img.data_ptr(), Hs, Ws, Cn, out.data_ptr(),
                                                     Hd, Wd, self.stream), "resize_bilinear_u8")
        return out

    def seg_preprocess(self, page: torch.Tensor, patch: int = PATCH, out: torch.Tensor | None = None):
        """page u8 HWC on device -> (tiles u8 {0,1} [N,P,P], meta) as Detection._preprocess_image.
        `out` (optional) is a preallocated [N,P,P] u8 slice of a larger tile batch."""
        H0, W0 = int(page.shape[0]), int(page.shape[1])
        Hc, Wc = clamp_size(H0, W0, patch)
        page = page.contiguous()
        if (Hc, Wc) != (H0, W0):
            page = self.resize_u8(page, Hc, Wc)
        nty, ntx = math.ceil(Hc / patch), math.ceil(Wc / patch)
        tiles = out if out is not None else torch.empty((nty * ntx, patch, patch), dtype=torch.uint8,
                                                        device=self.device)
        assert tiles.shape[0] == nty * ntx and tiles.is_contiguous()
        self._check(self.lib.bdrc_seg_preprocess(self.ctx.h, page.data_ptr(), Hc, Wc, patch, tiles.data_ptr(),
                                                 self.stream), "seg_preprocess")
        meta = dict(H0=H0, W0=W0, Hc=Hc, Wc=Wc, nty=nty, ntx=ntx, pad_x=ntx * patch - Wc, pad_y=nty * patch - Hc)
        return tiles, meta

    def seg_preprocess_batch(self, pages: list, patch: int = PATCH):
        """a-1 for several device-resident pages in ONE launch -> (tiles u8 {0,1} [sum N,P,P], metas).  Pages that
        need the clamp resize are resized first (one small kernel each, rare)."""
        metas, srcs = [], []
        for page in pages:
            H0, W0 = int(page.shape[0]), int(page.shape[1])
            Hc, Wc = clamp_size(H0, W0, patch)
            page = page.contiguous()
            if (Hc, Wc) != (H0, W0):
                page = self.resize_u8(page, Hc, Wc)
            nty, ntx = math.ceil(Hc / patch), math.ceil(Wc / patch)
            srcs.append(page)
            metas.append(dict(H0=H0, W0=W0, Hc=Hc, Wc=Wc, nty=nty, ntx=ntx, pad_x=ntx * patch - Wc,
                              pad_y=nty * patch - Hc))
        n = len(srcs)
        tiles = torch.empty((sum(m["nty"] * m["ntx"] for m in metas), patch, patch), dtype=torch.uint8, device=self.device)
        if n == 0:
            return tiles, metas
        ptrs = np.array([t.data_ptr() for t in srcs], dtype=np.uint64)
        hs = np.array([m["Hc"] for m in metas], dtype=np.int32)
        ws = np.array([m["Wc"] for m in metas], dtype=np.int32)
        self._check(self.lib.bdrc_seg_preprocess_batch(self.ctx.h, ptrs.ctypes.data, hs.ctypes.data, ws.ctypes.data, n,
                                                       patch, tiles.data_ptr(), self.stream), "seg_preprocess_batch")
        return tiles, metas

    def seg_postprocess_lowres_batch(self, logits: torch.Tensor, metas: list, mode: int, thr: float) -> list:
        """a-3 for the pages of one forward call in ONE launch: low-res logits [sum N,128,128,C] -> one mask per page
        (slices of one allocation)."""
        Cn = int(logits.shape[3])
        logits = logits.contiguous()
        for m in metas:
            if (m["Hc"], m["Wc"]) != (m["H0"], m["W0"]):
                raise _lib.BdrcError("low-res post-processing does not resize; use the full-res path")
        sizes = [m["H0"] * m["W0"] * Cn for m in metas]
        flat = torch.empty((sum((s + 15) // 16 * 16 for s in sizes),), dtype=torch.uint8, device=self.device)
        masks, lptr, mptr, o, t0 = [], [], [], 0, 0
        per_tile = 128 * 128 * Cn * 4
        for m, sz in zip(metas, sizes):
            shape = (m["H0"], m["W0"]) if Cn == 1 else (m["H0"], m["W0"], Cn)
            masks.append(flat[o:o + sz].view(shape))
            mptr.append(flat.data_ptr() + o)
            lptr.append(logits.data_ptr() + t0 * per_tile)
            o += (sz + 15) // 16 * 16
            t0 += m["nty"] * m["ntx"]
        n = len(metas)
        if n == 0:
            return masks
        arr = lambda v, dt: np.array(v, dtype=dt)  # noqa: E731
        lp, mp = arr(lptr, np.uint64), arr(mptr, np.uint64)
        nty, ntx = arr([m["nty"] for m in metas], np.int32), arr([m["ntx"] for m in metas], np.int32)
        hc, wc = arr([m["Hc"] for m in metas], np.int32), arr([m["Wc"] for m in metas], np.int32)
        self._check(self.lib.bdrc_seg_postprocess_lowres_batch(
            self.ctx.h, lp.ctypes.data, nty.ctypes.data, ntx.ctypes.data, hc.ctypes.data, wc.ctypes.data, n, Cn, PATCH,
            mode, thr, mp.ctypes.data, self.stream), "seg_postprocess_lowres_batch")
        return masks

    def seg_tile_rgb(self, page: torch.Tensor, patch: int = PATCH, out: torch.Tensor | None = None):
        """torch-path collate (utils.multi_image_collate_fn): fp32 RGB tiles [N,P,P,3] in [0,255].
        `out` (optional) is a preallocated [N,P,P,3] fp32 slice of a larger tile batch (several images per
        forward call, as the reference's collate function batches them)."""
        H0, W0 = int(page.shape[0]), int(page.shape[1])
        Hc, Wc = clamp_size_torch(H0, W0, patch)
        page = page.contiguous()
        nty, ntx = math.ceil(Hc / patch), math.ceil(Wc / patch)
        tiles = out if out is not None else torch.empty((nty * ntx, patch, patch, 3), dtype=torch.float32,
                                                        device=self.device)
        assert tuple(tiles.shape) == (nty * ntx, patch, patch, 3) and tiles.is_contiguous()
        self._check(self.lib.bdrc_seg_tile_rgb(self.ctx.h, page.data_ptr(), H0, W0, Hc, Wc, patch,
                                               tiles.data_ptr(), self.stream), "seg_tile_rgb")
        meta = dict(H0=H0, W0=W0, Hc=Hc, Wc=Wc, nty=nty, ntx=ntx, pad_x=ntx * patch - Wc, pad_y=nty * patch - Hc)
        return tiles, meta

    # ---- a-2 -----------------------------------------------------------------------------
    def seg_forward(self, model: Model, tiles: torch.Tensor, lowres: bool = False) -> torch.Tensor:
        n = int(tiles.shape[0])
        cin = 1 if tiles.dtype == torch.uint8 else 3
        Cn = model.n_out
        ws = self._workspace("seg", int(self.lib.bdrc_seg_workspace_bytes(model.h, n)))
        tiles = tiles.contiguous()
        if lowres:
            out = torch.empty((n, 128, 128, Cn), dtype=torch.float32, device=self.device)
            fn = self.lib.bdrc_seg_forward_lowres
        else:
            out = torch.empty((n, Cn, PATCH, PATCH), dtype=torch.float32, device=self.device)
            fn = self.lib.bdrc_seg_forward
        self._check(fn(self.ctx.h, model.h, tiles.data_ptr(), n, cin, out.data_ptr(), ws.data_ptr(),
                       ws.numel(), self.stream), "seg_forward")
        return out

    # ---- a-3 -----------------------------------------------------------------------------
    def seg_postprocess(self, logits: torch.Tensor, meta: dict, mode: int, thr: float,
                        lowres: bool = False, out: torch.Tensor | None = None) -> torch.Tensor:
        Cn = int(logits.shape[3] if lowres else logits.shape[1])
        H0, W0 = meta["H0"], meta["W0"]
        if lowres and (meta["Hc"], meta["Wc"]) != (H0, W0):
            raise _lib.BdrcError("low-res post-processing does not resize; use the full-res path")
        shape = (H0, W0) if Cn == 1 else (H0, W0, Cn)
        mask = out if out is not None else torch.empty(shape, dtype=torch.uint8, device=self.device)
        logits = logits.contiguous()
        if lowres:
            rc = self.lib.bdrc_seg_postprocess_lowres(self.ctx.h, logits.data_ptr(), meta["nty"], meta["ntx"], Cn,
                                                      PATCH, meta["Hc"], meta["Wc"], mode, thr, mask.data_ptr(),
                                                      self.stream)
        else:
            rc = self.lib.bdrc_seg_postprocess(self.ctx.h, logits.data_ptr(), meta["nty"], meta["ntx"], Cn, PATCH,
                                               meta["Hc"], meta["Wc"], H0, W0, mode, thr, mask.data_ptr(),
                                               self.stream)
        self._check(rc, "seg_postprocess")
        return mask

    def seg_postprocess_torch(self, logits: torch.Tensor, meta: dict, thr: float) -> torch.Tensor:
        Cn = int(logits.shape[1])
        mask = torch.empty((Cn, meta["H0"], meta["W0"]), dtype=torch.uint8, device=self.device)
        logits = logits.contiguous()
        self._check(self.lib.bdrc_seg_postprocess_torch(self.ctx.h, logits.data_ptr(), meta["nty"], meta["ntx"], Cn,
                                                        PATCH, meta["Hc"], meta["Wc"], meta["H0"], meta["W0"], thr,
                                                        mask.data_ptr(), self.stream), "seg_postprocess_torch")
        return mask

    # ---- a-4 / a-6 -----------------------------------------------------------------------
    def warp_affine(self, img: torch.Tensor, M: np.ndarray) -> torch.Tensor:
        """cv2.warpAffine(img, M, (W, H), INTER_LINEAR, borderValue=0) on an HWC / HW u8 tensor."""
        img = img.contiguous()
        H, W = int(img.shape[0]), int(img.shape[1])
        Cn = int(img.shape[2]) if img.dim() == 3 else 1
        out = torch.empty_like(img)
        Mh = np.ascontiguousarray(M, dtype=np.float64).reshape(6)
        self._check(self.lib.bdrc_warp_affine_u8(self.ctx.h, img.data_ptr(), H, W, Cn, Mh.ctypes.data, out.data_ptr(),
                                                 self.stream), "warp_affine_u8")
        return out

    def remap_bilinear(self, img: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
        """map_coordinates(order=1, mode='constant', cval=0); coords fp32 [2,H,W] = (row, col)."""
        img, coords = img.contiguous(), coords.contiguous().float()
        H, W = int(img.shape[0]), int(img.shape[1])
        Cn = int(img.shape[2]) if img.dim() == 3 else 1
        out = torch.empty_like(img)
        self._check(self.lib.bdrc_remap_bilinear_u8(self.ctx.h, img.data_ptr(), H, W, Cn, coords.data_ptr(),
                                                    out.data_ptr(), self.stream), "remap_bilinear_u8")
        return out

    def tps_remap(self, img: torch.Tensor, ctrl: np.ndarray, params: np.ndarray) -> torch.Tensor:
        """Thin-plate-spline dewarp: output(row, col) = bilinear(img, tps(row, col))."""
        img = img.contiguous()
        H, W = int(img.shape[0]), int(img.shape[1])
        Cn = int(img.shape[2]) if img.dim() == 3 else 1
        out = torch.empty_like(img)
        c = np.ascontiguousarray(ctrl, dtype=np.float64)
        p = np.ascontiguousarray(params, dtype=np.float64)
        ws = self._workspace("tps", int(self.lib.bdrc_tps_workspace_bytes()))
        self._check(self.lib.bdrc_tps_remap_u8(self.ctx.h, img.data_ptr(), H, W, Cn, c.ctypes.data, p.ctypes.data,
                                               int(c.shape[0]), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                               self.stream), "tps_remap_u8")
        return out

    def spline3_prefilter(self, img: torch.Tensor) -> torch.Tensor:
        """u8 HWC image -> fp64 cubic B-spline coefficients [H, W, C] (scipy.ndimage.spline_filter, order 3)."""
        img = img.contiguous()
        H, W = int(img.shape[0]), int(img.shape[1])
        Cn = int(img.shape[2]) if img.dim() == 3 else 1
        coef = torch.empty((H, W, Cn), dtype=torch.float64, device=self.device)
        assert coef.numel() * 8 == int(self.lib.bdrc_spline3_coef_bytes(H, W, Cn))
        ws = self._workspace("spline3", int(self.lib.bdrc_spline3_workspace_bytes(H, W, Cn)))
        self._check(self.lib.bdrc_spline3_prefilter_u8(self.ctx.h, img.data_ptr(), H, W, Cn, coef.data_ptr(),
                                                       ws.data_ptr(), ws.numel(), self.stream), "spline3_prefilter_u8")
        return coef

    def remap_cubic(self, img: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
        """scipy.ndimage.map_coordinates(order=3, mode='constant', cval=0) per channel; coords fp64 [2, H, W]."""
        coef = self.spline3_prefilter(img)
        H, W, Cn = (int(v) for v in coef.shape)
        coords = coords.to(device=self.device, dtype=torch.float64).contiguous()
        if tuple(coords.shape) != (2, H, W):
            raise ValueError("remap_cubic: coords must be [2, H, W]")
        out = torch.empty(img.shape, dtype=torch.uint8, device=self.device)
        self._check(self.lib.bdrc_remap_cubic_u8(self.ctx.h, coef.data_ptr(), H, W, Cn, coords.data_ptr(), out.data_ptr(),
                                                 self.stream), "remap_cubic_u8")
        return out

    def tps_remap_cubic(self, img: torch.Tensor, ctrl: np.ndarray, params: np.ndarray) -> torch.Tensor:
        """Thin-plate-spline dewarp with cubic B-spline sampling (the reference's intended run_tps)."""
        coef = self.spline3_prefilter(img)
        H, W, Cn = (int(v) for v in coef.shape)
        out = torch.empty(img.shape, dtype=torch.uint8, device=self.device)
        c = np.ascontiguousarray(ctrl, dtype=np.float64)
        p = np.ascontiguousarray(params, dtype=np.float64)
        ws = self._workspace("tps", int(self.lib.bdrc_tps_workspace_bytes()))
        self._check(self.lib.bdrc_tps_remap_cubic_u8(self.ctx.h, coef.data_ptr(), H, W, Cn, c.ctypes.data, p.ctypes.data,
                                                     int(c.shape[0]), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                     self.stream), "tps_remap_cubic_u8")
        return out

    # ---- a-5 -----------------------------------------------------------------------------
    def pack_masks(self, masks: list[np.ndarray], rois: list[tuple[int, int, int, int]]):
        """ROI masks (filled contours) -> one device buffer + host geometry arrays."""
        roi = np.array(rois, dtype=np.int32).reshape(-1, 4)
        sizes = [int(m.shape[0]) * int(m.shape[1]) for m in masks]
        off = np.zeros(len(masks), dtype=np.int64)
        if len(masks) > 1:
            off[1:] = np.cumsum([(s + 15) // 16 * 16 for s in sizes[:-1]])
        total = int(off[-1] + sizes[-1]) if masks else 0
        host = self._pinned("masks", total)
        hv = host.numpy()
        for m, o, s in zip(masks, off, sizes):
            hv[o:o + s] = np.ascontiguousarray(m, dtype=np.uint8).reshape(-1)
        dev = self._upload("masks", host)
        return dict(dev=dev, roi=roi, off=off)

    def _sub_pack(self, pack, idx):
        roi = np.ascontiguousarray(pack["roi"][idx])
        off = np.ascontiguousarray(pack["off"][idx])
        sat = int(((roi[:, 2].astype(np.int64) + 1) * (roi[:, 3].astype(np.int64) + 1)).sum())
        return roi, off, sat

    def line_crop_heights(self, page: torch.Tensor, pack, idx: list[int], cand: list[list[tuple[int, int]]]):
        """Crop heights for lines pack[idx] x candidate SEs; returns host int32 [len(idx), n_cand]."""
        page = page.contiguous()
        H, W = int(page.shape[0]), int(page.shape[1])
        Cn = int(page.shape[2]) if page.dim() == 3 else 1
        roi, off, sat = self._sub_pack(pack, idx)
        L, NC = len(idx), len(cand[0])
        kw = np.array([[c[0] for c in row] for row in cand], dtype=np.int32)
        kh = np.array([[c[1] for c in row] for row in cand], dtype=np.int32)
        heights = torch.empty((L, NC), dtype=torch.int32, device=self.device)
        ws = self._workspace("extract", int(self.lib.bdrc_extract_lines_workspace_bytes(H, W, L * NC, sat)))
        self._check(self.lib.bdrc_line_crop_heights(
            self.ctx.h, page.data_ptr(), H, W, Cn, pack["dev"].data_ptr(), roi.ctypes.data, off.ctypes.data, L, NC,
            kw.ctypes.data, kh.ctypes.data, heights.data_ptr(), ws.data_ptr(), ws.numel(), self.stream),
            "line_crop_heights")
        return heights.cpu().numpy()

    def extract_lines(self, page: torch.Tensor, pack, idx: list[int], ses: list[tuple[int, int]]):
        """Masked ragged crops for lines pack[idx] with SE ses[i]; returns (crops u8, hw [L,2], off [L])."""
        page = page.contiguous()
        H, W = int(page.shape[0]), int(page.shape[1])
        Cn = int(page.shape[2]) if page.dim() == 3 else 1
        roi, off, sat = self._sub_pack(pack, idx)
        L = len(idx)
        kw = np.array([s[0] for s in ses], dtype=np.int32)
        kh = np.array([s[1] for s in ses], dtype=np.int32)
        cap = 0
        for (x, y, w, h), a, b in zip(roi.tolist(), kw.tolist(), kh.tolist()):
            bw = min(W - 1, x + w - 1 + a // 2) - max(0, x - (a - 1 - a // 2)) + 1
            bh = min(H - 1, y + h - 1 + b // 2) - max(0, y - (b - 1 - b // 2)) + 1
            cap += (max(bw, 0) * max(bh, 0) * Cn + 15) // 16 * 16
        crops = torch.empty(max(cap, 16), dtype=torch.uint8, device=self.device)
        hw = torch.empty((L, 2), dtype=torch.int32, device=self.device)
        coff = torch.empty((L,), dtype=torch.int64, device=self.device)
        ws = self._workspace("extract", int(self.lib.bdrc_extract_lines_workspace_bytes(H, W, L, sat)))
        self._check(self.lib.bdrc_extract_lines(
            self.ctx.h, page.data_ptr(), H, W, Cn, pack["dev"].data_ptr(), roi.ctypes.data, off.ctypes.data, L,
            kw.ctypes.data, kh.ctypes.data, crops.data_ptr(), crops.numel(), hw.data_ptr(), coff.data_ptr(),
            ws.data_ptr(), ws.numel(), self.stream), "extract_lines")
        return crops, hw, coff

    # ---- a-7 -----------------------------------------------------------------------------
    def pack_crops(self, crops: list[np.ndarray], key: str = "crops"):
        """Host list of ragged HWC u8 crops -> (device buffer, hw int32 [L,2], off int64 [L]).  `key` names the
        pinned staging buffer: callers that pack several groups per call give each its own, so that packing group
        k + 1 overlaps the upload of group k."""
        hw = np.array([[c.shape[0], c.shape[1]] for c in crops], dtype=np.int32).reshape(-1, 2)
        sizes = [int(c.shape[0]) * int(c.shape[1]) * 3 for c in crops]
        off = np.zeros(len(crops), dtype=np.int64)
        if len(crops) > 1:
            off[1:] = np.cumsum([(s + 15) // 16 * 16 for s in sizes[:-1]])
        total = int(off[-1] + sizes[-1]) if crops else 0
        host = self._pinned(key, total)
        hv = host.numpy()
        # contiguous 3-channel crops go through the native gather (a few threads, GIL released once)
        native = [j for j, c in enumerate(crops) if sizes[j] and c.ndim == 3 and c.shape[2] == 3
                  and c.dtype == np.uint8 and c.flags["C_CONTIGUOUS"]]
        if self._gather_threads > 0 and total >= (2 << 20) and len(native) >= 2:
            ptrs = np.array([crops[j].ctypes.data for j in native], dtype=np.uint64)
            nb = np.array([sizes[j] for j in native], dtype=np.int64)
            do = np.ascontiguousarray(off[native])
            self._check(self.lib.bdrc_host_gather(ptrs.ctypes.data, nb.ctypes.data, do.ctypes.data, len(native),
                                                  host.data_ptr(), self._gather_threads), "host_gather")
            rest = sorted(set(range(len(crops))) - set(native))
        else:
            rest = range(len(crops))
        for j in rest:
            if sizes[j]:
                hv[off[j]:off[j] + sizes[j]] = np.ascontiguousarray(crops[j][..., :3]).reshape(-1)
        dev = self._upload(key, host)
        # crop geometry rides in its own small pinned buffer: a pageable .to() would synchronise the stream
        n = len(crops)
        meta = self._pinned(key + ".meta", 16 * n + 16)
        mv = meta.numpy()
        mv[:8 * n].view(np.int32)[:] = hw.reshape(-1)
        mv[8 * n:16 * n].view(np.int64)[:] = off
        meta_dev = self._upload(key + ".meta", meta)
        return dev, meta_dev[:8 * n].view(torch.int32).view(n, 2), meta_dev[8 * n:16 * n].view(torch.int64)

    def ocr_prepare_lines(self, crops_dev, hw_dev, off_dev, L: int, Ht: int, Wt: int, pre_pad: bool,
                          want_f32: bool = False, out: torch.Tensor | None = None, want_bf16: bool = True):
        """`out` (optional): a contiguous bf16 [L, Wt, OCR_CPAD] slice of a larger line batch to fill.
        want_bf16=False (with want_f32): only the reference-layout fp32 [L, Ht, Wt] lines, any Ht."""
        if not want_bf16:
            if not want_f32 or out is not None:
                raise ValueError("ocr_prepare_lines: want_bf16=False needs want_f32=True and no out")
            f32 = torch.empty((L, Ht, Wt), dtype=torch.float32, device=self.device)
            ws = self._workspace("ocr_prep", int(self.lib.bdrc_ocr_prepare_workspace_bytes(L, Ht, Wt)))
            self._check(self.lib.bdrc_ocr_prepare_lines(
                self.ctx.h, crops_dev.data_ptr(), hw_dev.data_ptr(), off_dev.data_ptr(), L, Ht, Wt, int(pre_pad),
                OCR_CPAD, None, f32.data_ptr(), ws.data_ptr(), ws.numel(), self.stream), "ocr_prepare_lines")
            return None, f32
        if out is not None and (tuple(out.shape) != (L, Wt, OCR_CPAD) or out.dtype != torch.bfloat16
                                or not out.is_contiguous()):
            raise ValueError("ocr_prepare_lines: out must be a contiguous bf16 [L, Wt, OCR_CPAD] tensor")
        lines = out if out is not None else torch.empty((L, Wt, OCR_CPAD), dtype=torch.bfloat16, device=self.device)
        f32 = torch.empty((L, Ht, Wt), dtype=torch.float32, device=self.device) if want_f32 else None
        ws = self._workspace("ocr_prep", int(self.lib.bdrc_ocr_prepare_workspace_bytes(L, Ht, Wt)))
        self._check(self.lib.bdrc_ocr_prepare_lines(
            self.ctx.h, crops_dev.data_ptr(), hw_dev.data_ptr(), off_dev.data_ptr(), L, Ht, Wt, int(pre_pad),
            OCR_CPAD, lines.data_ptr(), f32.data_ptr() if want_f32 else None, ws.data_ptr(), ws.numel(),
            self.stream), "ocr_prepare_lines")
        return lines, f32

    # ---- a-8 / a-9 -----------------------------------------------------------------------
    def ocr_forward(self, model: Model, lines: torch.Tensor) -> torch.Tensor:
        L, Wt = int(lines.shape[0]), int(lines.shape[1])
        T, V = Wt // 4, model.n_out
        logits = torch.empty((L, T, V), dtype=torch.float32, device=self.device)
        ws = self._workspace("ocr", int(self.lib.bdrc_ocr_workspace_bytes(model.h, L, Wt)))
        lines = lines.contiguous()
        self._check(self.lib.bdrc_ocr_forward(self.ctx.h, model.h, lines.data_ptr(), L, Wt, logits.data_ptr(),
                                              ws.data_ptr(), ws.numel(), self.stream), "ocr_forward")
        return logits

    def ctc_greedy(self, logits: torch.Tensor, layout: int = LAYOUT_TV, blank: int = 0):
        logits = logits.contiguous()
        L = int(logits.shape[0])
        T, V = (int(logits.shape[1]), int(logits.shape[2])) if layout == LAYOUT_TV else \
            (int(logits.shape[2]), int(logits.shape[1]))
        ids = torch.empty((L, T), dtype=torch.int32, device=self.device)
        lens = torch.zeros((L,), dtype=torch.int32, device=self.device)
        self._check(self.lib.bdrc_ctc_greedy(self.ctx.h, logits.data_ptr(), L, V, T, layout, blank, ids.data_ptr(),
                                             lens.data_ptr(), self.stream), "ctc_greedy")
        return ids, lens

    def ctc_topk(self, logits: torch.Tensor, layout: int = LAYOUT_TV, min_logp: float = -5.0):
        """Beam-search front end: (tokens int32 [L,T,K], log-probs fp32 [L,T,K], counts int32 [L,T])."""
        logits = logits.contiguous()
        L = int(logits.shape[0])
        T, V = (int(logits.shape[1]), int(logits.shape[2])) if layout == LAYOUT_TV else \
            (int(logits.shape[2]), int(logits.shape[1]))
        K = int(self.lib.bdrc_ctc_topk_width())
        tok = torch.zeros((L, T, K), dtype=torch.int32, device=self.device)
        lp = torch.full((L, T, K), float("-inf"), dtype=torch.float32, device=self.device)
        cnt = torch.zeros((L, T), dtype=torch.int32, device=self.device)
        self._check(self.lib.bdrc_ctc_topk(self.ctx.h, logits.data_ptr(), L, V, T, layout, min_logp, tok.data_ptr(),
                                           lp.data_ptr(), cnt.data_ptr(), self.stream), "ctc_topk")
        return tok, lp, cnt

    # ---- a-10, GPU half ------------------------------------------------------------------------
    CENSUS_HDR_BYTES = 64

    def mask_census(self, mask: torch.Tensor, slice_width: int = 20) -> torch.Tensor:
        """Connected-component census of a line mask u8 [H, W] on the device (bdrc_mask_census) -> ONE u8 device
        buffer [64 B header | 4 * max_centers B centres | H B row occupancy]; parse_census() reads its host copy."""
        mask = mask.contiguous()
        H, W = int(mask.shape[0]), int(mask.shape[1])
        mc = int(self.lib.bdrc_census_max_centers())
        out = torch.empty((self.CENSUS_HDR_BYTES + 4 * mc + H,), dtype=torch.uint8, device=self.device)
        ws = self._workspace("census", int(self.lib.bdrc_mask_census_workspace_bytes(H, W, slice_width)))
        base = out.data_ptr()
        self._check(self.lib.bdrc_mask_census(self.ctx.h, mask.data_ptr(), H, W, slice_width,
                                              base + self.CENSUS_HDR_BYTES + 4 * mc, base, base + self.CENSUS_HDR_BYTES,
                                              ws.data_ptr(), ws.numel(), self.stream), "mask_census")
        return out

    def parse_census(self, buf: np.ndarray, H: int):
        """Host copy of mask_census()'s buffer -> dict(bbox, x_steps, n, centers, row_any, overflow)."""
        from .line_detection import parse_census

        return parse_census(buf, H, int(self.lib.bdrc_census_max_centers()))

    def mask_components(self, mask: torch.Tensor, max_comp: int = 4096):
        """8-connected components of a mask on the device -> int array [n, 6] {x0, y0, x1, y1, area, root} (host)."""
        mask = mask.contiguous()
        H, W = int(mask.shape[0]), int(mask.shape[1])
        n = torch.zeros((1,), dtype=torch.int32, device=self.device)
        stats = torch.empty((max_comp, 6), dtype=torch.int32, device=self.device)
        ws = self._workspace("components", int(self.lib.bdrc_mask_components_workspace_bytes(H, W)))
        self._check(self.lib.bdrc_mask_components(self.ctx.h, mask.data_ptr(), H, W, n.data_ptr(), stats.data_ptr(), max_comp,
                                                  ws.data_ptr(), ws.numel(), self.stream), "mask_components")
        k = int(n.item())
        if k > max_comp:
            raise _lib.BdrcError(f"mask_components: {k} components exceed max_comp = {max_comp}")
        return stats[:k].cpu().numpy()
