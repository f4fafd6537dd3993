"""Drop-in for the hot path of BDRC/inference.py: same class names, constructor signatures, stage
methods and (OpStatus, payload) conventions -- the numpy/cv2/onnxruntime internals are replaced
by libbdrc_b200.so kernels on a B200.  No CPU fallback: constructing any of these without the
CUDA library / a B200 raises.

Batched entry points (`predict_batch`, `run_batch`, `run_ocr_batch`) are additions for
throughput; the single-image methods give identical results.
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Sequence

import numpy as np
import torch

from .data import (CharsetEncoder, DewarpingResult, Encoding, KenLMConfig, LayoutDetectionConfig,  # noqa: F401
                   Line, LineDetectionConfig, OCRLine, OCRModelConfig, OpStatus)
from .engine import BLOCKING_EVENTS, LAYOUT_TV, LAYOUT_VT, OCR, POST_SIGMOID, POST_SOFTMAX, SEG, Engine
from .weights import MAGIC, load_reference_checkpoint

_ENGINES: dict[int, Engine] = {}


def get_engine(device: int | None = None) -> Engine:
    """One Engine (ctx + workspaces) per process and GPU. LOCAL_RANK selects the GPU by default."""
    if device is None:
        device = int(os.environ.get("BDRC_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _ENGINES:
        _ENGINES[device] = Engine(device)
    return _ENGINES[device]


def _read_weights(src):
    """dict of tensors / BDRCW1 bytes / path of a BDRCW1 container / path of a torch checkpoint of the
    reference's torch path (an un-folded state dict: BatchNorm is folded here, weights.load_reference_checkpoint)."""
    if isinstance(src, dict) or isinstance(src, (bytes, bytearray)):
        return src
    with open(src, "rb") as f:
        head = f.read(8)
        if head == MAGIC:
            return head + f.read()
    if str(src).endswith((".pth", ".pt", ".ckpt")):
        return load_reference_checkpoint(src)
    raise ValueError(f"{src}: not a BDRCW1 weight container (convert the ONNX / checkpoint first, INTEGRATION.md)")


class CTCDecoder:
    """CTC decoder with the reference's vocabulary conventions (BDRC/inference.py:75-118): index 0 is '<blk>' when
    add_blank, label i maps to charset[i-1].  ctc_decode is the greedy best path (north_star); ctc_beam_decode is
    the beam search of pyctcdecode without a language model (spec: ocr_inference_b200/ctc_beam.py; the
    reference's fork is unpinned, so parity with it is unpinned)."""

    def __init__(self, charset: str | list[str], add_blank: bool, kenlm_config: KenLMConfig | None = None):
        self.blank_sign = "<blk>"
        self.ctc_beam_width = 64                      # (sic) set but never passed on by the reference (:84)
        self.charset = list(charset) if isinstance(charset, str) else charset
        self.ctc_vocab = self.charset.copy()
        if add_blank:
            self.ctc_vocab.insert(0, self.blank_sign)
        self.blank_index = self.ctc_vocab.index(self.blank_sign) if self.blank_sign in self.ctc_vocab else 0
        if kenlm_config is not None:
            # LM-fused beam search needs KenLM + the un-vendored pyctcdecode fork (requirements.txt:13): out of scope
            print("KenLM disabled: beam search with a language model is not part of the B200 hot path")

    def encode(self, label: str):
        return [self.charset.index(x) + 1 for x in label]

    def decode(self, inputs: Sequence[int]) -> str:
        return "".join(self.charset[x - 1] for x in inputs)

    def ids_to_text(self, ids: Sequence[int]) -> str:
        return "".join(self.ctc_vocab[i] for i in ids).replace(self.blank_sign, "")

    def _device_logits(self, logits):
        """(T, V) or (V, T) array-like -> (device tensor [1, ., .], layout): the reference transposes when the
        first axis has the vocabulary's length (BDRC/inference.py:373-376, 387-390)."""
        eng = get_engine()
        a = np.ascontiguousarray(logits, dtype=np.float32)
        if a.ndim != 2:
            raise ValueError(f"logits must be 2-D (time x vocab or vocab x time), got shape {a.shape}")
        layout = LAYOUT_VT if a.shape[0] == len(self.ctc_vocab) else LAYOUT_TV
        return eng, torch.from_numpy(a).to(eng.device)[None], layout

    def ctc_decode(self, logits) -> str:
        """logits (T, V) or (V, T) array-like -> text; runs the greedy kernel on the GPU."""
        eng, t, layout = self._device_logits(logits)
        ids, lens = eng.ctc_greedy(t, layout=layout, blank=self.blank_index)
        n = int(lens[0].item())
        return self.ids_to_text(ids[0, :n].tolist())

    def beams_from_candidates(self, tokens: np.ndarray, logps: np.ndarray, counts: np.ndarray):
        from .ctc_beam import beams_to_output, merge_frames

        return beams_to_output(merge_frames(tokens, logps, counts, self.blank_index), self.ctc_vocab, self.blank_index)

    def ctc_beam_decode(self, logits):
        """logits (T, V) or (V, T) -> list of OutputBeam (text, logit_score, ...), best first."""
        eng, t, layout = self._device_logits(logits)
        tok, lp, cnt = eng.ctc_topk(t, layout=layout)
        return self.beams_from_candidates(tok[0].cpu().numpy(), lp[0].cpu().numpy(), cnt[0].cpu().numpy())


class Detection:
    """BDRC/inference.py:121-163.  config.onnx_file names a BDRCW1 weight container."""

    _mode = POST_SIGMOID
    _default_threshold = 0.8

    def __init__(self, config: LineDetectionConfig | LayoutDetectionConfig, engine: Engine | None = None):
        self.config = config
        self._config_file = config
        self._onnx_model_file = config.onnx_file
        self._patch_size = config.patch_size
        if self._patch_size != 512:
            raise ValueError("the B200 segmentation path is built for 512x512 patches")
        self._classes = config.classes
        self.engine = engine or get_engine()
        n_out = 1 if self._mode == POST_SIGMOID else len(config.classes)
        self._inference = self.engine.load_model(SEG, _read_weights(config.onnx_file), n_out)
        self._tls = threading.local()   # .last = (host image, its device copy) of this thread's latest predict()

    # device-resident stages ----------------------------------------------------------------
    def _predict_device(self, page_dev: torch.Tensor, class_threshold: float) -> torch.Tensor:
        tiles, meta = self.engine.seg_preprocess(page_dev, self._patch_size)            # a-1
        if (meta["Hc"], meta["Wc"]) == (meta["H0"], meta["W0"]):
            low = self.engine.seg_forward(self._inference, tiles, lowres=True)           # a-2
            return self.engine.seg_postprocess(low, meta, self._mode, class_threshold, lowres=True)   # a-3
        logits = self.engine.seg_forward(self._inference, tiles)
        return self.engine.seg_postprocess(logits, meta, self._mode, class_threshold)

    def _predict_device_batch(self, pages_dev: Sequence[torch.Tensor], class_threshold: float, census: bool = False):
        """Several device-resident pages through ONE forward call (tiles of all pages share the
        batch dimension; per-tile reductions are per batch element, so this is exact).
        census=True (single-channel masks only) also returns, per page, the device buffer of the mask's
        connected-component census (Engine.mask_census: row occupancy + the slice census of get_line_threshold)."""
        eng = self.engine
        tiles, metas = eng.seg_preprocess_batch(list(pages_dev), self._patch_size)         # a-1, one launch
        need_full = any((m["Hc"], m["Wc"]) != (m["H0"], m["W0"]) for m in metas)
        logits = eng.seg_forward(self._inference, tiles, lowres=not need_full)             # a-2
        if not need_full:
            outs = eng.seg_postprocess_lowres_batch(logits, metas, self._mode, class_threshold)   # a-3, one launch
        else:
            outs, o = [], 0
            for meta in metas:
                n = meta["nty"] * meta["ntx"]
                outs.append(eng.seg_postprocess(logits[o:o + n], meta, self._mode, class_threshold))
                o += n
        if not census:
            return outs
        return outs, [eng.mask_census(m) if m.dim() == 2 else None for m in outs]         # a-10, GPU half

    def predict(self, image: np.ndarray, class_threshold: float | None = None) -> np.ndarray:
        thr = self._default_threshold if class_threshold is None else class_threshold
        ring = getattr(self._tls, "ring", None)          # per calling thread: predict() may run on several
        if ring is None:
            ring = self._tls.ring = [[None, None], [None, None]]
        self._tls.ring_k = getattr(self._tls, "ring_k", 0) + 1
        page, _ = self.engine.upload_page(np.ascontiguousarray(image[..., :3], dtype=np.uint8), ring, self._tls.ring_k)
        self._tls.last = (image, page)
        mask_dev = self._predict_device(page, thr)
        if getattr(self, "with_census", False) and mask_dev.dim() == 2:
            # the mask leaves with its connected-component census (computed on the GPU while the mask is still
            # there); line_detection.build_raw_line_data / get_line_threshold pick it up from the array
            from .line_detection import Mask3, parse_census

            cen = self.engine.mask_census(mask_dev).cpu().numpy()
            mask = mask_dev.cpu().numpy().view(Mask3)
            mask.census = parse_census(cen, int(mask.shape[0]))
            return mask
        return mask_dev.cpu().numpy()

    def predict_batch(self, images: Sequence[np.ndarray], class_threshold: float | None = None,
                      pinned: Sequence[torch.Tensor] | None = None, copy: bool = True,
                      group_sizes: Sequence[int] | None = None) -> list[np.ndarray]:
        """Same result as [predict(i) for i in images], as a three-stage pipeline: a helper thread moves the
        (pageable) pages through a ring of pinned slots onto the copy-in stream, this thread launches one forward
        per group of pages as soon as the group's uploads are queued, and a second helper thread hands the masks
        out as their group's D2H copy completes.  With copy=False the returned arrays alias this object's pinned
        staging buffers and stay valid until the next predict_batch call (the io-binding style of use);
        copy=True (default) hands out fresh arrays like the reference."""
        thr = self._default_threshold if class_threshold is None else class_threshold
        eng = self.engine
        dev = eng.device
        n = len(images)
        if n == 0:
            return []
        compute = torch.cuda.current_stream(dev)
        copy_in = getattr(self, "_copy_in", None) or torch.cuda.Stream(dev)
        copy_out = getattr(self, "_copy_out", None) or torch.cuda.Stream(dev)
        self._copy_in, self._copy_out = copy_in, copy_out
        if not hasattr(self, "_host_bufs"):
            self._host_bufs = {}
        ups: list = [None] * n
        evs: list = [None] * n
        # the upload is the only host work of this call: all the copy threads the box can spare (measured on 16 cores,
        # 16 pages: 35.2 ms per call with 4 threads, 32.3 ms with 8)
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))      # processes sharing this host's cores
        up_threads = getattr(eng, "upload_threads", None) or max(2, min(8, (os.cpu_count() or 8) // (2 if ranks == 1 else ranks)))
        ready = [threading.Event() for _ in range(n)]
        failed: list = []
        copy_in.wait_stream(compute)

        def uploader():
            torch.cuda.set_device(dev)
            ring = getattr(self, "_up_slots", None)
            if ring is None:
                ring = self._up_slots = [[None, None] for _ in range(4)]
            try:
                with torch.cuda.stream(copy_in):
                    for i, img in enumerate(images):
                        if pinned is not None:
                            ups[i] = pinned[i].to(dev, non_blocking=True)
                            evs[i] = torch.cuda.Event()
                            evs[i].record(copy_in)
                        else:       # pageable page: through the pinned slot ring (native copy threads, GIL released)
                            ups[i], evs[i] = eng.upload_page(np.ascontiguousarray(img[..., :3], dtype=np.uint8), ring, i,
                                                             threads=up_threads)
                        ready[i].set()
            except Exception as e:  # noqa: BLE001
                failed.append(e)
            finally:
                for r in ready:
                    r.set()

        up_thread = threading.Thread(target=uploader, name="bdrc-upload", daemon=True)
        up_thread.start()
        # pages per forward call: a small first group starts computing after two uploads, then groups of 6
        # landscape pages (144 tiles: one 148-tile chunk, ~whole waves per ResNet stage); later groups upload meanwhile
        sizes, left = [], n
        plan = list(group_sizes) if group_sizes else list(getattr(self, "group_plan", None) or [2])
        while left > 0:
            want = plan[len(sizes)] if len(sizes) < len(plan) else 6
            sizes.append(max(1, min(left, int(want))))
            left -= sizes[-1]
        starts = [sum(sizes[:i]) for i in range(len(sizes))]
        outs: list = [None] * n
        group_done: "list[tuple[int, int, torch.cuda.Event, list]]" = []
        hand_out = threading.Semaphore(0)

        def finisher():
            # fresh, caller-owned arrays: one native multi-threaded copy per group out of the pinned staging buffers
            for _ in sizes:
                hand_out.acquire()
                if failed:
                    return
                g0, group, ev, hosts = group_done.pop(0)
                ev.synchronize()
                if not copy:
                    for k, h in enumerate(hosts):
                        outs[g0 + k] = h.numpy()
                    continue
                fresh = [np.empty(tuple(h.shape), dtype=np.uint8) for h in hosts]
                eng.host_copy_many([h.data_ptr() for h in hosts], [f.ctypes.data for f in fresh], [f.nbytes for f in fresh])
                for k, f in enumerate(fresh):
                    outs[g0 + k] = f

        fin_thread = threading.Thread(target=finisher, name="bdrc-masks", daemon=True)
        fin_thread.start()
        try:
            for g0, group in zip(starts, sizes):
                for i in range(g0, g0 + group):
                    if not ready[i].wait(timeout=300.0):
                        raise TimeoutError("page upload did not finish")
                    if failed:
                        raise failed[0]
                    compute.wait_event(evs[i])
                    ups[i].record_stream(compute)
                masks = self._predict_device_batch(ups[g0:g0 + group], thr)
                done = torch.cuda.Event()
                done.record(compute)
                # masks go home on their own stream: a D2H copy queued on the compute stream would hold back the
                # next group's kernels for its whole duration
                copy_out.wait_event(done)
                hosts = []
                with torch.cuda.stream(copy_out):
                    for k, m in enumerate(masks):
                        key = ("mask_host", g0 + k, tuple(m.shape))
                        h = self._host_bufs.get(key)
                        if h is None:
                            h = torch.empty(m.shape, dtype=torch.uint8, pin_memory=True)
                            self._host_bufs[key] = h
                        h.copy_(m, non_blocking=True)
                        m.record_stream(copy_out)
                        hosts.append(h)
                    ev = torch.cuda.Event(blocking=BLOCKING_EVENTS)     # the hand-out thread sleeps on it
                    ev.record(copy_out)
                group_done.append((g0, group, ev, hosts))
                hand_out.release()
        except Exception as e:  # noqa: BLE001
            failed.append(e)
            for _ in sizes:
                hand_out.release()
        up_thread.join()
        fin_thread.join()
        compute.synchronize()
        if failed:
            raise failed[0]
        return outs


class LineDetection(Detection):
    """BDRC/inference.py:166-183: sigmoid > 0.9 -> (H, W) u8 {0,255}."""

    _mode = POST_SIGMOID
    _default_threshold = 0.9


class LayoutDetection(Detection):
    """BDRC/inference.py:186-276: softmax > 0.8 -> (H, W, C) u8 {0,255}."""

    _mode = POST_SOFTMAX
    _default_threshold = 0.8

    def __init__(self, config: LayoutDetectionConfig, debug: bool = False, engine: Engine | None = None):
        super().__init__(config, engine)
        self._debug = debug


class OCRInference:
    """BDRC/inference.py:279-429.  config.model_file names a BDRCW1 container of Easter2 weights."""

    def __init__(self, ocr_config: OCRModelConfig, kenlm_config: KenLMConfig | None = None,
                 engine: Engine | None = None):
        self.config = ocr_config
        self._onnx_model_file = ocr_config.model_file
        self._input_width = ocr_config.input_width
        self._input_height = ocr_config.input_height
        self._input_layer = ocr_config.input_layer
        self._output_layer = ocr_config.output_layer
        self._characters = ocr_config.charset
        self._squeeze_channel_dim = ocr_config.squeeze_channel
        self._swap_hw = ocr_config.swap_hw
        self._add_blank = ocr_config.add_blank
        self.engine = engine or get_engine()
        self.ctc_decoder = CTCDecoder(self._characters, self._add_blank, kenlm_config=None)
        self.ctc_decoder_lm = None
        vocab = len(self.ctc_decoder.ctc_vocab)
        self.ocr_session = self.engine.load_model(OCR, _read_weights(ocr_config.model_file), vocab)

    def _check_fit(self, img: np.ndarray, pre_pad: bool):
        h, w = img.shape[:2]
        if pre_pad:
            w = w + 2 * h
        Wt, Ht = self._input_width, self._input_height
        wr, hr = Wt / w, Ht / h
        rh, rw = (int(h * wr), Wt) if wr <= hr else (Ht, int(w * hr))
        if rh <= 0 or rw <= 0:
            # cv2.resize raises on an empty target in the reference (utils.py:468-476)
            raise ValueError(f"line image {img.shape} collapses to an empty OCR input")

    def run_batch(self, line_images: Sequence[np.ndarray], pre_pad: bool = True, use_lm: bool = False,
                  return_ids: bool = False):
        """All lines of a page (or many pages) in one pass: prep -> Easter2 -> greedy CTC."""
        if len(line_images) == 0:
            return []
        for img in line_images:
            self._check_fit(img, pre_pad)
        eng = self.engine
        # (packing in four groups so that the host copy of group k + 1 overlaps the upload of group k was measured:
        #  no gain -- the call is bound by ~16 GB/s of host copy into pinned memory, tools/pack_bench.py)
        dev, hw, off = eng.pack_crops(list(line_images))
        lines, _ = eng.ocr_prepare_lines(dev, hw, off, len(line_images), self._input_height, self._input_width,
                                         pre_pad)                                          # a-7
        logits = eng.ocr_forward(self.ocr_session, lines)                                  # a-8
        ids, lens = eng.ctc_greedy(logits, layout=LAYOUT_TV, blank=self.ctc_decoder.blank_index)   # a-9
        ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
        seqs = [ids[i, : lens[i]].tolist() for i in range(len(line_images))]
        if return_ids:
            return seqs
        return [self.ctc_decoder.ids_to_text(s) for s in seqs]

    def run(self, line_image: np.ndarray, pre_pad: bool = True, use_lm: bool = False) -> str:
        return self.run_batch([line_image], pre_pad, use_lm)[0]

    # ---- the reference's lower-level entry points (BDRC/inference.py:359-414) -----------------------------
    def _predict(self, image_batch: np.ndarray) -> np.ndarray:
        """A prepared line in ANY of the reference's input layouts -- [1, H, W], [1, W, H] (swap_hw) or
        [1, 1, H, W] (squeeze_channel off), values in {-1, +1} -- -> logits [T, V] (np.squeeze'd like :368).
        swap_hw / squeeze_channel only describe how an ONNX graph wants the same image handed over; the B200
        network consumes the line in its own time-major layout, so they change nothing but this unpacking."""
        x = np.asarray(image_batch, dtype=np.float32)
        if not self._squeeze_channel_dim:
            if x.ndim != 4 or x.shape[1] != 1:
                raise ValueError(f"squeeze_channel is off: expected [1, 1, H, W], got {x.shape}")
            x = x[:, 0]
        if x.ndim != 3:
            raise ValueError(f"expected a [1, H, W] line batch, got {x.shape}")
        if self._swap_hw:
            x = np.transpose(x, (0, 2, 1))
        if x.shape[1:] != (self._input_height, self._input_width):
            raise ValueError(f"line batch {x.shape} does not match the model input {self._input_height}x{self._input_width}")
        from .engine import OCR_CPAD

        eng = self.engine
        lines = torch.zeros((x.shape[0], self._input_width, OCR_CPAD), dtype=torch.bfloat16, device=eng.device)
        lines[:, :, : self._input_height] = torch.from_numpy(np.ascontiguousarray(x)).to(eng.device).permute(0, 2, 1)
        return np.squeeze(eng.ocr_forward(self.ocr_session, lines).cpu().numpy())

    def _decode(self, logits: np.ndarray, use_lm: bool = False) -> str:
        return self.ctc_decoder.ctc_decode(logits)          # [V, T] inputs are transposed there (:373-376)

    def _decode_beams(self, logits: np.ndarray, use_lm: bool = False):
        return self.ctc_decoder.ctc_beam_decode(logits)

    def run_beam_batch(self, line_images: Sequence[np.ndarray], pre_pad: bool = True) -> list:
        """Beam-search decode of many lines: a-7 and a-8 as in run_batch, per-step candidates on the GPU
        (bdrc_ctc_topk), prefix merge per line on the host.  -> list (per line) of lists of OutputBeam."""
        if len(line_images) == 0:
            return []
        for img in line_images:
            self._check_fit(img, pre_pad)
        eng = self.engine
        dev, hw, off = eng.pack_crops(list(line_images))
        lines, _ = eng.ocr_prepare_lines(dev, hw, off, len(line_images), self._input_height, self._input_width, pre_pad)
        tok, lp, cnt = eng.ctc_topk(eng.ocr_forward(self.ocr_session, lines), layout=LAYOUT_TV)
        tok, lp, cnt = tok.cpu().numpy(), lp.cpu().numpy(), cnt.cpu().numpy()
        return [self.ctc_decoder.beams_from_candidates(tok[i], lp[i], cnt[i]) for i in range(len(line_images))]

    def run_beam_code(self, line_image: np.ndarray, pre_pad: bool = True) -> list:
        """BDRC/inference.py:400-414."""
        return self.run_beam_batch([line_image], pre_pad)[0]

    def run_packed(self, crops_dev, hw_dev, off_dev, n: int, pre_pad: bool):
        """Lines already resident on the GPU (from extract_line_images): text per line."""
        eng = self.engine
        if n == 0:
            return []
        lines, _ = eng.ocr_prepare_lines(crops_dev, hw_dev, off_dev, n, self._input_height, self._input_width, pre_pad)
        logits = eng.ocr_forward(self.ocr_session, lines)
        ids, lens = eng.ctc_greedy(logits, layout=LAYOUT_TV, blank=self.ctc_decoder.blank_index)
        ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
        return [self.ctc_decoder.ids_to_text(ids[i, : lens[i]].tolist()) for i in range(n)]


class _IdentityConverter:
    """Stand-in for pyewts.pyewts() (Wylie <-> Unicode conversion is third-party host code outside
    the hot path); a real converter can be assigned to OCRPipeline.converter."""

    def toUnicode(self, s):
        return s

    def toWylie(self, s):
        return s


class OCRPipeline:
    """BDRC/inference.py:432-826 -- same constructor, stage methods and (OpStatus, payload) returns."""

    def __init__(self, ocr_config: OCRModelConfig, line_config: LineDetectionConfig | LayoutDetectionConfig,
                 kenlm_config: KenLMConfig | None = None, use_line_prepadding: bool = False,
                 engine: Engine | None = None):
        self.ready = False
        self.ocr_model_config = ocr_config
        self.line_config = line_config
        self.encoder = ocr_config.encoder
        self.engine = engine or get_engine()
        self.ocr_inference = OCRInference(self.ocr_model_config, kenlm_config=kenlm_config, engine=self.engine)
        try:
            import pyewts  # type: ignore

            self.converter = pyewts.pyewts()
        except Exception:  # noqa: BLE001
            self.converter = _IdentityConverter()
        self.use_line_prepadding = use_line_prepadding
        self._tls = threading.local()   # .fast: run_ocr keeps line crops on the device between stages 4 and 5
        if isinstance(self.line_config, LineDetectionConfig):
            self.line_inference = LineDetection(self.line_config, engine=self.engine)
            self.line_inference.with_census = True      # masks travel with their GPU census (a-10, GPU half)
            self.ready = True
        elif isinstance(self.line_config, LayoutDetectionConfig):
            self.line_inference = LayoutDetection(self.line_config, engine=self.engine)
            self.ready = True
        else:
            self.line_inference = None
            self.ready = False

    def update_ocr_model(self, config: OCRModelConfig, kenlm_config: KenLMConfig | None = None):
        self.ocr_model_config = config
        self.ocr_inference = OCRInference(config, kenlm_config, engine=self.engine)

    def update_line_detection(self, config):
        # inference.py:469-482: only swaps when the config TYPE changes
        if isinstance(config, LineDetectionConfig) and isinstance(self.line_config, LayoutDetectionConfig):
            self.line_inference = LineDetection(config, engine=self.engine)
        elif isinstance(config, LayoutDetectionConfig) and isinstance(self.line_config, LineDetectionConfig):
            self.line_inference = LayoutDetection(config, engine=self.engine)

    # ---- stage 1 (inference.py:488-506) --------------------------------------------------
    def detect_lines(self, image: np.ndarray):
        if isinstance(self.line_config, LineDetectionConfig) and self.line_inference is not None:
            line_mask = self.line_inference.predict(image)
        elif isinstance(self.line_config, LayoutDetectionConfig) and self.line_inference is not None:
            layout_mask = self.line_inference.predict(image)
            line_mask = layout_mask[:, :, self.line_config.classes.index("line")]
        return OpStatus.SUCCESS, line_mask

    # ---- stage 2 (inference.py:508-533) --------------------------------------------------
    def build_lines(self, image: np.ndarray, line_mask: np.ndarray):
        from .line_detection import build_raw_line_data, filter_line_contours

        rot_img, rot_mask, line_contours, page_angle = build_raw_line_data(image, line_mask)
        if len(line_contours) == 0:
            return OpStatus.FAILED, "No lines detected"
        filtered = filter_line_contours(rot_mask, line_contours)
        if len(filtered) == 0:
            return OpStatus.FAILED, "No valid lines after filtering"
        return OpStatus.SUCCESS, (rot_img, rot_mask, line_contours, filtered, page_angle)

    # ---- stage 3 (inference.py:535-589) --------------------------------------------------
    def apply_dewarping(self, rot_img, rot_mask, filtered_contours, page_angle, use_tps: bool = False,
                        tps_threshold: float = 0.25, spline_order: int = 1):
        """`spline_order` (addition): 1 = bilinear resampling (north_star), 3 = scipy's default cubic B-spline,
        what the reference's map_coordinates call computes."""
        import cv2

        from .image_dewarping import apply_global_tps, check_for_tps
        from .line_detection import build_raw_line_data, filter_line_contours

        if not use_tps:
            return OpStatus.SUCCESS, DewarpingResult(rot_img, rot_mask, filtered_contours, page_angle, False)
        ratio, tps_line_data = check_for_tps(rot_img, filtered_contours)
        if ratio <= tps_threshold:
            return OpStatus.SUCCESS, DewarpingResult(rot_img, rot_mask, filtered_contours, page_angle, False,
                                                     tps_ratio=ratio)
        dew_img, dew_mask = apply_global_tps(rot_img, rot_mask, tps_line_data, order=spline_order)
        if dew_mask.ndim == 3:
            dew_mask = cv2.cvtColor(dew_mask, cv2.COLOR_RGB2GRAY)
        d_img, d_mask, contours, new_angle = build_raw_line_data(dew_img, dew_mask)
        return OpStatus.SUCCESS, DewarpingResult(d_img, d_mask, filter_line_contours(d_mask, contours), new_angle,
                                                 True, tps_ratio=ratio, dewarped_img=dew_img, dewarped_mask=dew_mask)

    # ---- stage 4 (inference.py:591-616) --------------------------------------------------
    def extract_lines(self, work_img, rot_mask, filtered_contours, merge_lines: bool = True, k_factor: float = 2.5,
                      bbox_tolerance: float = 4.0):
        from .line_detection import build_line_data, extract_line_images, sort_lines_by_threshold2

        line_data = [build_line_data(c) for c in filtered_contours]
        sorted_lines, _ = sort_lines_by_threshold2(rot_mask, line_data, group_lines=merge_lines)
        # the page uploaded by detect_lines is still resident when the page was not rotated / dewarped
        last = getattr(self.line_inference._tls, "last", None) if self.line_inference is not None else None
        page_dev = last[1] if last is not None and last[0] is work_img and work_img.ndim == 3 \
            and work_img.shape[2] == 3 else None
        line_images = extract_line_images(work_img, sorted_lines, k_factor, bbox_tolerance, page_dev=page_dev,
                                          download=not getattr(self._tls, "fast", False))
        if not line_images:
            return OpStatus.FAILED, "No valid line images extracted"
        return OpStatus.SUCCESS, (sorted_lines, line_images)

    # ---- stage 5 (inference.py:618-666) --------------------------------------------------
    def _convert(self, pred: str, target_encoding) -> str:
        if self.encoder == CharsetEncoder.WYLIE and target_encoding == Encoding.UNICODE:
            return self.converter.toUnicode(pred)
        if self.encoder == CharsetEncoder.STACK and target_encoding == Encoding.WYLIE:
            return self.converter.toWylie(pred)
        return pred

    def run_text_recognition(self, line_images: list, sorted_lines: list, target_encoding: Encoding = Encoding.UNICODE):
        """All non-empty lines of the page go through the recogniser as ONE batch."""
        keep_idx = [i for i, img in enumerate(line_images) if img.shape[0] != 0 and img.shape[1] != 0]
        keep = [(line_images[i], sorted_lines[i]) for i in keep_idx]
        pack = getattr(line_images, "device_pack", None)
        if keep and pack is not None and list(pack[3]) == keep_idx and len(line_images) == len(sorted_lines):
            # crops are still on the GPU in exactly this order: no download / re-upload
            for img, _ in keep:
                self.ocr_inference._check_fit(img, self.use_line_prepadding)
            texts = self.ocr_inference.run_packed(pack[0], pack[1], pack[2], len(keep), self.use_line_prepadding)
        else:
            texts = self.ocr_inference.run_batch([np.asarray(k[0]) for k in keep], self.use_line_prepadding) if keep else []
        return OpStatus.SUCCESS, self._make_ocr_lines([k[1] for k in keep], texts, target_encoding)

    def run_text_recognition_eval(self, line_images: list, sorted_lines: list,
                                  target_encoding: Encoding = Encoding.UNICODE, top_k_beams: int = 10):
        """Stage 5 in eval mode (BDRC/inference.py:668-726): beam-search decode, the top beam's text (spaces
        removed, section sign -> space), ctc_conf = exp(length-normalised log score), logits = the beams' scores."""
        import math

        keep = [(np.asarray(img), ln) for img, ln in zip(line_images, sorted_lines)
                if img.shape[0] != 0 and img.shape[1] != 0]
        all_beams = self.ocr_inference.run_beam_batch([k[0] for k in keep], self.use_line_prepadding) if keep else []
        ocr_lines = []
        for (_, ln), beams in zip(keep, all_beams):
            if not beams:
                continue
            beams = beams[:top_k_beams]
            pred = beams[0].text.strip().replace(" ", "").replace("§", " ")
            pred = self._convert(pred, target_encoding)
            norm_logp = beams[0].logit_score / max(len(beams[0].text), 1)
            ocr_lines.append(OCRLine(
                guid=ln.guid, text=pred,
                encoding=Encoding.WYLIE.name if target_encoding == Encoding.WYLIE._name_ else Encoding.UNICODE.name,
                ctc_conf=float(math.exp(norm_logp)), logits=[float(b.logit_score) for b in beams], lm_scores=None))
        return OpStatus.SUCCESS, ocr_lines

    def _make_ocr_lines(self, lines: list, texts: list, target_encoding) -> list:
        ocr_lines = []
        for ln, t in zip(lines, texts):
            pred = self._convert(t.strip().replace("§", " "), target_encoding)
            ocr_lines.append(OCRLine(
                guid=ln.guid, text=pred,
                # (sic) the reference compares the enum with its name, so this is always UNICODE (:655-659)
                encoding=Encoding.WYLIE.name if target_encoding == Encoding.WYLIE.name else Encoding.UNICODE.name,
                ctc_conf=None, logits=None, lm_scores=None))
        return ocr_lines

    # ---- whole page (inference.py:732-826) ------------------------------------------------
    def run_ocr(self, image: np.ndarray, k_factor: float = 2.5, bbox_tolerance: float = 4.0, merge_lines: bool = True,
                use_tps: bool = False, tps_threshold: float = 0.25, target_encoding: Encoding = Encoding.UNICODE,
                eval_mode: bool = False):
        self._tls.fast = True    # line crops are not part of run_ocr's result: they can stay on the GPU
        try:
            return self._run_ocr(image, k_factor, bbox_tolerance, merge_lines, use_tps, tps_threshold, target_encoding,
                                 eval_mode)
        finally:
            self._tls.fast = False

    def _run_ocr(self, image, k_factor, bbox_tolerance, merge_lines, use_tps, tps_threshold, target_encoding, eval_mode):
        try:
            if not self.ready:
                return OpStatus.FAILED, "OCR pipeline not ready"
            if image is None:
                return OpStatus.FAILED, "Input image is None"
            try:
                status, result = self.detect_lines(image)
                if status == OpStatus.FAILED:
                    return status, result
                line_mask = result
            except Exception as e:  # noqa: BLE001
                return OpStatus.FAILED, f"Line detection failed: {str(e)}"
            try:
                status, result = self.build_lines(image, line_mask)
                if status == OpStatus.FAILED:
                    return status, result
                rot_img, rot_mask, _, filtered_contours, page_angle = result
            except Exception as e:  # noqa: BLE001
                return OpStatus.FAILED, f"Line data building failed: {str(e)}"
            try:
                status, result = self.apply_dewarping(rot_img, rot_mask, filtered_contours, page_angle, use_tps=use_tps,
                                                      tps_threshold=tps_threshold)
                if status == OpStatus.FAILED:
                    return status, result
                dew = result
            except Exception as e:  # noqa: BLE001
                return OpStatus.FAILED, f"Line processing failed: {str(e)}"
            try:
                status, result = self.extract_lines(dew.work_img, rot_mask, dew.filtered_contours, merge_lines=merge_lines,
                                                    k_factor=k_factor, bbox_tolerance=bbox_tolerance)
                if status == OpStatus.FAILED:
                    return status, result
                sorted_lines, line_images = result
            except Exception as e:  # noqa: BLE001
                return OpStatus.FAILED, f"Line extraction failed: {str(e)}"
            try:
                if eval_mode:
                    status, result = self.run_text_recognition_eval(line_images, sorted_lines,
                                                                    target_encoding=target_encoding)
                else:
                    status, result = self.run_text_recognition(line_images, sorted_lines, target_encoding=target_encoding)
                if status == OpStatus.FAILED:
                    return status, result
                ocr_lines = result
            except Exception as e:  # noqa: BLE001
                return OpStatus.FAILED, f"OCR processing failed: {str(e)}"
            return OpStatus.SUCCESS, [rot_mask, sorted_lines, ocr_lines, float(page_angle)]
        except Exception as e:  # noqa: BLE001
            return OpStatus.FAILED, f"OCR pipeline failed: {str(e)}"

    def run_ocr_batch(self, images: Sequence[np.ndarray], workers: int | None = None, staged: bool = True,
                      chunk_pages: int = 48, geometry_procs: int | None = None, **kw):
        """Pages are independent units: same results as [run_ocr(i, **kw) for i in images].

        staged=True (default) runs the stages of run_ocr across pages instead of page by page:
        uploads run ahead on a copy stream, the segmentation network sees groups of up to 6 pages per
        forward call (whole waves on 148 SMs), `workers` host threads do the cv2 geometry and the line
        extraction of one page each on their own CUDA stream while the next group is in the network, and
        the recogniser reads the crops of ALL pages of a chunk as one batch.  Batching is exact (per-tile
        and per-line reductions only), so the result equals the page-by-page one.  staged=False keeps the
        page-per-thread schedule.

        `geometry_procs` > 0 moves the cv2 geometry of the staged schedule (contours, reading order, ROI masks:
        line_detection.page_geometry) into that many worker PROCESSES (geometry_pool.GeometryPool; no CUDA in
        them, masks through shared memory), which is what lets one GPU-owning process go past ~120 pages/s;
        None = 8 (half the cores at most) once a call brings at least 16 pages, 0 = threads only.  Like any
        `spawn`-based pool it needs the caller's script to be import-safe (`if __name__ == "__main__":`)."""
        workers_given = workers
        if workers is None:
            # measured on a 16-core B200 host: 2 geometry threads beside the launching and the upload thread are
            # the best staged schedule (120 pages/s; 4: 71), 4 page threads the best page-by-page one (113)
            workers = max(1, min(len(images), len(os.sched_getaffinity(0)), 2 if staged else 4))
        if len(images) <= 1 or kw.get("eval_mode") or not self.ready:
            return [self.run_ocr(img, **kw) for img in images]
        if staged:
            eng = self.engine
            if not getattr(eng, "blocking_sync", False) and os.environ.get("BDRC_SPIN_SYNC", "0") != "1":
                # ~10 host threads per GPU wait on the device in this schedule: they sleep instead of spinning on
                # cores the geometry (and, on a full box, the other GPUs' pipelines) need
                eng.set_blocking_sync(True)
            gpool = self._geometry_pool(geometry_procs, len(images)) if not kw.get("use_tps") else None
            if gpool is not None and workers_given is None:
                workers = len(gpool)            # these threads mostly wait for their page's geometry
            import cv2

            # OpenCV's own thread pool spins after every parallel_for: with one small cv2 call after another on
            # several page threads that burnt ~7 ms of CPU per page for nothing (measured: 6.6 -> 4.4 busy cores at
            # the same pages/s).  Pages are the unit of parallelism here, so cv2 runs single-threaded meanwhile.
            cv_threads = cv2.getNumThreads()
            cv2.setNumThreads(1)
            try:
                # chunks are software-pipelined: the segmentation of chunk k+1 is queued on the GPU before this
                # thread waits for the geometry of chunk k and runs its recognition (on a stream of its own), so the
                # device does not idle through the geometry tail, the text decode and the next chunk's first uploads
                out, pending = [], None
                for k, c0 in enumerate(range(0, len(images), chunk_pages)):
                    fin = self._run_ocr_staged(images[c0:c0 + chunk_pages], workers, gpool, _slot=k & 1, _defer=True, **kw)
                    if pending is not None:
                        out.extend(pending())
                    pending = fin
                if pending is not None:
                    out.extend(pending())
                return out
            finally:
                cv2.setNumThreads(cv_threads)
        if workers <= 1:
            return [self.run_ocr(img, **kw) for img in images]
        dev = self.engine.device
        tls = self._tls

        def one(img):
            if not hasattr(tls, "stream"):
                torch.cuda.set_device(dev)
                tls.stream = torch.cuda.Stream(dev)
            with torch.cuda.stream(tls.stream):
                return self.run_ocr(img, **kw)

        return list(self._thread_pool(workers).map(one, images))

    def _geometry_pool(self, procs: int | None, n_images: int):
        """The GeometryPool of this pipeline (started on first use, kept for later calls), or None."""
        have = getattr(self, "_gpool", None)
        if procs == 0 or have is False:
            return None
        if procs is None:
            if have is not None:
                return have
            if n_images < 16:
                return None
            procs = max(2, min(8, len(os.sched_getaffinity(0)) // 2))
        if have is not None and len(have) == procs:
            return have
        from .geometry_pool import GeometryPool

        if have is not None:
            have.close()
        try:
            self._gpool = GeometryPool(procs)
        except Exception as e:  # noqa: BLE001  (e.g. a caller script that is not import-safe under `spawn`)
            print(f"geometry worker processes unavailable ({e}); using threads")
            self._gpool = False
            return None
        return self._gpool

    def close(self):
        """Stops the geometry worker processes and the page threads of run_ocr_batch."""
        g = getattr(self, "_gpool", None)
        if g:
            g.close()
        self._gpool = None
        p = getattr(self, "_pool", None)
        if p is not None:
            p.shutdown(wait=False)
            self._pool = None

    def _thread_pool(self, workers: int) -> ThreadPoolExecutor:
        pool = getattr(self, "_pool", None)
        if pool is None or pool._max_workers != workers:
            pool = self._pool = ThreadPoolExecutor(workers, thread_name_prefix="bdrc-page")
        return pool

    def _run_ocr_staged(self, images, workers: int, gpool=None, k_factor: float = 2.5, bbox_tolerance: float = 4.0,
                        merge_lines: bool = True, use_tps: bool = False, tps_threshold: float = 0.25,
                        target_encoding: Encoding = Encoding.UNICODE, eval_mode: bool = False, _slot: int = 0,
                        _defer: bool = False):
        """One chunk of pages through the staged schedule (see run_ocr_batch).  A page on which any
        stage raises is re-run through run_ocr, which owns the reference's error messages.
        _defer=True returns, once every segmentation group of the chunk is queued and its page stages are submitted,
        a function that collects them, runs the recognition and returns the results (run_ocr_batch calls it after it
        has queued the next chunk); _slot selects the set of pinned mask buffers (two chunks are in flight)."""
        eng, det = self.engine, self.line_inference
        dev = eng.device
        n = len(images)
        tls = self._tls
        results: list = [None] * n
        bad = [i for i, img in enumerate(images)
               if not isinstance(img, np.ndarray) or img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8
               or img.shape[0] == 0 or img.shape[1] == 0]
        good = [i for i in range(n) if i not in bad]
        for i in bad:
            results[i] = self.run_ocr(images[i], k_factor, bbox_tolerance, merge_lines, use_tps, tps_threshold,
                                      target_encoding, eval_mode)
        if not good:
            return (lambda: results) if _defer else results
        compute = torch.cuda.current_stream(dev)
        copy_in = getattr(det, "_copy_in", None) or torch.cuda.Stream(dev)
        copy_out = getattr(det, "_copy_out", None) or torch.cuda.Stream(dev)
        det._copy_in, det._copy_out = copy_in, copy_out
        if not hasattr(det, "_host_bufs"):
            det._host_bufs = {}
        layout_ch = self.line_config.classes.index("line") if isinstance(self.line_config, LayoutDetectionConfig) else None

        # ---- uploads: one thread fills the copy stream (pageable H2D releases the GIL) ----------
        pages_dev: dict[int, torch.Tensor] = {}
        up_ev: dict[int, torch.cuda.Event] = {}
        up_ready = {i: threading.Event() for i in good}
        copy_in.wait_stream(compute)

        def uploader():
            # pageable pages go through a ring of pinned slots (Engine.upload_page): a plain pageable .to() capped
            # the whole pipeline at ~270 pages/s (tools/c3_limits.py)
            torch.cuda.set_device(dev)
            ring = getattr(det, "_up_slots", None)
            if ring is None:
                ring = det._up_slots = [[None, None] for _ in range(4)]
            try:
                with torch.cuda.stream(copy_in):
                    for k, i in enumerate(good):
                        # a page whose upload fails (pinned-alloc OOM, CUDA error) is left without an entry in
                        # up_ev: its group fails, the page is retried through run_ocr, later pages still upload
                        try:
                            pages_dev[i], up_ev[i] = eng.upload_page(images[i], ring, k)
                        except Exception:  # noqa: BLE001
                            pages_dev.pop(i, None)
                            up_ev.pop(i, None)
                        finally:
                            up_ready[i].set()
            finally:
                for ev in up_ready.values():      # whatever happened above, nobody waits forever
                    ev.set()

        up_thread = threading.Thread(target=uploader, name="bdrc-upload", daemon=True)
        up_thread.start()

        # ---- host geometry + line extraction of one page (worker threads) ------------------------
        def page_stage(i: int, mask_host: np.ndarray, mask_ev: torch.cuda.Event, census_host=None):
            if not hasattr(tls, "stream"):
                torch.cuda.set_device(dev)
                # high priority: the small extraction kernels of a page go ahead of the next conv launch of the
                # segmentation group that is in flight, instead of queueing behind the whole forward
                tls.stream = torch.cuda.Stream(dev, priority=-1)
            tls.fast = True
            try:
                with torch.cuda.stream(tls.stream):
                    mask_ev.synchronize()
                    image = images[i]
                    line_mask = mask_host if layout_ch is None else mask_host[:, :, layout_ch]
                    det._tls.last = (image, pages_dev[i])
                    census = None
                    if census_host is not None:
                        from .line_detection import Mask3, parse_census

                        census = parse_census(census_host, int(line_mask.shape[0]))
                        if gpool is None:
                            line_mask = line_mask.view(Mask3)
                            line_mask.census = census
                    geo = gpool.submit(line_mask, merge_lines, census).result() if gpool is not None else None
                    if geo is not None and geo[0] == "failed":
                        return OpStatus.FAILED, geo[1]
                    if geo is not None and geo[0] == "ok":
                        # stages 2-4 with the geometry done by a worker process (no rotation, no dewarping)
                        import cv2

                        from .data import BBox
                        from .line_detection import extract_line_images

                        _, page_angle, lines, rois, masks, hs = geo
                        rot_mask = cv2.cvtColor(np.ascontiguousarray(line_mask), cv2.COLOR_GRAY2RGB)
                        sorted_lines = [Line(g, c, BBox(*bb), ctr) for g, c, bb, ctr in lines]
                        line_images = extract_line_images(image, sorted_lines, k_factor, bbox_tolerance,
                                                          page_dev=pages_dev[i], download=False,
                                                          _rois=(rois, masks, hs))
                        if not line_images:
                            return OpStatus.FAILED, "No valid line images extracted"
                    else:       # threads only, or a page that has to be rotated first (GPU work)
                        status, res = self.build_lines(image, line_mask)
                        if status == OpStatus.FAILED:
                            return status, res
                        rot_img, rot_mask, _, filtered, page_angle = res
                        status, dew = self.apply_dewarping(rot_img, rot_mask, filtered, page_angle, use_tps=use_tps,
                                                           tps_threshold=tps_threshold)
                        if status == OpStatus.FAILED:
                            return status, dew
                        status, res = self.extract_lines(dew.work_img, rot_mask, dew.filtered_contours,
                                                         merge_lines=merge_lines, k_factor=k_factor,
                                                         bbox_tolerance=bbox_tolerance)
                        if status == OpStatus.FAILED:
                            return status, res
                        sorted_lines, line_images = res
                    keep_idx = [j for j, im in enumerate(line_images) if im.shape[0] != 0 and im.shape[1] != 0]
                    pack = getattr(line_images, "device_pack", None)
                    if pack is None or list(pack[3]) != keep_idx or len(line_images) != len(sorted_lines):
                        return None       # crops not (all) on the device: this page takes the page-by-page route
                    for j in keep_idx:
                        self.ocr_inference._check_fit(line_images[j], self.use_line_prepadding)
                    done = torch.cuda.Event()
                    done.record(tls.stream)
                    if getattr(rot_mask, "gray", None) is not None:
                        rot_mask.gray = None      # the 1-channel mask is a pinned staging buffer of this object
                    return OpStatus.SUCCESS, (rot_mask, sorted_lines, [sorted_lines[j] for j in keep_idx], pack, done,
                                              float(page_angle))
            except Exception:  # noqa: BLE001  (run_ocr reproduces the failure with the reference's message)
                return None
            finally:
                tls.fast = False
                det._tls.last = None

        # ---- segmentation: groups of pages per forward call --------------------------------------
        pool = self._thread_pool(workers)
        futures = {}
        sizes, left = [], len(good)
        while left > 0:
            # a small first group starts the GPU after two uploads; a chunk queued behind another one needs none
            sizes.append(min(left, 2 if (not sizes and not (_defer and _slot)) else 6))
            left -= sizes[-1]
        g0 = 0
        thr = det._default_threshold
        for group in sizes:
            idx = good[g0:g0 + group]
            g0 += group
            try:
                for i in idx:
                    if not up_ready[i].wait(timeout=300.0):
                        raise TimeoutError("page upload did not finish")
                    if i not in up_ev:
                        raise RuntimeError("page upload failed")
                    compute.wait_event(up_ev[i])
                want_census = layout_ch is None and getattr(det, "with_census", False)
                r_ = det._predict_device_batch([pages_dev[i] for i in idx], thr, census=want_census)
                masks, cens = r_ if want_census else (r_, [None] * len(r_))
                seg_done = torch.cuda.Event()
                seg_done.record(compute)
                copy_out.wait_event(seg_done)
                with torch.cuda.stream(copy_out):
                    for i, m, c in zip(idx, masks, cens):
                        key = ("mask_host", _slot, i, tuple(m.shape))
                        h = det._host_bufs.get(key)
                        if h is None:
                            h = det._host_bufs[key] = torch.empty(m.shape, dtype=torch.uint8, pin_memory=True)
                        h.copy_(m, non_blocking=True)
                        m.record_stream(copy_out)
                        ch = None
                        if c is not None:
                            ckey = ("census_host", _slot, i, int(c.numel()))
                            ch = det._host_bufs.get(ckey)
                            if ch is None:
                                ch = det._host_bufs[ckey] = torch.empty(c.shape, dtype=torch.uint8, pin_memory=True)
                            ch.copy_(c, non_blocking=True)
                            c.record_stream(copy_out)
                        ev = torch.cuda.Event(blocking=BLOCKING_EVENTS)     # host threads wait on it: sleep, do not spin
                        ev.record(copy_out)
                        futures[i] = pool.submit(page_stage, i, h.numpy(), ev, ch.numpy() if ch is not None else None)
            except Exception:  # noqa: BLE001
                for i in idx:
                    futures.pop(i, None)
                    results[i] = "retry"
        up_thread.join()

        # ---- recognition: the crops of all pages as one batch -------------------------------------
        def finish():
            staged_pages = []
            for i in good:
                if i not in futures:
                    continue
                r = futures[i].result()
                if r is None:
                    results[i] = "retry"
                elif r[0] == OpStatus.FAILED:
                    results[i] = r
                else:
                    staged_pages.append((i, r[1]))
            # on a high-priority stream of its own: the next chunk's segmentation may already be queued on the compute
            # stream, and the text of this chunk must not wait behind it
            rec = getattr(self, "_rec_stream", None)
            if rec is None:
                rec = self._rec_stream = torch.cuda.Stream(dev, priority=-1)
            try:
                with torch.cuda.stream(rec):
                    texts = self._recognise_packs([(p[3], p[4]) for _, p in staged_pages])
                for (i, (rot_mask, sorted_lines, kept, _, _, angle)), tx in zip(staged_pages, texts):
                    results[i] = (OpStatus.SUCCESS, [rot_mask, sorted_lines, self._make_ocr_lines(kept, tx, target_encoding),
                                                     angle])
            except Exception:  # noqa: BLE001
                for i, _ in staged_pages:
                    results[i] = "retry"
            # the recognition waited for every page's extraction (their `done` events) and those for the masks, i.e. for
            # the chunk's segmentation and uploads: once it has drained, no stream reads this chunk's pages or crops
            rec.synchronize()
            if any(isinstance(r, str) for r in results):
                torch.cuda.synchronize(dev)   # a failed page may have left kernels behind on its thread's stream
            pages_dev.clear()
            for i in range(n):
                if isinstance(results[i], str):
                    results[i] = self.run_ocr(images[i], k_factor, bbox_tolerance, merge_lines, use_tps, tps_threshold,
                                              target_encoding, eval_mode)
            return results

        return finish if _defer else finish()

    def _recognise_packs(self, packs, max_lines: int = 512) -> list[list[str]]:
        """packs: [((crops_dev, hw, off, real), done_event)] per page -> texts per page.  a-7 runs per page into
        one shared line batch; a-8 / a-9 run over up to `max_lines` lines per call."""
        eng, ocr = self.engine, self.ocr_inference
        counts = [len(p[0][3]) for p in packs]
        total = sum(counts)
        if total == 0:
            return [[] for _ in packs]
        from .engine import OCR_CPAD

        stream = torch.cuda.current_stream(eng.device)
        lines = torch.empty((total, ocr._input_width, OCR_CPAD), dtype=torch.bfloat16, device=eng.device)
        o = 0
        for (pack, done), c in zip(packs, counts):
            if c:
                stream.wait_event(done)
                eng.ocr_prepare_lines(pack[0], pack[1], pack[2], c, ocr._input_height, ocr._input_width,
                                      self.use_line_prepadding, out=lines[o:o + c])
                o += c
        seqs = []
        for b0 in range(0, total, max_lines):
            logits = eng.ocr_forward(ocr.ocr_session, lines[b0:b0 + max_lines])
            ids, lens = eng.ctc_greedy(logits, layout=LAYOUT_TV, blank=ocr.ctc_decoder.blank_index)
            ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
            seqs.extend(ids[j, : lens[j]].tolist() for j in range(len(lens)))
        texts, o = [], 0
        for c in counts:
            texts.append([ocr.ctc_decoder.ids_to_text(sq) for sq in seqs[o:o + c]])
            o += c
        return texts


class ModernBookFormatLayoutDetection:
    """BDRC/inference.py:860-961 (torch path, config C4): float-bilinear resize_clamp, pad 255,
    tile, /255, DeepLabV3+ forward, softmax, stitch, crop, bilinear resize back, > thr, per-class
    union bounding box written as {output_dir}/{image_name}.json.  config.checkpoint (or
    config.onnx_file) names a BDRCW1 container of the same network."""

    def __init__(self, config: LayoutDetectionConfig, engine: Engine | None = None):
        self.config = config
        self.classes = config.classes
        self.engine = engine or get_engine()
        self.device = "cuda"
        self.model = self.engine.load_model(SEG, _read_weights(self._weights_source(config)), len(self.classes))

    @staticmethod
    def _weights_source(config: LayoutDetectionConfig):
        """The reference's torch path loads config.checkpoint (BDRC/inference.py:864).  Here: the BDRCW1 container
        named by onnx_file (after utils._weights_path's `.bdrcw` substitution) when there is one, else the
        checkpoint -- a BDRCW1 container, or the reference's own `.pth` (BatchNorm folded at load)."""
        def is_blob(x):
            if isinstance(x, (dict, bytes, bytearray)):
                return True
            try:
                with open(x, "rb") as f:
                    return f.read(8) == MAGIC
            except (OSError, TypeError):
                return False

        for cand in (config.onnx_file, config.checkpoint):
            if cand is not None and not (isinstance(cand, str) and cand == "") and is_blob(cand):
                return cand
        for cand in (config.checkpoint, config.onnx_file):
            if isinstance(cand, str) and cand and os.path.exists(cand):
                return cand
        raise FileNotFoundError(f"no weights found: checkpoint={config.checkpoint!r}, onnx_file={config.onnx_file!r}")

    def predict_masks(self, image: np.ndarray, class_threshold: float = 0.8) -> np.ndarray:
        """RGB HWC u8 page -> u8 {0,255} masks [C, H, W] (what run() thresholds, :949-956)."""
        eng = self.engine
        page = torch.from_numpy(np.ascontiguousarray(image[..., :3], dtype=np.uint8)).to(eng.device)
        tiles, meta = eng.seg_tile_rgb(page)                                   # multi_image_collate_fn
        logits = eng.seg_forward(self.model, tiles)
        return eng.seg_postprocess_torch(logits, meta, class_threshold).cpu().numpy()

    def predict_masks_batch(self, images: Sequence[np.ndarray], class_threshold: float = 0.8,
                            max_tiles: int = 148, copy: bool = True) -> list[np.ndarray]:
        """[predict_masks(i) for i in images] with the tiles of several images sharing one forward call -- what
        the reference's DataLoader + multi_image_collate_fn (utils.py:1142-1172, batch_size images per step) does.
        Uploads run ahead on a copy stream and masks return through pinned buffers on a third one; images are
        grouped while their tiles fit `max_tiles` (one 148-tile chunk of the network).  copy=False returns views of
        this object's pinned staging buffers (valid until the next call) instead of fresh arrays."""
        from .engine import clamp_size_torch

        eng, dev = self.engine, self.engine.device
        compute = torch.cuda.current_stream(dev)
        copy_in = getattr(self, "_copy_in", None) or torch.cuda.Stream(dev)
        copy_out = getattr(self, "_copy_out", None) or torch.cuda.Stream(dev)
        self._copy_in, self._copy_out = copy_in, copy_out
        if not hasattr(self, "_host_bufs"):
            self._host_bufs = {}
        n_tiles = []
        for img in images:
            Hc, Wc = clamp_size_torch(int(img.shape[0]), int(img.shape[1]), 512)
            n_tiles.append((-(-Hc // 512)) * (-(-Wc // 512)))
        ups, evs = [], []
        copy_in.wait_stream(compute)
        ring = getattr(self, "_up_slots", None)
        if ring is None:
            ring = self._up_slots = [[None, None] for _ in range(4)]
        with torch.cuda.stream(copy_in):
            for k, img in enumerate(images):
                d, ev = eng.upload_page(np.ascontiguousarray(img[..., :3], dtype=np.uint8), ring, k)
                ups.append(d)
                evs.append(ev)
        outs: list = [None] * len(images)
        i = 0
        while i < len(images):
            j, tot = i, 0
            while j < len(images) and (j == i or tot + n_tiles[j] <= max_tiles):
                tot += n_tiles[j]
                j += 1
            for k in range(i, j):
                compute.wait_event(evs[k])
                ups[k].record_stream(compute)
            tiles = torch.empty((tot, 512, 512, 3), dtype=torch.float32, device=dev)
            metas, o = [], 0
            for k in range(i, j):
                _, meta = eng.seg_tile_rgb(ups[k], out=tiles[o:o + n_tiles[k]])
                metas.append(meta)
                o += n_tiles[k]
            logits = eng.seg_forward(self.model, tiles)
            o, masks = 0, []
            for k, meta in zip(range(i, j), metas):
                masks.append(eng.seg_postprocess_torch(logits[o:o + n_tiles[k]], meta, class_threshold))
                o += n_tiles[k]
            done = torch.cuda.Event()
            done.record(compute)
            copy_out.wait_event(done)
            with torch.cuda.stream(copy_out):
                for k, m in zip(range(i, j), masks):
                    key = (k, tuple(m.shape))
                    h = self._host_bufs.get(key)
                    if h is None:
                        h = self._host_bufs[key] = torch.empty(m.shape, dtype=torch.uint8, pin_memory=True)
                    h.copy_(m, non_blocking=True)
                    m.record_stream(copy_out)
                    outs[k] = h
            i = j
        copy_out.synchronize()
        compute.synchronize()
        return [o.numpy().copy() if copy else o.numpy() for o in outs]

    def post_process_sample(self, prediction: np.ndarray, class_index: int):
        import cv2

        contours, _ = cv2.findContours(prediction[class_index], cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
        if len(contours) == 0:
            return None
        x, y, w, h = cv2.boundingRect(np.vstack(contours))                     # utils.get_union_bbox :1081-1097
        return {"bbox": {"x": x, "y": y, "w": w, "h": h}}

    def extract_json_data(self, meta: dict, prediction: np.ndarray, filter_classes, output_dir: str):
        import json

        found = {}
        for idx, name in enumerate(self.classes):
            if filter_classes and name not in filter_classes:
                continue
            bbox = self.post_process_sample(prediction, idx)
            if bbox is not None:
                found[name] = bbox
        stem = os.path.splitext(os.path.basename(meta["image_name"]))[0]
        with open(f"{output_dir}/{stem}.json", "w", encoding="UTF-8") as f:
            json.dump(found, f, ensure_ascii=False, indent=1)
        return found

    def run(self, directory: str, output_dir: str, filter_classes=None, batch_size: int = 4, num_workers: int = 4,
            class_threshold: float = 0.8):
        if filter_classes:
            for c in filter_classes:
                if c not in self.classes:
                    raise ValueError(f"ERROR: provided filter classes: {filter_classes} are not part of the model's classes!")
        os.makedirs(output_dir, exist_ok=True)
        from .dataio import PageLoader

        # the files are decoded ahead by `num_workers` host threads (the reference's DataLoader workers, :935-941);
        # `batch_size` pages share one forward call
        loader = PageLoader(directory, workers=max(1, int(num_workers)), prefetch=2 * max(1, int(batch_size)))
        for imgs, metas in loader.batches(max(1, int(batch_size))):
            for meta, masks in zip(metas, self.predict_masks_batch(imgs, class_threshold, copy=False)):
                self.extract_json_data(meta, masks, filter_classes, output_dir)
