"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (ocr_inference_b200/).

Import shim that lets the UNMODIFIED reference (/root/reference/BDRC/*.py) run in the build
container, where onnxruntime, pyctcdecode, pyewts, tps, segmentation_models_pytorch,
matplotlib and botok are absent.  It is used for two things only:
  * tests/golden/make_golden.py generates the committed golden vectors with it;
  * tests (marked `needs_reference`, skipped when /root/reference is absent, e.g. on the GPU
    box) pin oracle/*.py against the real reference functions.

Stand-ins injected (everything else is the reference's own code):
  onnxruntime.InferenceSession  -> FakeSession wrapping a python callable (numpy in/out),
                                   honouring run_with_ort_values / OrtValue (inference.py:154-158)
  pyctcdecode.build_ctcdecoder  -> greedy best path (north_star semantics; the reference's beam
                                   search lives in an unpinned fork, requirements.txt:13)
  pyewts.pyewts                 -> identity converter
  tps.ThinPlateSpline           -> oracle.tps_numpy.ThinPlateSpline restatement
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("BDRC_REFERENCE_ROOT", "/root/reference")

# model_file -> callable(np.ndarray) -> np.ndarray, looked up by FakeSession
SESSION_REGISTRY: dict[str, object] = {}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "BDRC", "inference.py"))


def _stub(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    sys.modules[name] = m
    return m


class _OrtValue:
    def __init__(self, arr):
        self._arr = arr

    @staticmethod
    def ortvalue_from_numpy(arr):
        return _OrtValue(np.asarray(arr))

    def numpy(self):
        return self._arr


class FakeSession:
    """Stands in for ort.InferenceSession(model_file[, providers=...])."""

    def __init__(self, model_file, providers=None, **_):
        self.model_file = model_file
        self.fn = SESSION_REGISTRY[str(model_file)]
        self.calls = []

    def run_with_ort_values(self, out_names, feeds):
        (x,) = [v.numpy() for v in feeds.values()]
        self.calls.append(tuple(x.shape))
        y = np.asarray(self.fn(x))
        return [_OrtValue(y)]


class _GreedyDecoder:
    """Greedy best path: argmax per step, collapse repeats, drop the blank token."""

    def __init__(self, vocab):
        self.vocab = list(vocab)

    def decode(self, logits):
        ids = np.argmax(np.asarray(logits), axis=-1)
        out, prev = [], -1
        for i in ids.tolist():
            if i != prev and self.vocab[i] != "<blk>":
                out.append(self.vocab[i])
            prev = i
        return "".join(out)

    def decode_beams(self, logits):
        raise NotImplementedError("beam search is out of scope (SURVEY.md 8c)")


class _Pyewts:
    def toUnicode(self, s):
        return s

    def toWylie(self, s):
        return s


_installed = False


def install():
    """Register the stub modules and put the reference on sys.path. Idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    ort = _stub("onnxruntime")
    ort.InferenceSession = FakeSession
    ort.OrtValue = _OrtValue
    ort.get_available_providers = lambda: ["CPUExecutionProvider"]
    ort.get_device = lambda: "CPU"
    _stub("segmentation_models_pytorch")
    ew = _stub("pyewts")
    ew.pyewts = _Pyewts
    pc = _stub("pyctcdecode")
    pc.build_ctcdecoder = lambda vocab, **kw: _GreedyDecoder(vocab)
    pcd = _stub("pyctcdecode.decoder")
    pcd.OutputBeam = object
    pc.decoder = pcd
    tps = _stub("tps")
    from oracle.tps_numpy import ThinPlateSpline

    tps.ThinPlateSpline = ThinPlateSpline
    bt = _stub("botok")
    bt.normalize_unicode = lambda s, *a, **k: s
    bt.tokenize_in_stacks = lambda s, *a, **k: list(s)
    for name in ("evaluate", "natsort", "huggingface_hub", "datasets"):
        try:
            importlib.import_module(name)
        except Exception:  # noqa: BLE001
            m = _stub(name)
            if name == "natsort":
                m.natsorted = sorted
            if name == "huggingface_hub":
                m.snapshot_download = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no network"))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def import_reference():
    """Returns the reference modules (utils, inference, line_detection, data, image_dewarping)."""
    install()
    mods = {}
    for name in ("data", "utils", "line_detection", "image_dewarping", "inference"):
        mods[name] = importlib.import_module(f"BDRC.{name}")
    return types.SimpleNamespace(**mods)
