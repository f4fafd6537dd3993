"""BDRCW1 weight container: what bdrc_load_weights() consumes in place of the ONNX / .pth files the
reference hands to onnxruntime / torch.load (BDRC/inference.py:128-130,294; BDRC/utils.py:1175-1180).

Layout (little endian):
    magic   8 bytes  b"BDRCW1\\0\\0"
    u32     n_tensors
    per tensor: u16 name_len, name (utf-8), u8 ndim, u32 dims[ndim], u64 offset, u64 nbytes
    padding to a 64-byte boundary, then the fp32 tensor data (offsets relative to that boundary)

Tensor names are the ONNX initializer names of the exported graph with BatchNorm folded
(PyTorch-LineInference.ipynb:L131-L230; '<conv>.weight_bias' is spelled '<conv>.bias').
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"BDRCW1\0\0"


def write_blob(tensors: dict[str, np.ndarray]) -> bytes:
    head = [MAGIC, struct.pack("<I", len(tensors))]
    off = 0
    datas = []
    for name, arr in tensors.items():
        a = np.ascontiguousarray(arr, dtype="<f4")
        nb = name.encode("utf-8")
        head.append(struct.pack("<H", len(nb)) + nb + struct.pack("<B", a.ndim) +
                    struct.pack(f"<{a.ndim}I", *a.shape) + struct.pack("<QQ", off, a.nbytes))
        datas.append(a.tobytes())
        pad = (-a.nbytes) % 64
        if pad:
            datas.append(b"\0" * pad)
        off += a.nbytes + pad
    h = b"".join(head)
    h += b"\0" * ((-len(h)) % 64)
    return h + b"".join(datas)


def read_blob(blob: bytes) -> dict[str, np.ndarray]:
    if blob[:8] != MAGIC:
        raise ValueError("not a BDRCW1 container")
    (n,) = struct.unpack_from("<I", blob, 8)
    off = 12
    ents = []
    for _ in range(n):
        (nl,) = struct.unpack_from("<H", blob, off)
        off += 2
        name = blob[off:off + nl].decode("utf-8")
        off += nl
        nd = blob[off]
        off += 1
        dims = struct.unpack_from(f"<{nd}I", blob, off)
        off += 4 * nd
        o, nb = struct.unpack_from("<QQ", blob, off)
        off += 16
        ents.append((name, dims, o, nb))
    data0 = (off + 63) // 64 * 64
    out = {}
    for name, dims, o, nb in ents:
        out[name] = np.frombuffer(blob, dtype="<f4", count=nb // 4, offset=data0 + o).reshape(dims).copy()
    return out


def from_torch_state_dict(sd) -> dict[str, np.ndarray]:
    """Convenience for users holding a BN-folded torch state dict (keys already in ONNX naming)."""
    return {k: v.detach().cpu().float().numpy() for k, v in sd.items()}


def fold_batchnorm_state_dict(sd, eps: float = 1e-5) -> dict[str, np.ndarray]:
    """An UN-folded torch state dict (what the reference's torch path loads: BDRC/utils.py:1175-1180,
    `sm.DeepLabV3Plus(...).load_state_dict(checkpoint["state_dict"])`) -> the BN-folded tensors the B200 loader
    expects, named like the exported ONNX initializers (PyTorch-LineInference.ipynb:L131-L230).

    A state dict lists tensors in module-registration order, so the convolution a BatchNorm normalises is the
    last 4-D `.weight` seen before the BatchNorm's `running_mean`: w' = w * g / sqrt(var + eps) per output
    channel, b' = beta + (b - mean) * g / sqrt(var + eps).  Depthwise convs of the separable blocks keep their
    names and no bias (the BatchNorm that follows the pointwise conv folds into the pointwise conv)."""
    keys = list(sd.keys())
    prefixes = [k[: -len(".running_mean")] for k in keys if k.endswith(".running_mean")]
    bn_keys = {p + sfx for p in prefixes for sfx in (".weight", ".bias", ".running_mean", ".running_var",
                                                     ".num_batches_tracked")}
    get = lambda k: np.asarray(sd[k].detach().cpu().float().numpy() if hasattr(sd[k], "detach") else sd[k],  # noqa: E731
                               dtype=np.float32)
    out: dict[str, np.ndarray] = {}
    last_conv = None
    for k in keys:
        if k in bn_keys:
            if k.endswith(".running_mean"):
                p = k[: -len(".running_mean")]
                if last_conv is None:
                    raise ValueError(f"BatchNorm '{p}' has no convolution before it")
                g, beta = get(p + ".weight"), get(p + ".bias")
                mean, var = get(p + ".running_mean"), get(p + ".running_var")
                scale = g / np.sqrt(var + eps)
                w = out[last_conv + ".weight"]
                b = out.get(last_conv + ".bias", np.zeros(w.shape[0], np.float32))
                out[last_conv + ".weight"] = (w * scale.reshape(-1, *([1] * (w.ndim - 1)))).astype(np.float32)
                out[last_conv + ".bias"] = (beta + (b - mean) * scale).astype(np.float32)
                last_conv = None
            continue
        v = get(k)
        out[k] = v
        if k.endswith(".weight") and v.ndim >= 3:
            last_conv = k[: -len(".weight")]
    return out


def load_reference_checkpoint(path: str) -> dict[str, np.ndarray]:
    """A `.pth` written by the reference's training code -> folded tensors.  Accepts a bare state dict or the
    {"state_dict": ...} wrapper of BDRC/utils.py:1176-1179, with an optional "model." / "module." key prefix."""
    import torch

    ck = torch.load(path, map_location="cpu", weights_only=True)
    sd = ck["state_dict"] if isinstance(ck, dict) and "state_dict" in ck else ck
    for pre in ("model.", "module."):
        if sd and all(k.startswith(pre) for k in sd):
            sd = {k[len(pre):]: v for k, v in sd.items()}
    return fold_batchnorm_state_dict(sd)
