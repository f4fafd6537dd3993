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
