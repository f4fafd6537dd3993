"""Seeded random-init weights of the two reference architectures and their algorithmic FLOP counts.

Shared by the GPU path (bench.py, tests) and the oracle (oracle/nets.py imports these), so both
provably run the same synthetic model.  Pure numpy; nothing here computes a forward pass.
Initializer names/shapes: PyTorch-LineInference.ipynb:L131-L230 (segmentation) and SURVEY.md 8(a)-8
(Easter2).
"""
from __future__ import annotations

import os

import numpy as np

R34_LAYERS = [(64, 3, 1, 1), (128, 4, 2, 1), (256, 6, 2, 1), (512, 3, 1, 2)]  # (ch, blocks, stride, dilation)


# ------------------------------------------------------------------------------------------
# weight generators (seeded, deterministic)
# ------------------------------------------------------------------------------------------


def _conv_w(rng, cout, cin, kh, kw, gain=1.0):
    fan_in = cin * kh * kw
    return (rng.standard_normal((cout, cin, kh, kw)) * (gain * np.sqrt(2.0 / fan_in))).astype(np.float32)


def _bias(rng, c, scale=0.05):
    return (rng.standard_normal(c) * scale).astype(np.float32)


def seg_weights(seed: int = 11, n_classes: int = 1, in_ch: int = 3) -> dict[str, np.ndarray]:
    """Random-init DeepLabV3+/R34 weights with the initializer names/shapes of
    PyTorch-LineInference.ipynb:L131-L230 ('.weight_bias' shortened to '.bias')."""
    rng = np.random.default_rng(seed)
    w: dict[str, np.ndarray] = {}

    def conv(name, cout, cin, k, gain=1.0, bias=True):
        w[name + ".weight"] = _conv_w(rng, cout, cin, k, k, gain)
        if bias:
            w[name + ".bias"] = _bias(rng, cout)

    conv("encoder.conv1", 64, in_ch, 7)
    cin = 64
    for li, (ch, blocks, stride, _dil) in enumerate(R34_LAYERS, start=1):
        for b in range(blocks):
            p = f"encoder.layer{li}.{b}"
            conv(p + ".conv1", ch, cin if b == 0 else ch, 3)
            conv(p + ".conv2", ch, ch, 3, gain=0.25)  # keeps the residual stream O(1)
            if b == 0 and (cin != ch or stride != 1):
                conv(p + ".downsample.0", ch, cin, 1, gain=0.7)
        cin = ch
    conv("decoder.aspp.0.convs.0.0", 256, 512, 1)
    for i in (1, 2, 3):
        w[f"decoder.aspp.0.convs.{i}.0.0.weight"] = _conv_w(rng, 512, 1, 3, 3, gain=1.5)
        conv(f"decoder.aspp.0.convs.{i}.0.1", 256, 512, 1)
    conv("decoder.aspp.0.convs.4.1", 256, 512, 1)
    conv("decoder.aspp.0.project.0", 256, 1280, 1)
    w["decoder.aspp.1.0.weight"] = _conv_w(rng, 256, 1, 3, 3, gain=1.5)
    conv("decoder.aspp.1.1", 256, 256, 1)
    conv("decoder.block1.0", 48, 64, 1)
    w["decoder.block2.0.0.weight"] = _conv_w(rng, 304, 1, 3, 3, gain=1.5)
    conv("decoder.block2.0.1", 256, 304, 1)
    conv("segmentation_head.0", n_classes, 256, 1, gain=2.0)
    return w


def easter2_weights(seed: int = 7, vocab: int = 77, in_h: int = 80) -> dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    w: dict[str, np.ndarray] = {}

    def conv(name, cout, cin, k, gain=1.0):
        w[name + ".weight"] = _conv_w(rng, cout, cin, 1, k, gain)[:, :, 0, :]  # [O, I, K]
        w[name + ".bias"] = _bias(rng, cout)

    conv("conv1", 128, in_h, 3)
    conv("conv2", 128, 128, 3)
    cin = 128
    for u, k in zip((1, 2, 3), (5, 7, 9)):
        p = f"unit{u}"
        conv(p + ".res_old", 256, cin, 1, gain=0.5)
        conv(p + ".res_data", 256, cin, 1, gain=0.5)
        conv(p + ".conv_a", 256, cin, k)
        conv(p + ".conv_b", 256, 256, k)
        conv(p + ".conv_c", 256, 256, k, gain=0.7)
        w[p + ".se_fc1.weight"] = (rng.standard_normal((32, 256)) * np.sqrt(2.0 / 256)).astype(np.float32)
        w[p + ".se_fc1.bias"] = _bias(rng, 32)
        w[p + ".se_fc2.weight"] = (rng.standard_normal((256, 32)) * np.sqrt(1.0 / 32)).astype(np.float32)
        w[p + ".se_fc2.bias"] = _bias(rng, 256)
        cin = 256
    conv("conv3", 512, 256, 11)
    conv("conv4", 512, 512, 1)
    conv("head", vocab, 512, 1, gain=3.0)
    return w



FITTED_HEADS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "fitted_heads.npz")


def _fitted(prefix: str, path: str | None = None) -> dict[str, np.ndarray]:
    z = np.load(path or FITTED_HEADS)
    return {k[len(prefix):]: z[k].astype(np.float32) for k in z.files if k.startswith(prefix)}


def seg_weights_fitted(seed: int = 11, path: str | None = None) -> dict[str, np.ndarray]:
    """seg_weights(seed, 1) with the whole DeepLabV3+ decoder (ASPP, block1, block2, head: 1.15 M parameters)
    replaced by the tensors tests/golden/train_standins.py trained against the thinned text-band mask of
    synthetic text / stroke-band pages (the ResNet-34 body keeps the seeded random init): the network then emits
    one clean blob per text line, like a trained line model, so the stages after it see realistic work."""
    assert seed == 11, "the committed tensors belong to seed 11"
    w = seg_weights(seed, 1, 3)
    w.update(_fitted("seg.", path))
    return w


def easter2_weights_fitted(path: str | None = None) -> dict[str, np.ndarray]:
    """Easter2 (V = 77, H = 80) trained with CTC by tests/golden/train_standins.py on the line crops the pipeline's
    own extraction cuts out of synthetic.text_page() pages plus synthetic.glyph_line() crops; values are the
    committed fp16 tensors widened to fp32.  It reads its pages (316 / 316 held-out page lines equal to the
    generator's transcription, tests/golden/fitted_heads_report.json), which is what the >= 95 % identical-
    decoded-lines parity target presupposes."""
    w = _fitted("ocr.", path)
    if "conv1.weight" not in w:           # fixture holds only the output layers: rest is the seeded init
        base = easter2_weights(7, 77, 80)
        base.update(w)
        w = base
    return w


def seg_flops_per_tile(n_classes: int = 1) -> float:
    """Algorithmic FLOPs of one 512x512 tile (dense conv MACs x2, reference 3-channel stem)."""
    f = 2 * 256 * 256 * 64 * 3 * 49
    cin = 64
    size = 128
    for ch, blocks, stride, _ in R34_LAYERS:
        for b in range(blocks):
            s = stride if b == 0 else 1
            size_out = size // s
            f += 2 * size_out * size_out * ch * (cin if b == 0 else ch) * 9
            f += 2 * size_out * size_out * ch * ch * 9
            if b == 0 and (cin != ch or stride != 1):
                f += 2 * size_out * size_out * ch * cin
            size = size_out
        cin = ch
    f += 2 * 32 * 32 * 256 * 512 * 4 + 2 * 256 * 512          # ASPP 1x1s + pooled branch
    f += 2 * 32 * 32 * 512 * 9 * 3                             # ASPP depthwise
    f += 2 * 32 * 32 * 256 * 1280                              # project
    f += 2 * 32 * 32 * 256 * 9 + 2 * 32 * 32 * 256 * 256       # aspp.1
    f += 2 * 128 * 128 * 48 * 64                               # block1
    f += 2 * 128 * 128 * 304 * 9 + 2 * 128 * 128 * 256 * 304   # block2
    f += 2 * 128 * 128 * n_classes * 256                       # head
    return float(f)


def easter2_flops_per_line(vocab: int = 77, in_h: int = 80, W: int = 2000) -> float:
    T2, T = W // 2, W // 4
    f = 2 * T2 * 128 * in_h * 3 + 2 * T * 128 * 128 * 3
    cin = 128
    for k in (5, 7, 9):
        f += 2 * T * 256 * cin * 2 + 2 * T * 256 * cin * k + 2 * T * 256 * 256 * k * 2
        cin = 256
    f += 2 * T * 512 * 256 * 11 + 2 * T * 512 * 512 + 2 * T * vocab * 512
    return float(f)
