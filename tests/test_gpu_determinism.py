"""Run-to-run bit reproducibility of both forward passes on the GPU.

Every kernel on the path sums in a fixed order, so two runs on the same input must agree bit for bit;
a difference means a data race (this test caught a TMA write overtaking pending shared-memory loads in
the depthwise row-ring kernel, which the tolerance-based parity tests let through)."""
import pytest
import torch

from oracle import nets

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_tiles", [4, 48])
def test_seg_forward_bit_reproducible(engine, n_tiles):
    model = engine.load_model(1, nets.seg_weights(11, 1, 3), 1)
    g = torch.Generator(device="cuda").manual_seed(n_tiles)
    x = (torch.rand((n_tiles, 512, 512), device="cuda", generator=g) > 0.2).to(torch.uint8)
    ref = engine.seg_forward(model, x, lowres=True).clone()
    for _ in range(6):
        assert torch.equal(engine.seg_forward(model, x, lowres=True), ref)
    full = engine.seg_forward(model, x[:2]).clone()
    assert torch.equal(engine.seg_forward(model, x[:2]), full)


def test_ocr_forward_bit_reproducible(engine):
    model = engine.load_model(2, nets.easter2_weights(7, 77, 80), 77)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.sign(torch.randn((9, 2000, 128), device="cuda", generator=g)).to(torch.bfloat16)
    x[:, :, 80:] = 0
    ref = engine.ocr_forward(model, x).clone()
    for _ in range(6):
        assert torch.equal(engine.ocr_forward(model, x), ref)


@pytest.mark.parametrize("shape", [(48, 128, 128, 304), (48, 32, 32, 256), (5, 64, 64, 64)])
def test_depthwise_ring_kernel_in_isolation(engine, shape):
    """The TMA row-ring depthwise kernel alone, on random data: equal to the register-strip kernel up to the
    fp32 summation order, and bit-identical from run to run (the slot-refill race showed up here 20/20 times)."""
    import ctypes as C
    import os

    import numpy as np

    N, H, W, Cn = shape
    g = torch.Generator(device="cpu").manual_seed(H + Cn)
    x = torch.randn((N, H, W, Cn), generator=g).to(torch.bfloat16).cuda()
    w = np.ascontiguousarray(np.random.default_rng(Cn).standard_normal((9, Cn)).astype(np.float32))

    def run():
        y = torch.empty_like(x)
        rc = engine.lib.bdrc_depthwise3x3_test(engine.ctx.h, x.data_ptr(), N, H, W, Cn, 1, w.ctypes.data, y.data_ptr(),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream))
        engine.ctx.check(rc, "depthwise3x3_test")
        return y

    os.environ["BDRC_NO_DW_RING"] = "1"
    try:
        ref = run()
    finally:
        del os.environ["BDRC_NO_DW_RING"]
    first = run()
    assert float((first.float() - ref.float()).abs().max()) <= 0.02 * float(ref.float().abs().max())
    for _ in range(10):
        assert torch.equal(run(), first)


def test_batching_is_exact(engine):
    """Tiles of different pages and lines of different pages share batches (SURVEY 8a note: every reduction is per
    batch element).  The results of an element must not depend on what else is in the batch, bit for bit --
    whatever tile shape, accumulator interleave or CTA pairing the batch size selects."""
    seg = engine.load_model(1, nets.seg_weights(11, 1, 3), 1)
    g = torch.Generator(device="cuda").manual_seed(77)
    x = (torch.rand((10, 512, 512), device="cuda", generator=g) > 0.2).to(torch.uint8)
    full = engine.seg_forward(seg, x, lowres=True).clone()
    for n in (1, 3, 4):
        assert torch.equal(engine.seg_forward(seg, x[:n], lowres=True), full[:n]), n
    ocr = engine.load_model(2, nets.easter2_weights(7, 77, 80), 77)
    xl = torch.sign(torch.randn((70, 2000, 128), device="cuda", generator=g)).to(torch.bfloat16)
    xl[:, :, 80:] = 0
    big = engine.ocr_forward(ocr, xl).clone()
    for n in (1, 6, 64):
        assert torch.equal(engine.ocr_forward(ocr, xl[:n]), big[:n]), n
