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
