"""GPU parity: the two network forward passes (bf16 operands, fp32 accumulate on tcgen05) vs the
torch-CPU fp32 oracle with the same seeded weights.

Tolerance (stated, as north_star asks): bf16 keeps 8 mantissa bits (relative rounding error 2^-9)
and both weights and every layer output are rounded once, so ~35-50 roundings accumulate to a few
per cent of the logit scale.  We require
    rms(delta) <= 0.03 * std(oracle logits)                      (aggregate error)
    |delta|    <= 0.08 * std(oracle) + 0.05 * |oracle|           for >= 99.9 % of the elements
and that the decisions (mask bit / argmax) agree wherever the oracle's margin to the decision
boundary exceeds that per-element tolerance.
"""
import numpy as np
import pytest
import torch

from oracle import hostops as ho
from oracle import nets, synth

pytestmark = pytest.mark.gpu


def _tol_check(got, ref, frac=0.999):
    err = np.abs(got - ref)
    std = float(ref.std())
    rms = float(np.sqrt((err.astype(np.float64) ** 2).mean()))
    tol = 0.08 * std + 0.05 * np.abs(ref)
    ok = err <= tol
    print(f"parity: rms_err/std={rms / std:.4f} max_err/std={err.max() / std:.4f} within_tol={ok.mean():.6f}")
    assert np.isfinite(got).all()
    assert rms <= 0.03 * std, f"rms error {rms:.4f} vs std {std:.4f}"
    assert ok.mean() >= frac, f"only {ok.mean():.5f} within tolerance; max err {err.max():.4f}, std {std:.4f}"
    return tol


@pytest.fixture(scope="module")
def seg1(engine):
    w = nets.seg_weights(11, 1, 3)
    return w, engine.load_model(1, w, 1)


@pytest.mark.parametrize("conv_engine", [0, 1])
def test_seg_forward_binarised_tiles(engine, seg1, conv_engine):
    w, model = seg1
    page, _ = synth.page(2002, 1024, 1024)
    tiles_ref, meta = ho.seg_preprocess(page, closed_form=False)
    tiles, _ = engine.seg_preprocess(torch.from_numpy(page).cuda())
    engine.set_conv_engine(conv_engine)
    try:
        got = engine.seg_forward(model, tiles).cpu().numpy()
        low = engine.seg_forward(model, tiles, lowres=True).cpu().numpy()
    finally:
        engine.set_conv_engine(0)
    x = np.repeat(tiles_ref[:, None].astype(np.float32), 3, axis=1)
    ref = nets.seg_forward(w, x)
    ref_low = nets.seg_forward(w, x, return_lowres=True)
    assert got.shape == ref.shape == (4, 1, 512, 512)
    tol = _tol_check(got, ref)
    _tol_check(low.transpose(0, 3, 1, 2), ref_low)
    # mask decisions agree away from the threshold
    thr = float(ho.sigmoid_logit_threshold(0.9))
    decided = np.abs(ref - thr) > tol
    assert ((got >= thr) == (ref >= thr))[decided].all()
    assert decided.mean() > 0.9


def test_seg_forward_layout_5_classes_rgb(engine):
    w = nets.seg_weights(13, 5, 3)
    model = engine.load_model(1, w, 5)
    page, _ = synth.page(4004, 700, 900, rgb=True)
    chw = np.ascontiguousarray(page.transpose(2, 0, 1))
    x, meta = ho.torch_path_pre(chw)
    tiles, m2 = engine.seg_tile_rgb(torch.from_numpy(page).cuda())
    assert (m2["Hc"], m2["Wc"], m2["pad_x"], m2["pad_y"]) == (meta["Hc"], meta["Wc"], meta["pad_x"], meta["pad_y"])
    np.testing.assert_allclose(tiles.cpu().numpy().transpose(0, 3, 1, 2) / np.float32(255.0), x, rtol=0, atol=1e-6)
    got = engine.seg_forward(model, tiles).cpu().numpy()
    ref = nets.seg_forward(w, x)
    _tol_check(got, ref)
    mask = engine.seg_postprocess_torch(torch.from_numpy(ref).cuda(), m2, 0.8).cpu().numpy()
    mref = ho.torch_path_post(ref, meta, 0.8)
    assert mask.shape == mref.shape
    assert (mask != mref).mean() < 1e-5


def test_seg_torch_path_resized_page(engine):
    """resize_clamp branch (H < 512 -> upsize) of the torch path: tiles within fp32 rounding."""
    page, _ = synth.page(4005, 400, 1500, rgb=True)
    x, meta = ho.torch_path_pre(np.ascontiguousarray(page.transpose(2, 0, 1)))
    tiles, m2 = engine.seg_tile_rgb(torch.from_numpy(page).cuda())
    assert (m2["Hc"], m2["Wc"]) == (meta["Hc"], meta["Wc"])
    np.testing.assert_allclose(tiles.cpu().numpy().transpose(0, 3, 1, 2) / np.float32(255.0), x, rtol=0, atol=2e-6)
    rng = np.random.default_rng(1)
    low = rng.normal(0, 3.0, (x.shape[0], 5, 32, 32)).astype(np.float32)
    logits = torch.nn.functional.interpolate(torch.from_numpy(low), size=(512, 512), mode="bilinear").numpy()
    mask = engine.seg_postprocess_torch(torch.from_numpy(logits).cuda(), m2, 0.8).cpu().numpy()
    mref = ho.torch_path_post(logits, meta, 0.8)
    assert (mask != mref).mean() < 1e-4


@pytest.mark.parametrize("conv_engine", [0, 1])
def test_ocr_forward_and_decode(engine, conv_engine):
    V = 77
    w = nets.easter2_weights(7, V, 80)
    model = engine.load_model(2, w, V)
    crops = synth.line_crops(1001, 6)
    dev, hw, off = engine.pack_crops(crops)
    lines, f32 = engine.ocr_prepare_lines(dev, hw, off, len(crops), 80, 2000, False, want_f32=True)
    engine.set_conv_engine(conv_engine)
    try:
        logits = engine.ocr_forward(model, lines)
    finally:
        engine.set_conv_engine(0)
    got = logits.cpu().numpy()                       # [L,T,V]
    ref = nets.easter2_forward(w, f32.cpu().numpy())  # [L,V,T]
    ref_tv = ref.transpose(0, 2, 1)
    assert got.shape == ref_tv.shape == (6, 500, V)
    tol = _tol_check(got, ref_tv)
    ids, lens = engine.ctc_greedy(logits)
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    # argmax agrees wherever the oracle's top-2 margin exceeds the tolerance
    srt = np.sort(ref_tv, axis=-1)
    margin = srt[..., -1] - srt[..., -2]
    safe = margin > 2 * tol.max(axis=-1)
    assert (got.argmax(-1) == ref_tv.argmax(-1))[safe].all()
    same = 0
    for l in range(len(crops)):
        same += ids[l, : lens[l]].tolist() == ho.ctc_greedy_ids(ref[l], V)
    print("identical decoded lines:", same, "of", len(crops), "; safe steps:", safe.mean())


def test_modern_layout_batch_equals_per_image(engine):
    """predict_masks_batch (tiles of several images in one forward call, mixed page sizes incl. the clamp
    branches) returns exactly what predict_masks returns image by image."""
    from ocr_inference_b200.data import LayoutDetectionConfig
    from ocr_inference_b200.inference import ModernBookFormatLayoutDetection

    classes = ["background", "image", "line", "caption", "margin"]
    det = ModernBookFormatLayoutDetection(LayoutDetectionConfig(nets.seg_weights(13, 5, 3), "", "deeplabv3plus", 512, classes),
                                          engine=engine)
    sizes = [(472, 1300), (700, 1100), (1024, 1536), (2300, 1200), (400, 4300), (600, 600)]
    pages = [synth.page(4100 + i, h, w, rgb=True)[0] for i, (h, w) in enumerate(sizes)]
    want = [det.predict_masks(p, 0.8) for p in pages]
    for max_tiles in (148, 8):
        got = det.predict_masks_batch(pages, 0.8, max_tiles=max_tiles)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g.shape == w.shape and np.array_equal(g, w)


def test_modern_layout_run_writes_json(engine, tmp_path):
    """ModernBookFormatLayoutDetection.run (config C4): directory of images -> per-image bbox JSON,
    masks equal to the oracle's torch-path post-processing of the oracle logits away from ties."""
    import json

    import cv2

    from ocr_inference_b200.data import LayoutDetectionConfig
    from ocr_inference_b200.inference import ModernBookFormatLayoutDetection

    classes = ["background", "image", "line", "caption", "margin"]
    w = nets.seg_weights(13, 5, 3)
    det = ModernBookFormatLayoutDetection(LayoutDetectionConfig(w, "", "deeplabv3plus", 512, classes), engine=engine)
    src = tmp_path / "in"
    src.mkdir()
    for i, (H, W) in enumerate([(600, 800), (400, 1500)]):
        page, _ = synth.page(40 + i, H, W, rgb=True)
        cv2.imwrite(str(src / f"p{i}.png"), cv2.cvtColor(page, cv2.COLOR_RGB2BGR))
    det.run(str(src), str(tmp_path / "out"), None, class_threshold=0.3)
    for i in range(2):
        rec = json.load(open(tmp_path / "out" / f"p{i}.json"))
        assert set(rec) <= set(classes)
        for v in rec.values():
            assert set(v["bbox"]) == {"x", "y", "w", "h"}
    page, _ = synth.page(40, 600, 800, rgb=True)
    masks = det.predict_masks(page, 0.3)
    x, meta = ho.torch_path_pre(np.ascontiguousarray(page.transpose(2, 0, 1)))
    ref_logits = nets.seg_forward(w, x)
    mref = ho.torch_path_post(ref_logits, meta, 0.3)
    assert masks.shape == mref.shape == (5, 600, 800)
    assert (masks != mref).mean() < 0.02      # bf16 logits: only pixels near the 0.3 probability level differ


@pytest.mark.parametrize("n_classes,n_tiles", [(1, 7), (5, 3)])
def test_fused_decoder_tail_equals_the_three_kernel_form(engine, n_classes, n_tiles):
    """dec_tail.cu (depthwise 3x3 -> 1x1 304 -> 256 + ReLU -> head in one kernel) against the three kernels it
    replaces: the depthwise taps accumulate in the same order and both forms round to bf16 at the same two places, so
    the only difference is the fp32 summation order of the head's 256-term dot product."""
    from ocr_inference_b200.synthetic_weights import seg_weights

    model = engine.load_model(1, seg_weights(11 + n_classes, n_classes, 3), n_classes)
    x = (torch.rand((n_tiles, 512, 512), device="cuda", generator=torch.Generator("cuda").manual_seed(5)) > 0.25).to(torch.uint8)
    try:
        engine.set_dec_tail(False)
        ref = engine.seg_forward(model, x, lowres=True).clone()
        engine.set_dec_tail(True)
        got = engine.seg_forward(model, x, lowres=True).clone()
        again = engine.seg_forward(model, x, lowres=True)
    finally:
        engine.set_dec_tail(True)
    torch.cuda.synchronize()
    assert torch.equal(got, again)                          # deterministic
    scale = float(ref.std())
    err = float((got - ref).abs().max())
    print(f"fused vs three-kernel decoder tail: max |diff| {err:.3e} at logit std {scale:.3f}")
    assert err <= 2e-4 * max(1.0, scale) + 1e-4 * float(ref.abs().max())
    assert float(ref.abs().max()) > 0.1


@pytest.mark.parametrize("N,H,residual,relu", [(2, 128, True, True), (5, 128, False, False), (3, 8, True, False), (1, 4, False, True)])
def test_wide_conv_kernel_against_fp32_conv_and_the_slab_kernel(engine, N, H, residual, relu):
    """conv_tc_wide_kernel (64 -> 64 channel 3x3 layers on 128-pixel-wide maps: three horizontal taps in the MMA's N
    dimension, shifted sum in the epilogue, residual through TMA) against torch's fp32 conv2d on the bf16-rounded
    operands, and against the slab kernel it replaces (same products, different fp32 summation order)."""
    import ctypes as C

    g = torch.Generator().manual_seed(100 + N + H)
    x = torch.randn(N, H, 128, 64, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24.0).float()
    b = torch.randn(64, generator=g).float()
    res = torch.randn(N, H, 128, 64, generator=g).to(torch.bfloat16).cuda() if residual else None
    wc = np.ascontiguousarray(w.numpy())
    bc = np.ascontiguousarray(b.numpy())

    def run(wide):
        engine.set_wide_conv(wide)
        y = torch.full((N, H, 128, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        rc = engine.lib.bdrc_conv2d_test(engine.ctx.h, x.data_ptr(), N, H, 128, 64, wc.ctypes.data, bc.ctypes.data, 64, 3, 3,
                                         1, 1, 1, 1, 1, 1, res.data_ptr() if residual else None, int(relu), 0,
                                         y.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, engine.lib.bdrc_last_error(engine.ctx.h)
        torch.cuda.synchronize()
        return y.float()

    try:
        slab = run(False)
        wide = run(True)
        wide2 = run(True)
    finally:
        engine.set_wide_conv(True)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float().cuda(), b.cuda(),
                                     padding=1).permute(0, 2, 3, 1)
    if residual:
        ref = ref + res.float()
    if relu:
        ref = torch.relu(ref)
    assert torch.equal(wide, wide2)                                        # deterministic
    tol = 2e-2 * float(ref.abs().max()) + 1e-3                            # bf16 output rounding
    assert not torch.isnan(wide).any()
    assert float((wide - ref).abs().max()) <= tol
    # the two kernels differ by fp32 summation order only: at most one bf16 ulp of the output
    assert float((wide - slab).abs().max()) <= 2.0 ** -7 * float(ref.abs().max()) + 1e-6
    assert float((wide != slab).float().mean()) < 0.02
