"""GPU parity: pre/post-processing kernels vs the CPU oracle, bit-exact (integer / byte work)."""
import numpy as np
import pytest
import torch

from oracle import hostops as ho
from oracle import synth

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape", [(600, 700), (512, 512), (472, 2550), (2100, 1500), (1300, 4500), (1024, 1536)])
def test_seg_preprocess_bit_exact(engine, shape):
    page, _ = synth.page(7 + shape[0], *shape)
    tiles, meta = engine.seg_preprocess(_dev(page))
    ref, rmeta = ho.seg_preprocess(page, closed_form=False)   # cv2 path, exactly what the reference runs
    for k in ("Hc", "Wc", "pad_x", "pad_y"):
        assert meta[k] == rmeta[k]
    got = tiles.cpu().numpy()
    assert got.shape == ref.shape
    assert int((got != ref).sum()) == 0


def test_seg_preprocess_full_page_properties(engine):
    """BASELINE size (2048x3072): identical to the oracle on a subset of tiles + value range."""
    page, _ = synth.page(2002, 2048, 3072)
    tiles, meta = engine.seg_preprocess(_dev(page))
    got = tiles.cpu().numpy()
    assert got.shape == (24, 512, 512) and set(np.unique(got)) <= {0, 1}
    ref, _ = ho.seg_preprocess(page, closed_form=False)
    assert int((got != ref).sum()) == 0


@pytest.mark.parametrize("src,dst", [((100, 300), (80, 240)), ((60, 1500), (80, 2000)), ((472, 2550), (512, 2766)),
                                      ((300, 200), (150, 100)), ((61, 61), (80, 80)), ((2100, 1500), (2048, 1462))])
def test_resize_u8_matches_cv2(engine, src, dst):
    import cv2

    rng = np.random.default_rng(src[0] * 7 + dst[1])
    img = rng.integers(0, 256, (*src, 3), dtype=np.uint8)
    got = engine.resize_u8(_dev(img), dst[0], dst[1]).cpu().numpy()
    ref = cv2.resize(img, (dst[1], dst[0]), interpolation=cv2.INTER_LINEAR)
    assert int((got != ref).sum()) == 0


@pytest.mark.parametrize("shape", [(600, 700), (2048, 3072), (472, 2550), (2100, 1500), (1300, 4500)])
def test_line_postprocess_bit_exact(engine, shape):
    """sigmoid > 0.9 -> stitch -> crop -> (cv2 float resize) -> u8*255, on seeded fp32 logits."""
    H0, W0 = shape
    Hc, Wc = ho.clamp_size(H0, W0)
    nty, ntx = -(-Hc // 512), -(-Wc // 512)
    rng = np.random.default_rng(H0 + W0)
    # smooth blobs around the threshold so both decisions occur in runs (exercises the resize)
    low = rng.normal(1.5, 3.0, (nty * ntx, 1, 64, 64)).astype(np.float32)
    logits = torch.nn.functional.interpolate(torch.from_numpy(low), size=(512, 512), mode="bilinear").numpy()
    logits[0, 0, :4, :4] = np.float32(2.1972246)   # exactly the last "False" logit and the first "True"
    logits[0, 0, 4:8, :4] = np.float32(2.1972249)
    meta = dict(H0=H0, W0=W0, Hc=Hc, Wc=Wc, nty=nty, ntx=ntx, y_steps=nty, x_steps=ntx,
                pad_x=ntx * 512 - Wc, pad_y=nty * 512 - Hc)
    got = engine.seg_postprocess(_dev(logits), meta, 0, 0.9).cpu().numpy()
    ref = ho.line_postprocess(logits, meta, 0.9, closed_form=False)
    assert got.shape == ref.shape and got.dtype == np.uint8
    assert int((got != ref).sum()) == 0
    assert 0.05 < (ref == 255).mean() < 0.95


@pytest.mark.parametrize("shape", [(600, 700), (472, 2550), (2100, 1500)])
def test_layout_postprocess(engine, shape):
    H0, W0 = shape
    Hc, Wc = ho.clamp_size(H0, W0)
    nty, ntx = -(-Hc // 512), -(-Wc // 512)
    rng = np.random.default_rng(H0 * 3 + W0)
    low = rng.normal(0, 3.0, (nty * ntx, 5, 32, 32)).astype(np.float32)
    logits = torch.nn.functional.interpolate(torch.from_numpy(low), size=(512, 512), mode="bilinear").numpy()
    meta = dict(H0=H0, W0=W0, Hc=Hc, Wc=Wc, nty=nty, ntx=ntx, y_steps=nty, x_steps=ntx,
                pad_x=ntx * 512 - Wc, pad_y=nty * 512 - Hc)
    got = engine.seg_postprocess(_dev(logits), meta, 1, 0.8).cpu().numpy()
    ref = ho.layout_postprocess(logits, meta, 0.8, closed_form=False)
    assert got.shape == ref.shape
    # softmax is fp32 transcendental: decisions may differ only where p is within 1e-6 of 0.8
    from scipy.special import softmax
    p = softmax(np.transpose(logits, (0, 2, 3, 1)), axis=-1)
    near = int((np.abs(p - 0.8) < 1e-6).sum())
    assert int((got != ref).sum()) <= 4 * near
    assert (ref == 255).mean() > 0.02


@pytest.mark.parametrize("pre_pad", [False, True])
def test_ocr_prepare_lines_bit_exact(engine, pre_pad):
    crops = synth.line_crops(1001, 12) + [synth.line_crop(5, 80, 2000), synth.line_crop(6, 30, 400),
                                          synth.line_crop(7, 200, 900), synth.line_crop(8, 40, 4100),
                                          synth.line_crop(9, 160, 4000)]
    dev, hw, off = engine.pack_crops(crops)
    lines, f32 = engine.ocr_prepare_lines(dev, hw, off, len(crops), 80, 2000, pre_pad, want_f32=True)
    got = f32.cpu().numpy()
    bf = lines.float().cpu().numpy()
    for i, c in enumerate(crops):
        ref = ho.prepare_ocr_line(c, 80, 2000, pre_pad=pre_pad, closed_form=False)[0]
        assert set(np.unique(got[i])) <= {-1.0, 1.0}
        assert int((got[i] != ref).sum()) == 0, f"line {i} shape {c.shape}"
        # network layout: [W][Cpad] time-major, zero padded channels
        assert np.array_equal(bf[i, :, :80], ref.T)
        assert not bf[i, :, 80:].any()


@pytest.mark.parametrize("Ht,Wt", [(80, 2000), (76, 1000), (33, 130)])
def test_ocr_prepare_lines_fills_a_poisoned_buffer(engine, Ht, Wt):
    """The threshold kernel's sink writes EVERY element of the [L][Wt][128] network input -- rows of the image and
    the zero padding channels (16-byte and scalar paths, canvas heights that are not a multiple of the 32-row block)
    -- so a caller-provided slice of a larger batch needs no memset."""
    import torch

    crops = synth.line_crops(77, 5) + [synth.line_crop(3, 30, 400), synth.line_crop(4, 160, 3000)]
    dev, hw, off = engine.pack_crops(crops)
    out = torch.full((len(crops), Wt, 128), 7.0, dtype=torch.bfloat16, device=engine.device)
    lines, f32 = engine.ocr_prepare_lines(dev, hw, off, len(crops), Ht, Wt, False, want_f32=True, out=out)
    bf = lines.float().cpu().numpy()
    got = f32.cpu().numpy()
    for i, c in enumerate(crops):
        ref = ho.prepare_ocr_line(c, Ht, Wt, pre_pad=False, closed_form=False)[0]
        assert int((got[i] != ref).sum()) == 0, f"line {i} shape {c.shape}"
        assert np.array_equal(bf[i, :, :Ht], ref.T)
        assert not bf[i, :, Ht:].any()


@pytest.mark.parametrize("Ht,Wt", [(1, 64), (25, 300), (128, 500), (132, 420), (133, 420), (200, 333), (300, 96)])
def test_ocr_prepare_lines_any_canvas_height(engine, Ht, Wt):
    """The threshold kernel keeps canvases of up to 132 rows whole in its row ring (replicate border materialised)
    and streams taller ones through the ring with its first rows kept twice: both sides of that switch, ring wrap
    included, against cv2 (fp32 reference-layout output only: the network layout holds 128 rows at most)."""
    crops = synth.line_crops(55, 3) + [synth.line_crop(3, 30, 400), synth.line_crop(4, 160, 3000)]
    dev, hw, off = engine.pack_crops(crops)
    _, f32 = engine.ocr_prepare_lines(dev, hw, off, len(crops), Ht, Wt, False, want_f32=True, want_bf16=False)
    got = f32.cpu().numpy()
    for i, c in enumerate(crops):
        ref = ho.prepare_ocr_line(c, Ht, Wt, pre_pad=False, closed_form=False)[0]
        assert int((got[i] != ref).sum()) == 0, f"line {i} shape {c.shape}"


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("L,V,T", [(1, 77, 500), (7, 77, 500), (3, 5, 33), (2, 200, 1000), (4, 77, 1),
                                   (450, 77, 500), (5, 76, 512), (3, 8, 130), (2, 77, 36)])
def test_ctc_greedy_bit_exact(engine, layout, L, V, T):
    rng = np.random.default_rng(L * 1000 + V + T)
    logits = rng.normal(0, 1, (L, T, V)).astype(np.float32)
    logits[:, :, 0] += 1.5                       # blank prior -> runs of blanks and repeats
    logits[0, : min(T, 10), :] = 0.25           # exact ties -> lowest index must win
    if T > 40:
        logits[0, 20:30, 3] = 9.0               # a repeated label that must collapse
        logits[0, 30, 0] = 20.0
        logits[0, 31:35, 3] = 9.0               # same label after a blank must be emitted again
    if T > 40 and L > 1 and V > 9:
        logits[1, 5, 7] = np.nan                # a NaN counts as the maximum (np.argmax), the first one wins
        logits[1, 5, 9] = np.nan
        logits[1, 6, 0] = np.nan                # ... also in column 0
        logits[1, 7, 2] = -0.0                  # -0.0 == 0.0: the first of the two wins
        logits[1, 7, :] = np.minimum(logits[1, 7, :], 0.0)
        logits[1, 7, 4] = 0.0
    dev = _dev(logits if layout == 0 else np.ascontiguousarray(logits.transpose(0, 2, 1)))
    ids, lens = engine.ctc_greedy(dev, layout=layout, blank=0)
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    for l in range(L):
        ref = ho.ctc_greedy_ids(logits[l], V + 1000)   # V+1000: never transpose, input is [T,V]
        assert lens[l] == len(ref)
        assert ids[l, : lens[l]].tolist() == ref


def test_batched_seg_preprocess_bit_exact(engine):
    """a-1 over a page table (one launch for several pages, ragged sizes incl. clamp / upsize / pad) equals the
    cv2 oracle page by page."""
    shapes = [(2048, 3072), (600, 700), (472, 2550), (1024, 1536), (2100, 1500), (512, 512)]
    pages = [synth.page(40 + i, *s)[0] for i, s in enumerate(shapes)]
    tiles, metas = engine.seg_preprocess_batch([_dev(p) for p in pages])
    got = tiles.cpu().numpy()
    o = 0
    for page, meta in zip(pages, metas):
        ref, rmeta = ho.seg_preprocess(page, closed_form=False)
        for k in ("Hc", "Wc", "pad_x", "pad_y"):
            assert meta[k] == rmeta[k]
        n = ref.shape[0]
        assert int((got[o:o + n] != ref).sum()) == 0
        o += n
    assert o == got.shape[0]


def test_batched_line_postprocess_equals_oracle_upsample_decisions(engine):
    """a-3 over a page table with the x4 upsample fused: the mask equals threshold(oracle upsample of the same
    low-res logits) wherever the upsampled logit is not within fp32 rounding of the threshold, and equals the
    per-page kernel bit for bit.  Covers cells that are uniformly above / below the threshold (settled from the
    corners alone), cells that straddle it, +-inf and NaN corners, and ragged page sizes."""
    rng = np.random.default_rng(12)
    shapes = [(2048, 3072), (600, 700), (1024, 1536), (512, 2000)]
    lows, metas = [], []
    for (H, W) in shapes:
        nty, ntx = -(-H // 512), -(-W // 512)
        base = rng.normal(0.0, 6.0, (nty * ntx, 1, 16, 16)).astype(np.float32)
        low = torch.nn.functional.interpolate(torch.from_numpy(base), size=(128, 128), mode="bilinear").numpy()
        low = np.ascontiguousarray(low.transpose(0, 2, 3, 1))            # NHWC [N,128,128,1]
        low[0, 5, 7, 0] = np.inf
        low[0, 9, 9, 0] = -np.inf
        low[0, 20, 20, 0] = np.nan
        low[0, 40:44, 40:44, 0] = np.float32(2.1972246)                  # cells sitting exactly on the threshold
        lows.append(low)
        metas.append(dict(H0=H, W0=W, Hc=H, Wc=W, nty=nty, ntx=ntx, pad_x=ntx * 512 - W, pad_y=nty * 512 - H))
    logits = _dev(np.concatenate(lows))
    masks = engine.seg_postprocess_lowres_batch(logits, metas, 0, 0.9)
    thr = float(ho.sigmoid_logit_threshold(0.9))
    o = 0
    for low, meta, m in zip(lows, metas, masks):
        n = low.shape[0]
        per_page = engine.seg_postprocess(logits[o:o + n], meta, 0, 0.9, lowres=True)
        assert torch.equal(m, per_page)
        o += n
        got = m.cpu().numpy()
        assert got.shape == (meta["H0"], meta["W0"]) and set(np.unique(got)) <= {0, 255}
        with np.errstate(invalid="ignore"):
            up = torch.nn.functional.interpolate(torch.from_numpy(low.transpose(0, 3, 1, 2).copy()), scale_factor=4,
                                                 mode="bilinear", align_corners=True).numpy()[:, 0]
        full = np.vstack([np.hstack(list(r)) for r in np.split(up, meta["nty"])])[:meta["H0"], :meta["W0"]]
        decided = np.isfinite(full) & (np.abs(full - thr) > 1e-4 * np.maximum(1.0, np.abs(full)))
        assert np.array_equal(got[decided] == 255, full[decided] >= thr)
        assert 0.1 < (got == 255).mean() < 0.9
