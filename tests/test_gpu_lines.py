"""GPU parity: rotation, remap, per-line masked crops and the page -> text pipeline vs the oracle."""
import cv2
import numpy as np
import pytest
import torch

from ocr_inference_b200 import synthetic
from oracle import geometry as og
from oracle import hostops as ho
from oracle import nets
from oracle.tps_numpy import ThinPlateSpline

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape,angle", [((300, 400), 1.3), ((512, 700), -2.2), ((2048, 3072), 0.9), ((333, 257), 4.7)])
def test_warp_affine_bit_exact(engine, shape, angle):
    rng = np.random.default_rng(shape[0])
    H, W = shape
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    M = cv2.getRotationMatrix2D((W / 2, H / 2), angle, 1)
    got = engine.warp_affine(_dev(img), M).cpu().numpy()
    ref = ho.rotate_from_angle_ref(img, angle)
    assert int((got != ref).sum()) == 0
    mask = (rng.random((H, W)) > 0.5).astype(np.uint8) * 255
    assert np.array_equal(engine.warp_affine(_dev(mask), M).cpu().numpy(), ho.rotate_from_angle_ref(mask, angle))


def test_remap_bilinear_matches_scipy(engine):
    rng = np.random.default_rng(3)
    H, W = 240, 320
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    rr, cc = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    coords = np.stack([rr + 6 * np.sin(cc / 40) - 2, cc + 4 * np.cos(rr / 30) + 1.5]).astype(np.float32)
    coords[:, 0, :5] = [[-0.4] * 5, [3.0] * 5]          # just outside -> constant 0
    coords[:, 1, :5] = [[H - 1.0] * 5, [W - 1.0] * 5]   # exactly on the last pixel -> inside
    got = engine.remap_bilinear(_dev(img), _dev(coords)).cpu().numpy()
    ref = ho.remap_bilinear_ref(img, coords.astype(np.float64))
    assert int((got != ref).sum()) == 0


def test_tps_remap_matches_oracle(engine):
    rng = np.random.default_rng(5)
    H, W = 300, 500
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    out_pts = np.array([[150, 50], [150, 150], [150, 250], [150, 350], [150, 450]], float)
    in_pts = out_pts + np.array([[8, 0], [-5, 0], [-9, 0], [-4, 0], [7, 0]], float)
    corners = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]) * [H, W]
    tps = ThinPlateSpline(0.5).fit(np.concatenate([out_pts, corners]), np.concatenate([in_pts, corners]))
    grid = np.indices((H, W), dtype=np.float64).transpose(1, 2, 0).reshape(-1, 2)
    coords = tps.transform(grid).reshape(H, W, 2).transpose(2, 0, 1)
    ref = ho.remap_bilinear_ref(img, coords)
    got = engine.tps_remap(_dev(img), tps.control_points, tps.parameters).cpu().numpy()
    # fp64 on both sides but a different summation order / log implementation: coordinates agree to
    # ~1e-12, so results may differ (a) by 1 where the bilinear value sits on a .5 rounding boundary
    # and (b) arbitrarily where the coordinate is within 1e-9 of the image border (inside/outside).
    r, c = coords
    safe = (r > 1e-9) & (r < H - 1 - 1e-9) & (c > 1e-9) & (c < W - 1 - 1e-9)
    outside = (r < -1e-9) | (r > H - 1 + 1e-9) | (c < -1e-9) | (c > W - 1 + 1e-9)
    d = np.abs(got.astype(int) - ref.astype(int)).max(axis=-1)
    assert d[safe].max() <= 1 and (d[safe] > 0).mean() < 1e-3
    assert d[outside].max() == 0
    assert (safe | outside).mean() > 0.99
    from ocr_inference_b200.image_dewarping import run_tps
    got2 = run_tps(img, out_pts.tolist(), in_pts.tolist())
    d2 = np.abs(got2.astype(int) - ref.astype(int)).max(axis=-1)
    assert d2[safe].max() <= 1 and d2[outside].max() == 0


@pytest.mark.parametrize("shape,C", [((240, 320), 3), ((70, 45), 1), ((2, 300), 3), ((130, 3), 1), ((1, 64), 1),
                                     ((700, 900), 3), ((257, 513), 1), ((321, 256), 2)])
def test_remap_cubic_matches_scipy(engine, shape, C):
    """bdrc_spline3_prefilter_u8 + bdrc_remap_cubic_u8 against scipy.ndimage.map_coordinates with its defaults
    (order 3, prefilter, constant 0): coefficients to a few ulp, samples bit-exact except where the fp64 value sits
    within 1e-9 of a .5 rounding boundary."""
    import scipy.ndimage as ndi

    rng = np.random.default_rng(31)
    H, W = shape
    img = rng.integers(0, 256, (H, W, C), dtype=np.uint8)
    coef = engine.spline3_prefilter(_dev(img)).cpu().numpy()
    for k in range(C):
        ref_coef = ndi.spline_filter(img[..., k], 3, output=np.float64, mode="constant")
        assert np.abs(coef[..., k] - ref_coef).max() <= 1e-9 * max(1.0, np.abs(ref_coef).max())
    rr, cc = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    coords = np.stack([rr + 3 * np.sin(cc / 17) - 0.7, cc + 2.5 * np.cos(rr / 13) + 0.4])
    coords[:, 0, : min(W, 4)] = np.array([[-0.4, 0.0, H - 1.0, H - 0.5], [0.0, W - 1.0, 0.0, 1.0]])[:, : min(W, 4)]
    got = engine.remap_cubic(_dev(img), torch.from_numpy(coords)).cpu().numpy()
    ref = ho.remap_cubic_ref(img, coords)
    exact = np.stack([ndi.map_coordinates(img[..., k], coords, output=np.float64) for k in range(C)], axis=-1)
    tie = np.abs((exact + 0.5) - np.round(exact + 0.5)) < 1e-9
    assert got.shape == ref.shape and np.array_equal(got[~tie], ref[~tie])
    assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1 and tie.mean() < 1e-3


def test_tps_remap_cubic_matches_scipy(engine):
    """run_tps(order=3): thin-plate-spline coordinates + cubic sampling, against the oracle's TPS restatement fed
    to scipy's map_coordinates defaults (the reference's intended call)."""
    import scipy.ndimage as ndi

    rng = np.random.default_rng(6)
    H, W = 300, 500
    img = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), 1.5)
    out_pts = np.array([[150, 50], [150, 150], [150, 250], [150, 350], [150, 450]], float)
    in_pts = out_pts + np.array([[8, 0], [-5, 0], [-9, 0], [-4, 0], [7, 0]], float)
    corners = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]) * [H, W]
    tps = ThinPlateSpline(0.5).fit(np.concatenate([out_pts, corners]), np.concatenate([in_pts, corners]))
    grid = np.indices((H, W), dtype=np.float64).transpose(1, 2, 0).reshape(-1, 2)
    coords = tps.transform(grid).reshape(H, W, 2).transpose(2, 0, 1)
    ref = ho.remap_cubic_ref(img, coords)
    from ocr_inference_b200.image_dewarping import run_tps
    got = run_tps(img, out_pts.tolist(), in_pts.tolist(), order=3)
    r, c = coords
    safe = (r > 1e-9) & (r < H - 1 - 1e-9) & (c > 1e-9) & (c < W - 1 - 1e-9)
    outside = (r < -1e-9) | (r > H - 1 + 1e-9) | (c < -1e-9) | (c > W - 1 + 1e-9)
    d = np.abs(got.astype(int) - ref.astype(int)).max(axis=-1)
    # coordinates agree to ~1e-12 (different summation order / log): at most a rounding tie away
    assert d[safe].max() <= 1 and (d[safe] > 0).mean() < 1e-3
    assert d[outside].max() == 0 and (safe | outside).mean() > 0.99
    assert (got != run_tps(img, out_pts.tolist(), in_pts.tolist(), order=1)).mean() > 0.2   # it is not the bilinear path


@pytest.mark.parametrize("order", [1, 3])
def test_global_and_local_tps_dewarp_on_a_curved_page(engine, order):
    """check_for_tps -> apply_global_tps / get_line_images_via_local_tps on a page whose text bands are bent:
    the dewarped page equals the oracle's TPS restatement resampled by scipy (order 1 or the reference's
    default order 3); straight lines of the local variant equal get_line_image."""
    from ocr_inference_b200 import image_dewarping as idw
    from ocr_inference_b200 import line_detection as ld

    H, W = 700, 1100
    page, boxes = synthetic.page(41, H, W, n_lines=(4, 4))
    mask = synthetic.band_mask(boxes, H, W, thin=0.4)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    bend = (yy + 45.0 * np.sin(xx / W * np.pi)).astype(np.float32)
    page_c = cv2.remap(page, xx, bend, cv2.INTER_LINEAR, borderValue=(215, 215, 215))
    mask_c = cv2.remap(mask, xx, bend, cv2.INTER_NEAREST)
    _, mask3, contours, _ = ld.build_raw_line_data(page_c, mask_c)
    filtered = ld.filter_line_contours(mask3, contours)
    ratio, data = idw.check_for_tps(page_c, filtered)
    assert ratio > 0.25 and any(d["tps"] for d in data)
    d = data[idw.get_global_tps_line(data)]
    dew_img, dew_mask = idw.apply_global_tps(page_c, np.asarray(mask3), data, order=order)
    corners = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]) * [H, W]
    tps = ThinPlateSpline(0.5).fit(np.concatenate([np.asarray(d["output_pts"], float), corners]),
                                   np.concatenate([np.asarray(d["input_pts"], float), corners]))
    grid = np.indices((H, W), dtype=np.float64).transpose(1, 2, 0).reshape(-1, 2)
    coords = tps.transform(grid).reshape(H, W, 2).transpose(2, 0, 1)
    ref_fn = ho.remap_cubic_ref if order == 3 else ho.remap_bilinear_ref
    r, c = coords
    safe = (r > 1e-9) & (r < H - 1 - 1e-9) & (c > 1e-9) & (c < W - 1 - 1e-9)
    outside = (r < -1e-9) | (r > H - 1 + 1e-9) | (c < -1e-9) | (c > W - 1 + 1e-9)
    for got, src in ((dew_img, page_c), (dew_mask, np.asarray(mask3))):
        dd = np.abs(got.astype(int) - ref_fn(src, coords).astype(int)).max(axis=-1)
        assert dd[safe].max() <= 1 and (dd[safe] > 0).mean() < 1e-3 and dd[outside].max() == 0
    # the dewarped bands are straighter than the bent ones
    _, m3d, cd, _ = ld.build_raw_line_data(dew_img, cv2.cvtColor(dew_mask, cv2.COLOR_RGB2GRAY))
    ratio_after, _ = idw.check_for_tps(dew_img, ld.filter_line_contours(m3d, cd))
    assert ratio_after <= ratio
    # local variant: one dewarp per curved line; straight lines are plain get_line_image crops
    lines = idw.get_line_images_via_local_tps(page_c, data, k_factor=1.7, order=order)
    assert len(lines) == len(data) and all(im.ndim == 3 and im.shape[0] > 0 for im in lines)
    k = 1.7
    for dline, im in zip(data, lines):
        m = np.zeros((H, W), np.uint8)
        cv2.drawContours(m, [dline["contour"]], -1, 255, -1)
        bh = cv2.boundingRect(dline["contour"])[3]
        if dline["tps"] is True:
            wi = idw.run_tps(page_c, dline["output_pts"], dline["input_pts"], order=order)
            wm = idw.run_tps(m, dline["output_pts"], dline["input_pts"], order=order)[..., 0]
            want, k2 = ld.get_line_image(wi, wm, bh, bbox_tolerance=2.0, k_factor=k)
            k = k2 if k2 != k else k
        else:
            want, _ = ld.get_line_image(page_c, m, bh, bbox_tolerance=2.0, k_factor=k)
        assert np.array_equal(im, want)


def _page_with_mask(seed, H, W, rgb=False, zero_blocks=False):
    page, boxes = synthetic.page(seed, H, W, rgb=rgb)
    if zero_blocks:
        page = page.copy()
        page[:, : W // 3, 1] = 0          # one channel entirely zero on part of the page
        page[H // 2: H // 2 + 3] = 0      # fully black rows crossing a line region
    return page, synthetic.band_mask(boxes, H, W)


@pytest.mark.parametrize("seed,shape,k,tol,rgb,zero", [
    (11, (1024, 1536), 2.5, 4.0, False, False),     # pipeline defaults (inference.py:597-598)
    (12, (900, 1400), 2.0, 2.5, False, False),      # Demo-OCR.ipynb:L159
    (13, (700, 1100), 1.7, 3.0, True, True),        # RGB page with all-zero channel / rows
    (14, (2048, 3072), 2.5, 4.0, False, False),     # BASELINE page size
])
def test_extract_line_images_bit_exact(engine, seed, shape, k, tol, rgb, zero):
    from ocr_inference_b200 import line_detection as ld
    page, mask = _page_with_mask(seed, *shape, rgb=rgb, zero_blocks=zero)
    contours, _ = cv2.findContours(mask, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    lines = [ld.build_line_data(c) for c in contours]
    got = ld.extract_line_images(page, lines, k, tol)
    ref = ho.extract_line_images_ref(page, [ln.contour for ln in lines], k, tol)
    assert len(got) == len(ref) > 5
    for g, r in zip(got, ref):
        assert g.shape == r.shape, (g.shape, r.shape)
        assert np.array_equal(g, r)


def test_mask_n_crop_and_get_line_image(engine):
    from ocr_inference_b200 import line_detection as ld
    page, mask = _page_with_mask(21, 600, 900, rgb=True, zero_blocks=True)
    contours, _ = cv2.findContours(mask, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    one = np.zeros(mask.shape, np.uint8)
    cv2.drawContours(one, [contours[0]], -1, 255, -1)
    assert np.array_equal(ld.mask_n_crop(page, one), ho.mask_n_crop_ref(page, one))
    h = cv2.boundingRect(contours[0])[3]
    g, gk = ld.get_line_image(page, one, h, 2.5, 2.0)
    r, rk = ho.get_line_image_ref(page, one, h, 2.5, 2.0)
    assert gk == rk and np.array_equal(g, r)


def test_pipeline_page_to_text_matches_oracle(engine, monkeypatch):
    """Stages 2-5 of OCRPipeline.run_ocr on a synthetic page with a line-shaped mask injected as
    the stage-1 output (random-init segmentation weights do not produce text lines)."""
    from ocr_inference_b200.data import CharsetEncoder, LineDetectionConfig, OCRModelConfig, OpStatus
    from ocr_inference_b200.inference import OCRPipeline
    from ocr_inference_b200.synthetic_weights import easter2_weights, seg_weights

    V = 77
    we = easter2_weights(7, V, 80)
    pipe = OCRPipeline(
        OCRModelConfig(we, "Easter2", 2000, 80, "input", "output", True, False, CharsetEncoder.STACK,
                       synthetic.CHARSET_76, True, "test"),
        LineDetectionConfig("", seg_weights(11, 1, 3), "deeplabv3plus", 512, ["line"]), None, engine=engine)
    page, mask = _page_with_mask(31, 1024, 1536)
    monkeypatch.setattr(pipe.line_inference, "predict", lambda image, class_threshold=None: mask)
    status, result = pipe.run_ocr(page, k_factor=2.0, bbox_tolerance=2.5)
    assert status == OpStatus.SUCCESS, result
    rot_mask, sorted_lines, ocr_lines, angle = result
    o = og.page_to_lines(page, mask, 2.0, 2.5)
    assert not isinstance(o, str)
    o_mask, o_lines, o_imgs, o_angle = o
    assert isinstance(angle, float) and angle == o_angle
    assert np.array_equal(rot_mask, o_mask)
    assert [(*ln.center,) for ln in sorted_lines] == [(*ln.center,) for ln in o_lines]
    assert [(ln.bbox.x, ln.bbox.y, ln.bbox.w, ln.bbox.h) for ln in sorted_lines] == [ln.bbox for ln in o_lines]
    texts, ids = og.recognise(o_imgs, we, synthetic.CHARSET_76)
    assert len(ocr_lines) == len(texts) == len(sorted_lines)
    same = sum(a.text == b for a, b in zip(ocr_lines, texts))
    print(f"pipeline: {len(texts)} lines, identical decoded lines {same}/{len(texts)} (random-init logits)")
    # failure conventions: an empty mask is a FAILED tuple, never an exception
    monkeypatch.setattr(pipe.line_inference, "predict", lambda image, class_threshold=None: np.zeros_like(mask))
    st, msg = pipe.run_ocr(page)
    assert st == OpStatus.FAILED and msg == "No lines detected"
    assert pipe.run_ocr(None)[0] == OpStatus.FAILED
