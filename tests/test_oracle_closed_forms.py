"""The pure-numpy closed forms in oracle/hostops.py reproduce the cv2 primitives bit-exactly --
these are the arithmetic definitions the CUDA kernels implement."""
import cv2
import numpy as np
import pytest

from oracle import hostops as ho


def test_gray_and_gaussian_kernel():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)
    assert np.array_equal(ho.gray_u8(img), cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))
    assert np.array_equal(ho.gaussian_kernel(51), cv2.getGaussianKernel(51, 0, cv2.CV_32F).ravel())


@pytest.mark.parametrize("shape", [(512, 512), (80, 2000), (130, 77)])
def test_adaptive_threshold(shape):
    rng = np.random.default_rng(shape[0])
    g = np.clip(rng.normal(190, 35, shape), 0, 255).astype(np.uint8)
    g[shape[0] // 3: shape[0] // 3 + 6, 5: shape[1] // 2] = 30
    g = cv2.GaussianBlur(g, (3, 3), 0)
    ref = cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 51, 13)
    assert np.array_equal(ho.adaptive_threshold(g), ref)


@pytest.mark.parametrize("src,dst", [((100, 300), (80, 240)), ((60, 1500), (80, 2000)), ((472, 2550), (512, 2766)),
                                      ((300, 200), (150, 100)), ((61, 61), (80, 80)), ((77, 333), (80, 345))])
def test_resize_linear_u8(src, dst):
    rng = np.random.default_rng(src[1])
    img = rng.integers(0, 256, (*src, 3), dtype=np.uint8)
    assert np.array_equal(ho.resize_linear_u8(img, dst[1], dst[0]), cv2.resize(img, (dst[1], dst[0]), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("src,dst", [((512, 2766), (472, 2550)), ((1183, 1000), (1300, 1099)), ((400, 300), (200, 150)),
                                      ((512, 1536), (300, 900)), ((1500, 900), (500, 300)), ((512, 768), (1024, 1536))])
def test_resize_linear_f64_binary_maps(src, dst):
    rng = np.random.default_rng(src[0])
    m = np.zeros(src, np.float64)
    for _ in range(20):
        y, x = rng.integers(0, src[0] - 40), rng.integers(0, src[1] - 120)
        m[y:y + rng.integers(8, 40), x:x + rng.integers(30, 120)] = 1.0
    m[:5] = 1.0
    m[:, -4:] = 1.0
    ref = cv2.resize(m, dsize=(dst[1], dst[0])).astype(np.uint8)
    assert np.array_equal(ho.resize_linear_f64(m, dst[1], dst[0]).astype(np.uint8), ref)
    assert 0 < ref.mean() < 1


@pytest.mark.parametrize("shape,angle", [((300, 400), 1.3), ((257, 333), -3.1), ((200, 900), 4.9)])
def test_warp_affine(shape, angle):
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
    M = cv2.getRotationMatrix2D((shape[1] / 2, shape[0] / 2), angle, 1)
    assert np.array_equal(ho.warp_affine_u8(img, M), cv2.warpAffine(img, M, (shape[1], shape[0]), borderValue=(0, 0, 0)))


def test_dilate_anchor_convention():
    m = np.zeros((40, 40), np.uint8)
    m[20, 20] = 255
    for kw, kh in [(4, 6), (5, 3), (1, 1), (7, 2)]:
        d = cv2.dilate(m, cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh)))
        ys, xs = np.nonzero(d)
        up, down, left, right = ho.dilate_rect_anchor(kw, kh)
        assert (20 - ys.min(), ys.max() - 20, 20 - xs.min(), xs.max() - 20) == (up, down, left, right)


def test_spline3_interpolation_restatement_is_scipy_bitwise():
    """The 4x4-tap cubic B-spline step (weights, mirrored taps, constant-0 outside, summation order) restated in
    oracle/hostops.py equals scipy.ndimage.map_coordinates(order=3) bit for bit on scipy's own coefficients --
    this is the arithmetic csrc/lines.cu:sample_cubic_const follows."""
    import scipy.ndimage as ndi

    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53)).astype(np.uint8)
    coef = ndi.spline_filter(img, 3, output=np.float64, mode="constant")
    pts = np.stack([rng.uniform(-2, 39, 1500), rng.uniform(-2, 55, 1500)])
    pts[:, :50] = np.round(pts[:, :50])
    pts[0, 50:60], pts[1, 60:70], pts[0, 70:80] = 36.0, 52.0, 0.0
    want = ndi.map_coordinates(img, pts, order=3, mode="constant", cval=0.0, output=np.float64)
    got = np.array([ho.spline3_interp_numpy(coef, pts[0, k], pts[1, k]) for k in range(pts.shape[1])])
    assert np.array_equal(got, want)
