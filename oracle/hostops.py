"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the host-side image stages of the hot path.

Two layers per stage:
  * `*_ref`   : the reference's algorithm restated with the same cv2/numpy calls it makes
                (each function cites the reference file:line it follows);
  * closed forms (`gray_u8`, `adaptive_threshold`, `resize_linear_u8`, `resize_linear_f64`,
    `warp_affine_u8`): pure-numpy integer/float restatements of the cv2 primitives, so that
    the arithmetic the CUDA kernels must reproduce is explicit.  tests/test_oracle_*.py pins
    closed form == cv2 (bit-exact) and, when /root/reference is present, == the real reference.

Nothing here is imported by ocr_inference_b200/.
"""
from __future__ import annotations

import math

import cv2
import numpy as np

# ------------------------------------------------------------------------------------------
# cv2 primitives, closed form
# ------------------------------------------------------------------------------------------


def gray_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, COLOR_RGB2GRAY) on u8: 15-bit fixed point (utils.py:799)."""
    i = img.astype(np.int64)
    return ((i[..., 0] * 9798 + i[..., 1] * 19235 + i[..., 2] * 3735 + 16384) >> 15).astype(np.uint8)


def gaussian_kernel(ksize: int = 51) -> np.ndarray:
    """cv2.getGaussianKernel(ksize, sigma<=0, CV_32F): sigma = 0.3*((k-1)*0.5-1)+0.8 (= 8.0 for 51)."""
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    t = np.exp(-0.5 / (sigma * sigma) * x * x)
    return (t / t.sum()).astype(np.float32)


def _fma32(acc: np.ndarray, a: np.ndarray, b) -> np.ndarray:
    """fl32(a*b + acc) with a single rounding (product and sum are exact in float64)."""
    return (acc.astype(np.float64) + a.astype(np.float64) * np.float64(b)).astype(np.float32)


def adaptive_threshold(gray: np.ndarray, block: int = 51, c: int = 13) -> np.ndarray:
    """cv2.adaptiveThreshold(gray, 255, ADAPTIVE_THRESH_GAUSSIAN_C, THRESH_BINARY, block, c).

    fp32 separable Gaussian with BORDER_REPLICATE, row pass then column pass, mean rounded
    half-to-even to u8, out = 255 where gray - mean > -c (utils.py:802-809).  Summation order:
    row pass k = 0..block-1 sequential FMA; column pass centre tap, then symmetric pairs
    (x[-k] + x[k]) * w[k] by FMA -- the order the CUDA kernel uses as well.
    """
    kern = gaussian_kernel(block)
    r = block // 2
    H, W = gray.shape
    f = gray.astype(np.float32)
    pad = np.pad(f, ((0, 0), (r, r)), mode="edge")
    acc = np.zeros((H, W), np.float32)
    for k in range(block):
        acc = _fma32(acc, pad[:, k:k + W], kern[k])
    pad2 = np.pad(acc, ((r, r), (0, 0)), mode="edge")
    s = (pad2[r:r + H].astype(np.float64) * np.float64(kern[r])).astype(np.float32)
    for k in range(1, r + 1):
        pr = (pad2[r - k:r - k + H] + pad2[r + k:r + k + H]).astype(np.float32)
        s = _fma32(s, pr, kern[r + k])
    mean = np.clip(np.rint(s), 0, 255).astype(np.int32)
    return np.where(gray.astype(np.int32) - mean > -c, 255, 0).astype(np.uint8)


def _linear_coefs(ssize: int, dsize: int, vertical: bool):
    """Source index / weight tables of cv2.resize(INTER_LINEAR): fx is formed in float32 from a
    double expression; horizontally out-of-range taps are folded (fx=0), vertically the row index
    is clipped but the weight is kept (that is what cv2 does)."""
    scale = 1.0 / (np.float64(dsize) / np.float64(ssize))
    d = np.arange(dsize, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo = sx < 0
        fx[lo] = 0
        sx[lo] = 0
        hi = sx >= ssize - 1
        fx[hi] = 0
        sx[hi] = ssize - 1
    i0 = np.clip(sx, 0, ssize - 1)
    i1 = np.clip(sx + 1, 0, ssize - 1)
    w0 = (np.float32(1.0) - fx).astype(np.float32)
    return i0, i1, w0, fx


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src u8, (dw, dh), interpolation=INTER_LINEAR): 11-bit fixed point; exact
    2x2 box average when both scales are exactly 1/2 (cv2's INTER_AREA fast path)."""
    squeeze = src.ndim == 2
    if squeeze:
        src = src[..., None]
    H, W = src.shape[:2]
    if (dw, dh) == (W, H):
        out = src.copy()
    elif W == 2 * dw and H == 2 * dh:
        s = src.astype(np.int64)
        out = ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    else:
        x0, x1, a0f, a1f = _linear_coefs(W, dw, False)
        y0, y1, b0f, b1f = _linear_coefs(H, dh, True)
        q = np.float32(2048)
        a0 = np.rint((a0f * q).astype(np.float32)).astype(np.int64)
        a1 = np.rint((a1f * q).astype(np.float32)).astype(np.int64)
        b0 = np.rint((b0f * q).astype(np.float32)).astype(np.int64)[:, None, None]
        b1 = np.rint((b1f * q).astype(np.float32)).astype(np.int64)[:, None, None]
        s = src.astype(np.int64)
        rows = s[:, x0, :] * a0[None, :, None] + s[:, x1, :] * a1[None, :, None]
        S0, S1 = rows[y0], rows[y1]
        v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
        out = np.clip(v, 0, 255).astype(np.uint8)
    return out[..., 0] if squeeze else out


def _linear_coefs_f64(ssize: int, dsize: int, vertical: bool):
    """cv2.resize on CV_64F data keeps the interpolation fraction in double precision (measured:
    cv2 4.13 agrees with a double-fraction lerp to 1e-14 and disagrees with float32 weights at 1e-5)."""
    scale = 1.0 / (np.float64(dsize) / np.float64(ssize))
    d = np.arange(dsize, dtype=np.float64)
    fx = (d + 0.5) * scale - 0.5
    sx = np.floor(fx).astype(np.int64)
    fx = fx - sx
    if not vertical:
        lo = sx < 0
        fx[lo] = 0
        sx[lo] = 0
        hi = sx >= ssize - 1
        fx[hi] = 0
        sx[hi] = ssize - 1
    return np.clip(sx, 0, ssize - 1), np.clip(sx + 1, 0, ssize - 1), fx


def resize_linear_f64(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(float64 {0,1} map, (dw, dh)) FOLLOWED BY astype(uint8) -- the pair of calls in
    Detection._crop_prediction + predict (inference.py:148, 180).  Returned as float64 {0., 1.}.

    The bilinear value of a {0,1} map reaches 1.0 (and survives the truncating cast) exactly when
    every tap with a non-zero weight is 1:  v = s0 + (s1 - s0) * f  is 1 iff s0 == 1 and
    (s1 == 1 or f == 0), horizontally then vertically.  Pinned bit-exact against cv2 on 20 M pixels
    (tests/test_oracle_closed_forms.py)."""
    squeeze = src.ndim == 2
    if squeeze:
        src = src[..., None]
    H, W, C = src.shape
    if (dw, dh) == (W, H):
        out = (src >= 1.0)
    elif C in (1, 3, 4):
        # cv2's IPP route for 1/3/4-channel CV_64F images: double fraction, lerp form
        x0, x1, fx = _linear_coefs_f64(W, dw, False)
        y0, y1, fy = _linear_coefs_f64(H, dh, True)
        b = src >= 1.0
        row_one = b[:, x0, :] & (b[:, x1, :] | (fx == 0)[None, :, None])
        out = row_one[y0] & (row_one[y1] | (fy == 0)[:, None, None])
    else:
        # other channel counts (the 5-class layout map): OpenCV's generic route -- float32 weights,
        # double accumulation; the value must reach 1.0 to survive the truncating cast
        x0, x1, a0, a1 = _linear_coefs(W, dw, False)
        y0, y1, b0, b1 = _linear_coefs(H, dh, True)
        s = src.astype(np.float64)
        rows = s[:, x0, :] * a0.astype(np.float64)[None, :, None] + s[:, x1, :] * a1.astype(np.float64)[None, :, None]
        v = rows[y0] * b0.astype(np.float64)[:, None, None] + rows[y1] * b1.astype(np.float64)[:, None, None]
        out = v >= 1.0
    out = out.astype(np.float64)
    return out[..., 0] if squeeze else out


def warp_affine_u8(src: np.ndarray, M: np.ndarray) -> np.ndarray:
    """cv2.warpAffine(src, M, (W, H), flags=INTER_LINEAR, borderValue=0) on u8
    (line_detection.py:74): M is inverted in double, coordinates are 10-bit fixed point rounded
    to 1/32 pixel, bilinear weights are 15-bit ((32-fx)*(32-fy)*32 ...), result rounded."""
    squeeze = src.ndim == 2
    if squeeze:
        src = src[..., None]
    H, W, C = src.shape
    M = np.asarray(M, dtype=np.float64).reshape(2, 3)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    A12, A21 = -M[0, 1] * D, -M[1, 0] * D
    b1 = -A11 * M[0, 2] - A12 * M[1, 2]
    b2 = -A21 * M[0, 2] - A22 * M[1, 2]
    AB = 1024.0
    xs = np.arange(W, dtype=np.float64)
    ys = np.arange(H, dtype=np.float64)
    adelta = np.rint(A11 * xs * AB).astype(np.int64)
    bdelta = np.rint(A21 * xs * AB).astype(np.int64)
    X0 = np.rint((A12 * ys + b1) * AB).astype(np.int64) + 16
    Y0 = np.rint((A22 * ys + b2) * AB).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx, sy = X >> 5, Y >> 5
    fx, fy = X & 31, Y & 31
    w00 = (32 - fx) * (32 - fy) * 32
    w01 = fx * (32 - fy) * 32
    w10 = (32 - fx) * fy * 32
    w11 = fx * fy * 32
    s = src.astype(np.int64)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = s[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return v * ok[..., None]

    v = (tap(sy, sx) * w00[..., None] + tap(sy, sx + 1) * w01[..., None] +
         tap(sy + 1, sx) * w10[..., None] + tap(sy + 1, sx + 1) * w11[..., None] + (1 << 14)) >> 15
    out = np.clip(v, 0, 255).astype(np.uint8)
    return out[..., 0] if squeeze else out


# ------------------------------------------------------------------------------------------
# a-1  segmentation pre-processing  (BDRC/inference.py:132-139, BDRC/utils.py:713-815)
# ------------------------------------------------------------------------------------------


def clamp_size(H: int, W: int, patch: int = 512, clamp_w: int = 4096, clamp_h: int = 2048):
    """Target (H, W) after the clamp/upsize branch of utils.preprocess_image (:764-773).
    resize_to_width/height truncate the other side with int() (:473, :491)."""
    if W > H and W > clamp_w:
        return int(H * (clamp_w / W)), clamp_w
    if H > W and H > clamp_h:
        return clamp_h, int(W * (clamp_h / H))
    if H < patch:
        return patch, int(W * (patch / H))
    return H, W


def preprocess_image_ref(image: np.ndarray, patch: int = 512):
    """utils.preprocess_image (:741-778) with the reference's own cv2 calls."""
    H, W = image.shape[:2]
    Hc, Wc = clamp_size(H, W, patch)
    if (Hc, Wc) != (H, W):
        image = cv2.resize(image, (Wc, Hc), interpolation=cv2.INTER_LINEAR)
    pad_x = math.ceil(Wc / patch) * patch - Wc
    pad_y = math.ceil(Hc / patch) * patch - Hc
    padded = np.pad(image, ((0, pad_y), (0, pad_x), (0, 0)), mode="constant", constant_values=255)
    return padded, pad_x, pad_y


def tile_image_ref(padded: np.ndarray, patch: int = 512):
    """utils.tile_image (:713-721): row-major list of patch x patch tiles."""
    ys, xs = padded.shape[0] // patch, padded.shape[1] // patch
    tiles = [padded[y * patch:(y + 1) * patch, x * patch:(x + 1) * patch] for y in range(ys) for x in range(xs)]
    return tiles, ys


def binarize_ref(img: np.ndarray, block: int = 51, c: int = 13) -> np.ndarray:
    """utils.binarize (:796-815), adaptive branch, with cv2 as the reference calls it."""
    g = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
    bw = cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, block, c)
    return cv2.cvtColor(bw, cv2.COLOR_GRAY2RGB)


def binarize(img: np.ndarray) -> np.ndarray:
    """Closed form of utils.binarize: one channel u8 {0,255} (the three are identical)."""
    return adaptive_threshold(gray_u8(img))


def seg_preprocess(image: np.ndarray, patch: int = 512, closed_form: bool = True):
    """Detection._preprocess_image (inference.py:132-139).

    Returns tiles u8 {0,1} [N, patch, patch] (the reference's (N,patch,patch,3) fp32 tensor has
    three identical channels with exactly these values) and a meta dict.
    """
    H0, W0 = image.shape[:2]
    Hc, Wc = clamp_size(H0, W0, patch)
    if (Hc, Wc) != (H0, W0):
        image = resize_linear_u8(image, Wc, Hc) if closed_form else cv2.resize(
            image, (Wc, Hc), interpolation=cv2.INTER_LINEAR)
    pad_x = math.ceil(Wc / patch) * patch - Wc
    pad_y = math.ceil(Hc / patch) * patch - Hc
    padded = np.pad(image, ((0, pad_y), (0, pad_x), (0, 0)), mode="constant", constant_values=255)
    tiles, y_steps = tile_image_ref(padded, patch)
    if closed_form:
        out = np.stack([binarize(t) // 255 for t in tiles])
    else:
        out = np.stack([binarize_ref(t)[..., 0] // 255 for t in tiles])
    meta = dict(H0=H0, W0=W0, Hc=Hc, Wc=Wc, pad_x=pad_x, pad_y=pad_y, y_steps=y_steps,
                x_steps=padded.shape[1] // patch)
    return out.astype(np.uint8), meta


# ------------------------------------------------------------------------------------------
# a-3  mask thresholding  (BDRC/inference.py:141-150, 170-183, 263-276; utils.py:563-573, 724-729)
# ------------------------------------------------------------------------------------------


def sigmoid_ref(x):
    """utils.sigmoid (:563-573): 1 / (1 + exp(-x)) in the array's dtype (fp32)."""
    return 1 / (1 + np.exp(-x))


def stitch_ref(pred: np.ndarray, y_steps: int) -> np.ndarray:
    """utils.stitch_predictions (:724-729)."""
    rows = np.split(pred, y_steps, axis=0)
    return np.vstack([np.hstack(list(r)) for r in rows])


def line_postprocess(logits: np.ndarray, meta: dict, thr: float = 0.9, closed_form: bool = True) -> np.ndarray:
    """LineDetection.predict after _predict (inference.py:175-181).
    logits fp32 [N,1,P,P] -> u8 mask [H0,W0] in {0,255}."""
    p = np.squeeze(logits, axis=1)
    p = sigmoid_ref(p)
    p = np.where(p > thr, 1.0, 0.0)
    merged = stitch_ref(p, meta["y_steps"])
    merged = merged[: merged.shape[0] - meta["pad_y"], : merged.shape[1] - meta["pad_x"]]
    if closed_form:
        merged = resize_linear_f64(merged, meta["W0"], meta["H0"])
    else:
        merged = cv2.resize(merged, dsize=(meta["W0"], meta["H0"]))
    merged = merged.astype(np.uint8)
    merged *= 255
    return merged


def layout_postprocess(logits: np.ndarray, meta: dict, thr: float = 0.8, closed_form: bool = True) -> np.ndarray:
    """LayoutDetection.predict after _predict (inference.py:268-274): [N,C,P,P] -> [H0,W0,C]."""
    from scipy.special import softmax

    p = np.transpose(logits, (0, 2, 3, 1))
    p = softmax(p, axis=-1)
    p = np.where(p > thr, 1.0, 0)
    merged = stitch_ref(p, meta["y_steps"])
    merged = merged[: merged.shape[0] - meta["pad_y"], : merged.shape[1] - meta["pad_x"]]
    if closed_form:
        merged = resize_linear_f64(merged, meta["W0"], meta["H0"])
    else:
        merged = cv2.resize(merged, dsize=(meta["W0"], meta["H0"]))
        if merged.ndim == 2:
            merged = merged[..., None]
    merged = merged.astype(np.uint8)
    merged *= 255
    return merged


def sigmoid_logit_threshold(thr: float = 0.9) -> np.float32:
    """Smallest fp32 logit x with fp32 sigmoid(x) > fp32(thr), found by bisection on the
    reference predicate (utils.py:563-573 + inference.py:177; numpy-2 weak scalar => the
    comparison is against float32(thr))."""
    def pred(x):
        return bool(sigmoid_ref(np.float32(x)) > np.float32(thr))
    lo, hi = np.float32(-30.0), np.float32(30.0)
    assert not pred(lo) and pred(hi)
    while True:
        mid = np.float32((np.float64(lo) + np.float64(hi)) / 2)
        if mid == lo or mid == hi:
            break
        if pred(mid):
            hi = mid
        else:
            lo = mid
    return hi


# ------------------------------------------------------------------------------------------
# a-7  OCR line preparation  (BDRC/inference.py:313-357; utils.py:457-494, 818-876)
# ------------------------------------------------------------------------------------------


def pre_pad_ref(image: np.ndarray) -> np.ndarray:
    """OCRInference._pre_pad (inference.py:349-357)."""
    h, _, c = image.shape
    patch = np.full((h, h, c), 255, np.uint8)
    return np.hstack([patch, image, patch])


def ocr_fit_geometry(h: int, w: int, Ht: int, Wt: int):
    """Resize target and padding of _pad_ocr_line (inference.py:319-328) + pad_to_width/height
    (utils.py:818-876).  Returns (rh, rw, top, left): resized size and its offset in the canvas."""
    wr, hr = Wt / w, Ht / h
    if wr <= hr:  # pad_to_width (also the tie case, :327-328)
        rw, rh = Wt, int(h * (Wt / w))
        return rh, rw, (Ht - rh) // 2, 0
    rh, rw = Ht, int(w * (Ht / h))
    return rh, rw, 0, (Wt - rw) // 2


def pad_ocr_line(img: np.ndarray, Ht: int, Wt: int, closed_form: bool = True) -> np.ndarray:
    """_pad_ocr_line: aspect-preserving resize, black padding, final (identity) resize."""
    h, w = img.shape[:2]
    rh, rw, top, left = ocr_fit_geometry(h, w, Ht, Wt)
    if rh > Ht or rw > Wt or rh <= 0 or rw <= 0:
        # the reference's np.zeros(shape=negative) raises here (utils.py:834-838)
        raise ValueError("line does not fit the OCR canvas")
    r = resize_linear_u8(img, rw, rh) if closed_form else cv2.resize(img, (rw, rh), interpolation=cv2.INTER_LINEAR)
    canvas = np.zeros((Ht, Wt, img.shape[2]), np.uint8)
    canvas[top:top + rh, left:left + rw] = r
    return canvas


def prepare_ocr_line(img: np.ndarray, Ht: int = 80, Wt: int = 2000, pre_pad: bool = False,
                     closed_form: bool = True) -> np.ndarray:
    """OCRInference._prepare_ocr_line (+ optional _pre_pad): -> fp32 [1,Ht,Wt] in {-1,+1}
    (inference.py:336-347)."""
    if pre_pad:
        img = pre_pad_ref(img)
    canvas = pad_ocr_line(img, Ht, Wt, closed_form)
    if closed_form:
        bw = binarize(canvas)
    else:
        bw = cv2.cvtColor(binarize_ref(canvas), cv2.COLOR_RGB2GRAY)
    x = bw.reshape((1, Ht, Wt))
    x = (x / 127.5) - 1.0
    return x.astype(np.float32)


# ------------------------------------------------------------------------------------------
# a-9  greedy CTC  (BDRC/inference.py:90-93, 108-115, 372-384, 634-638)
# ------------------------------------------------------------------------------------------


def ctc_greedy_ids(logits: np.ndarray, V: int, blank: int = 0) -> list[int]:
    """logits [V,T] or [T,V] (transposed like _decode :373-376) -> collapsed, blank-free ids."""
    if logits.shape[0] == V and logits.shape[-1] != V:
        logits = logits.T
    elif logits.shape[0] == V and logits.shape[1] == V:
        logits = logits.T  # the reference transposes whenever shape[0] == V
    ids = np.argmax(logits, axis=-1)
    out, prev = [], -1
    for i in ids.tolist():
        if i != prev and i != blank:
            out.append(i)
        prev = i
    return out


def ctc_ids_to_text(ids: list[int], charset: list[str]) -> str:
    """CTCDecoder.decode index convention (inference.py:111-112): charset[i-1]; then the
    pipeline's .strip().replace('§', ' ') (inference.py:634-638)."""
    return "".join(charset[i - 1] for i in ids).strip().replace("§", " ")


# ------------------------------------------------------------------------------------------
# a-5  per-line masked crops  (BDRC/line_detection.py:79-104, 267-359)
# ------------------------------------------------------------------------------------------


def mask_n_crop_ref(image: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """line_detection.mask_n_crop (:79-104): AND with the mask, then delete every row/column in
    which ANY channel is entirely zero (that is what `~x.any(axis)` on an HWC array does)."""
    image = image.astype(np.uint8)
    mask = mask.astype(np.uint8)
    if image.ndim == 2:
        image = image[..., None]
    m = cv2.bitwise_and(image, image, mask, mask)
    m = np.delete(m, np.where(~m.any(axis=1))[0], axis=0)
    m = np.delete(m, np.where(~m.any(axis=0))[0], axis=1)
    return m


def extract_line_ref(image, mask, bbox_h: int, k_factor: float):
    """line_detection.extract_line (:267-288)."""
    k = int(bbox_h * k_factor)
    se = cv2.getStructuringElement(cv2.MORPH_RECT, (k, int(k * k_factor)))
    dil = cv2.dilate(mask, kernel=se, iterations=1)
    return mask_n_crop_ref(image, dil)


def get_line_image_ref(image, mask, bbox_h: int, bbox_tolerance: float, k_factor: float):
    """line_detection.get_line_image (:291-327), including the shrink-k retry loop."""
    try:
        k = k_factor
        img = extract_line_ref(image, mask, bbox_h, k)
        attempts = 0
        while img.shape[0] > bbox_h * bbox_tolerance and attempts < 10:
            k = k - 0.1
            if k <= 0.1:
                break
            img = extract_line_ref(image, mask, bbox_h, k)
            attempts += 1
        return img, k
    except Exception:  # noqa: BLE001  (the reference swallows everything here, :322-327)
        return np.zeros((bbox_h, bbox_h * 2, 3), np.uint8), k_factor


def extract_line_images_ref(image, contours, default_k: float = 1.7, bbox_tolerance: float = 3):
    """line_detection.extract_line_images (:330-359) over a list of contours; the adapted k
    carries over to the next line (:353-357)."""
    k = default_k
    out = []
    for cnt in contours:
        _, _, _, h = cv2.boundingRect(cnt)
        m = np.zeros(image.shape[:2], np.uint8)
        cv2.drawContours(m, [cnt], -1, (255, 255, 255), -1)
        img, k2 = get_line_image_ref(image, m, h, bbox_tolerance, k)
        out.append(img)
        if k != k2:
            k = k2
    return out


def dilate_rect_anchor(kw: int, kh: int):
    """cv2.dilate with a kw x kh MORPH_RECT and the default anchor (kw//2, kh//2):
    out(y,x) = max over dy in [-(kh//2), kh-1-kh//2], dx in [-(kw//2), kw-1-kw//2] of in(y+dy, x+dx).
    Returns the (up, down, left, right) reach of a set pixel in the OUTPUT (mirrored)."""
    ax, ay = kw // 2, kh // 2
    # in(y', x') = 1 sets out(y, x) for y = y' - dy, dy in [-ay, kh-1-ay]  =>  y in [y'-(kh-1-ay), y'+ay]
    return kh - 1 - ay, ay, kw - 1 - ax, ax


# ------------------------------------------------------------------------------------------
# a-4  rotation  (BDRC/line_detection.py:62-76)
# ------------------------------------------------------------------------------------------


def rotate_from_angle_ref(image: np.ndarray, angle: float) -> np.ndarray:
    rows, cols = image.shape[:2]
    M = cv2.getRotationMatrix2D((cols / 2, rows / 2), angle, 1)
    return cv2.warpAffine(image, M, (cols, rows), borderValue=(0, 0, 0))


# ------------------------------------------------------------------------------------------
# a-6  remap  (BDRC/image_dewarping.py:22-74, bilinear per north_star)
# ------------------------------------------------------------------------------------------


def remap_bilinear_ref(image: np.ndarray, coords: np.ndarray) -> np.ndarray:
    """scipy.ndimage.map_coordinates(channel, coords, order=1, mode='constant', cval=0) per
    channel (image_dewarping.py:66-72 with order=1), u8 output rounds half to even like scipy."""
    import scipy.ndimage

    return np.stack([scipy.ndimage.map_coordinates(image[..., c], coords, order=1, mode="constant", cval=0)
                     for c in range(image.shape[2])], axis=-1)


def remap_cubic_ref(image: np.ndarray, coords: np.ndarray) -> np.ndarray:
    """The call the reference makes (image_dewarping.py:66-72): scipy.ndimage.map_coordinates with its
    defaults -- order=3 cubic B-spline, prefilter, mode='constant', cval=0 -- per channel."""
    import scipy.ndimage

    return np.stack([scipy.ndimage.map_coordinates(image[..., c], coords) for c in range(image.shape[2])], axis=-1)


def spline3_interp_numpy(coef: np.ndarray, r: float, c: float) -> float:
    """Restatement of scipy's cubic interpolation step on prefiltered coefficients (ni_interpolation.c
    NI_GeometricTransform + ni_splines.c get_spline_interpolation_weights, mode 'constant'): bitwise equal to
    scipy given scipy's own coefficients (tests/test_oracle_closed_forms.py); small cases only."""
    import math

    H, W = coef.shape
    if r < 0 or r > H - 1 or c < 0 or c > W - 1:
        return 0.0

    def weights(x):
        x = x - math.floor(x)
        y, z = x, 1.0 - x
        w = [z * z * z / 6.0, (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0, (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0, 1.0]
        for i in range(3):
            w[3] -= w[i]
        return w

    def mirror(i, n):
        if n <= 1:
            return 0
        s2 = 2 * n - 2
        if i < 0:
            i = s2 * int(-i / s2) + i
            return i + s2 if i <= 1 - n else -i
        if i >= n:
            i -= s2 * int(i / s2)
            if i >= n:
                i = s2 - i
        return i

    sr, sc = math.floor(r) - 1, math.floor(c) - 1
    wr, wc = weights(r), weights(c)
    t = 0.0
    for i in range(4):
        for j in range(4):
            v = coef[mirror(sr + i, H), mirror(sc + j, W)]
            v *= wr[i]
            v *= wc[j]
            t += v
    return t


# ------------------------------------------------------------------------------------------
# a-11 torch-path pre/post  (BDRC/utils.py:967-1051, 1100-1108, 444-454; inference.py:943-956)
# ------------------------------------------------------------------------------------------


def torch_path_pre(img_chw_u8, patch: int = 512):
    """multi_image_collate_fn for one image (utils.py:1142-1172): resize_clamp (float bilinear,
    round()), pad 255, unfold tiles, /255.  Returns tiles fp32 [N,3,P,P] and meta."""
    import torch
    import torch.nn.functional as F

    img = torch.as_tensor(img_chw_u8)
    _, H, W = img.shape
    sx = sy = 1.0
    scale = None
    if W > H and W > 4096:
        scale = 4096 / W
    elif H > W and H > 2048:
        scale = 2048 / H
    elif H < patch:
        scale = patch / H
    if scale is not None:
        nh, nw = int(round(H * scale)), int(round(W * scale))
        sx, sy = nw / W, nh / H
        img = F.interpolate(img.unsqueeze(0).float(), size=(nh, nw), mode="bilinear", align_corners=False).squeeze(0)
    _, H1, W1 = img.shape
    pad_h = (patch - H1 % patch) % patch
    pad_w = (patch - W1 % patch) % patch
    img = F.pad(img, (0, pad_w, 0, pad_h), value=255)
    C = img.shape[0]
    ys, xs = img.shape[1] // patch, img.shape[2] // patch
    tiles = img.unfold(1, patch, patch).unfold(2, patch, patch).permute(1, 2, 0, 3, 4).contiguous().view(-1, C, patch, patch)
    tiles = tiles.float().div_(255.0)
    meta = dict(H0=H, W0=W, Hc=H1, Wc=W1, pad_x=pad_w, pad_y=pad_h, x_steps=xs, y_steps=ys, scale_x=sx, scale_y=sy)
    return tiles.numpy(), meta


def torch_path_post(logits: np.ndarray, meta: dict, thr: float = 0.8) -> np.ndarray:
    """inference.py:947-956: softmax(dim=1) -> stitch -> crop -> bilinear resize to the original
    size (align_corners=False) -> > thr -> u8*255.  Returns [C,H0,W0]."""
    import torch
    import torch.nn.functional as F

    soft = torch.softmax(torch.as_tensor(logits), dim=1)
    N, C, P, _ = soft.shape
    t = soft.view(meta["y_steps"], meta["x_steps"], C, P, P)
    full = torch.cat([torch.cat(list(t[y]), dim=-1) for y in range(meta["y_steps"])], dim=-2)
    if meta["pad_y"] > 0:
        full = full[:, :-meta["pad_y"], :]
    if meta["pad_x"] > 0:
        full = full[:, :, :-meta["pad_x"]]
    full = F.interpolate(full.unsqueeze(0).float(), size=(meta["H0"], meta["W0"]), mode="bilinear",
                         align_corners=False).squeeze(0)
    return ((full > thr).to(torch.uint8) * 255).numpy()
