"""Drop-in for the TPS dewarping stage of BDRC/image_dewarping.py.

The distortion analysis (check_line_tps / check_for_tps / get_global_tps_line, :118-264) is host
geometry on a few contours and stays on the CPU.  The per-pixel part -- evaluating the thin-plate
spline on the H x W grid and resampling the page -- is ONE fused GPU kernel (bdrc_tps_remap_u8).

Deviation from the reference, on purpose: reference run_tps (:22-74) cannot run
(`npt.NDArray(list)` raises TypeError at :42, and the corner control points are scaled twice at
:55-56).  This module implements the INTENDED behaviour: np.asarray control points, corners scaled
once, thin-plate spline with alpha = 0.5 (thin-plate-spline==1.2.0 algorithm) and -- as
north_star specifies -- order-1 (bilinear) sampling with constant-0 border by default.  `order=3` selects
scipy's default cubic B-spline sampling, i.e. exactly what the reference's map_coordinates call would compute
(spline prefilter + 4x4 taps, both on the GPU in fp64).  Parity is checked against
scipy.ndimage.map_coordinates(order=1 / order=3).
"""
from __future__ import annotations

import cv2
import numpy as np
import torch


class ThinPlateSpline:
    """U(r) = r^2 log r spline with regularisation alpha (what `tps.ThinPlateSpline` computes)."""

    def __init__(self, alpha: float = 0.0):
        self.alpha = float(alpha)
        self.control_points = None
        self.parameters = None

    @staticmethod
    def _u(d2: np.ndarray) -> np.ndarray:
        out = np.zeros_like(d2)
        nz = d2 > 0
        out[nz] = 0.5 * d2[nz] * np.log(d2[nz])
        return out

    def fit(self, X, Y):
        X = np.asarray(X, dtype=np.float64).reshape(len(X), -1)
        Y = np.asarray(Y, dtype=np.float64).reshape(len(Y), -1)
        n, d = X.shape
        self.control_points = X
        K = self._u(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)) + self.alpha * np.eye(n)
        P = np.hstack([np.ones((n, 1)), X])
        A = np.block([[K, P], [P.T, np.zeros((d + 1, d + 1))]])
        self.parameters = np.linalg.solve(A, np.vstack([Y, np.zeros((d + 1, Y.shape[1]))]))
        return self

    def transform(self, X):
        X = np.asarray(X, dtype=np.float64).reshape(len(X), -1)
        U = self._u(((X[:, None, :] - self.control_points[None, :, :]) ** 2).sum(-1))
        return np.hstack([U, np.ones((len(X), 1)), X]) @ self.parameters


def run_tps(image: np.ndarray, input_pts, output_pts, add_corners: bool = True, alpha: float = 0.5,
            order: int = 1) -> np.ndarray:
    """Intended behaviour of image_dewarping.run_tps (:22-74): fit output grid -> input coordinates,
    resample every channel on the GPU, bilinearly (order=1) or with the cubic B-spline of scipy's
    map_coordinates default (order=3)."""
    if order not in (1, 3):
        raise ValueError("run_tps: order must be 1 (bilinear) or 3 (cubic B-spline)")
    from .inference import get_engine

    if image.ndim == 2:
        image = cv2.cvtColor(image, cv2.COLOR_GRAY2RGB)
    H, W = image.shape[:2]
    inp = np.asarray(input_pts, dtype=np.float64)
    out = np.asarray(output_pts, dtype=np.float64)
    if add_corners:
        corners = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]) * [H, W]
        inp = np.concatenate((inp, corners))
        out = np.concatenate((out, corners))
    tps = ThinPlateSpline(alpha).fit(inp, out)
    eng = get_engine()
    src = torch.from_numpy(np.ascontiguousarray(image, dtype=np.uint8)).to(eng.device)
    remap = eng.tps_remap_cubic if order == 3 else eng.tps_remap
    return remap(src, tps.control_points, tps.parameters).cpu().numpy()


def _slice_center(mask_roi: np.ndarray, s0: int, s1: int, x: int, y: int, page_w: int):
    """get_global_center (:77-115) on the page columns [x + s0, x + s1) of a line whose filled contour is
    `mask_roi` at page offset (x, y).  The reference slices a page-sized canvas, so a slice may reach beyond the
    line's bounding box (lines narrower than 4 slice widths: zeros there) and follows Python's slice rules at
    the page border; the x offset it adds back is the unclamped slice start."""
    a, b, _ = slice(x + s0, x + s1).indices(page_w)
    sl = np.zeros((mask_roi.shape[0], max(0, b - a)), np.uint8)
    lo, hi = max(a, x), min(b, x + mask_roi.shape[1])
    if hi > lo:
        sl[:, lo - a:hi - a] = mask_roi[:, lo - x:hi - x]
    gx = x + s0
    found, _ = cv2.findContours(sl, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE) if sl.size else ((), None)
    if not found:
        return gx + sl.shape[1] // 2, y + sl.shape[0] // 2, sl.shape[0]
    areas = [cv2.contourArea(c) for c in found]
    big = found[areas.index(max(areas))]
    _, _, _, bh = cv2.boundingRect(big)
    (cx, cy), _, _ = cv2.minAreaRect(big)
    return gx + int(cx), y + int(cy), bh


def check_line_tps(image: np.ndarray, contour: np.ndarray, slice_width: int = 40):
    """:118-193 -- five slices along the line; needs TPS when their centres spread more than the
    mean slice height.  Works on the line's ROI instead of a page-sized canvas."""
    x, y, w, h = cv2.boundingRect(contour)
    roi = np.zeros((h, w), np.uint8)
    cv2.drawContours(roi, [contour], 0, 255, -1, offset=(-x, -y))
    starts = [0, w // 4 - slice_width, w // 2, w // 2 + w // 4, w - slice_width]
    cs = [_slice_center(roi, s, s + slice_width, x, y, image.shape[1]) for s in starts]
    ys = [c[1] for c in cs]
    spread = max(ys) - min(ys)
    if spread > np.mean([c[2] for c in cs]):
        ty = round(np.mean(ys))
        return True, [[c[1], c[0]] for c in cs], [[ty, c[0]] for c in cs], spread
    return False, None, None, 0.0


def check_for_tps(image: np.ndarray, line_contours: list):
    """:196-227."""
    data = []
    for c in line_contours:
        st, ip, op, yd = check_line_tps(image, c)
        data.append({"contour": c, "tps": st, "input_pts": ip, "output_pts": op, "max_yd": yd})
    return len([d for d in data if d["tps"] is True]) / len(line_contours), data


def get_global_tps_line(line_data: list) -> int:
    """:230-264: the curved line whose spread is closest to the mean spread."""
    deltas = [d["max_yd"] if d["tps"] is True else 0.0 for d in line_data]
    mean, best, pick = np.mean(deltas), max(deltas), None
    for yd in deltas:
        if yd > 0 and abs(mean - yd) < best:
            best, pick = abs(mean - yd), yd
    return deltas.index(pick)


def apply_global_tps(image: np.ndarray, line_mask: np.ndarray, line_data: list, order: int = 1):
    """:267-289."""
    d = line_data[get_global_tps_line(line_data)]
    if d["input_pts"] is None or d["output_pts"] is None:
        raise ValueError("input_pts and output_pts must not be None when tps is True")
    return (run_tps(image, d["output_pts"], d["input_pts"], order=order),
            run_tps(line_mask, d["output_pts"], d["input_pts"], order=order))


def get_line_images_via_local_tps(image: np.ndarray, line_data: list, k_factor: float = 1.7, order: int = 1) -> list:
    """:292-342: one thin-plate-spline dewarp per curved line (page and that line's filled contour), then the
    shrink-k masked crop of line_detection.get_line_image with bbox_tolerance 2.0; the adapted k carries over
    to the following lines.  `line_data` is check_for_tps's second result."""
    from .line_detection import get_line_image

    current_k = k_factor
    out = []
    for line in line_data:
        mask = np.zeros(image.shape[:2], np.uint8)
        cv2.drawContours(mask, [line["contour"]], -1, (255, 255, 255), -1)
        _, _, _, bbox_h = cv2.boundingRect(line["contour"])
        if line["tps"] is True:
            if line["input_pts"] is None or line["output_pts"] is None:
                raise ValueError("input_pts and output_pts must not be None when tps is True")
            img = run_tps(image, line["output_pts"], line["input_pts"], order=order)
            # run_tps replicates a 1-channel mask to 3 channels (:36-37); get_line_image takes channel 0 of it
            mask = run_tps(mask, line["output_pts"], line["input_pts"], order=order)[..., 0]
            line_img, adapted_k = get_line_image(img, mask, bbox_h, bbox_tolerance=2.0, k_factor=current_k)
            if current_k != adapted_k:
                current_k = adapted_k
        else:
            line_img, _ = get_line_image(image, mask, bbox_h, bbox_tolerance=2.0, k_factor=current_k)
        out.append(line_img)
    return out
