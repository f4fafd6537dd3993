"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the host geometry between the two networks and the
page -> text orchestration (BDRC/line_detection.py:107-264, 362-634; BDRC/inference.py:508-666).

Restated compactly with the reference's own cv2 calls; pinned against the real reference by
tests/test_oracle_vs_reference.py when /root/reference is present.
"""
from __future__ import annotations

from dataclasses import dataclass

import cv2
import numpy as np

from . import hostops as ho


@dataclass
class OLine:
    contour: np.ndarray
    bbox: tuple[int, int, int, int]
    center: tuple[int, int]


def rotation_angle(mask: np.ndarray, max_angle: float = 5.0) -> float:
    """get_rotation_angle_from_lines (line_detection.py:107-142)."""
    cs, _ = cv2.findContours(mask, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    cs = [c for c in cs if cv2.contourArea(c) > mask.shape[0] * mask.shape[1] * 0.001]
    ang = [cv2.minAreaRect(c)[2] for c in cs]
    lo = [a for a in ang if abs(a) != 0.0 and a < max_angle]
    hi = [a for a in ang if abs(a) != 90.0 and a > (90 - max_angle)]
    if len(lo) > len(hi) and len(lo) > 0:
        return np.mean(lo)
    if len(hi) > 0:
        return -(90 - np.mean(hi))
    return 0


def build_raw_line_data(image, mask, rot_threshold: float = 0.5):
    """line_detection.py:212-244."""
    if mask.ndim == 3:
        mask = cv2.cvtColor(mask, cv2.COLOR_BGR2GRAY)
    angle = rotation_angle(mask)
    if angle > rot_threshold:
        mask = ho.rotate_from_angle_ref(mask, angle)
        image = ho.rotate_from_angle_ref(image, angle)
    cs, _ = cv2.findContours(mask, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    cs = [c for c in cs if cv2.contourArea(c) > 10]
    return image, cv2.cvtColor(mask, cv2.COLOR_GRAY2RGB), cs, angle


def filter_line_contours(image, contours, threshold: float = 0.01):
    """line_detection.py:247-264."""
    return [c for c in contours if cv2.boundingRect(c)[2] > image.shape[1] * threshold and cv2.boundingRect(c)[3] > 10]


def build_line(contour) -> OLine:
    """build_line_data (line_detection.py:188-209) without the guid."""
    contour = cv2.approxPolyDP(contour, 0.001 * cv2.arcLength(contour, True), True)
    x, y, w, h = cv2.boundingRect(contour)
    return OLine(contour, (x, y, w, h), (x + w // 2, y + h // 2))


def line_threshold(pred: np.ndarray, slice_width: int = 20) -> float:
    """get_line_threshold (line_detection.py:362-421)."""
    if pred.ndim == 3:
        pred = cv2.cvtColor(pred, cv2.COLOR_BGR2GRAY)
    x, y, w, h = cv2.boundingRect(pred)
    steps = (w // slice_width) // 2
    counts = []
    for step in range(1, steps + 1):
        xs = x + steps * step
        cs, _ = cv2.findContours(pred[y:y + h, xs:xs + slice_width], cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
        counts.append((len(cs), cs))
    if not counts:
        return 0.0
    n, cs = sorted(counts, key=lambda t: t[0], reverse=True)[0]
    if n == 0:
        return 0.0
    ys = [cv2.boundingRect(c)[1] + cv2.boundingRect(c)[3] // 2 for c in cs]
    return float(np.median(ys) // n) if ys else 0.0


def sort_centers(centers, thr: float):
    """sort_bbox_centers (line_detection.py:424-486)."""
    if not centers:
        return []
    rows, tmp = [], []
    for c in centers:
        if tmp and abs(np.mean([p[1] for p in tmp]) - c[1]) > thr:
            tmp.sort(key=lambda p: p[0])
            rows.append(tmp.copy())
            tmp.clear()
        tmp.append(c)
    if tmp:
        rows.append(tmp)
    for r in rows:
        r.sort(key=lambda p: p[0])
    return list(reversed(rows))


def sort_lines(mask, lines: list[OLine], group_lines: bool = True):
    """sort_lines_by_threshold2 + group_line_chunks (line_detection.py:489-548, 594-634)."""
    thr = line_threshold(mask)
    rows = sort_centers([ln.center for ln in lines], thr)
    out = []
    for row in rows:
        if group_lines and len(row) > 1:
            stack = []
            for c in row:
                for ln in lines:
                    if c == ln.center:
                        stack.append(ln.contour)
                        break
            hull = cv2.convexHull(np.vstack(stack))
            x, y, w, h = cv2.boundingRect(hull)
            out.append(OLine(hull, (x, y, w, h), (x + w // 2, y + h // 2)))
        elif group_lines:
            for c in row:
                for ln in lines:
                    if c == ln.center:
                        out.append(ln)
                        break
        else:
            for c in row:
                for ln in lines:
                    if c == ln.center:
                        out.append(ln)
    return out, thr


def page_to_lines(image, line_mask, k_factor: float = 2.5, bbox_tolerance: float = 4.0, merge_lines: bool = True):
    """Stages 2-4 of OCRPipeline.run_ocr (inference.py:758-797) with use_tps=False.
    Returns (rot_mask, sorted OLines, line images, angle) or a failure string."""
    rot_img, rot_mask, contours, angle = build_raw_line_data(image, line_mask)
    if len(contours) == 0:
        return "No lines detected"
    filt = filter_line_contours(rot_mask, contours)
    if len(filt) == 0:
        return "No valid lines after filtering"
    lines, _ = sort_lines(rot_mask, [build_line(c) for c in filt], merge_lines)
    imgs = ho.extract_line_images_ref(rot_img, [ln.contour for ln in lines], k_factor, bbox_tolerance)
    if not imgs:
        return "No valid line images extracted"
    return rot_mask, lines, imgs, float(angle)


def recognise(line_images, ocr_weights, charset, Ht: int = 80, Wt: int = 2000, pre_pad: bool = False):
    """run_text_recognition (inference.py:618-666) with greedy decode; returns (texts, id lists)."""
    from . import nets

    V = len(charset) + 1
    texts, ids = [], []
    for img in line_images:
        if img.shape[0] == 0 or img.shape[1] == 0:
            continue
        x = ho.prepare_ocr_line(img, Ht, Wt, pre_pad=pre_pad, closed_form=False)
        logits = np.squeeze(nets.easter2_forward(ocr_weights, x))
        seq = ho.ctc_greedy_ids(logits, V)
        ids.append(seq)
        texts.append(ho.ctc_ids_to_text(seq, charset))
    return texts, ids
