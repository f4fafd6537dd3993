"""Drop-in for BDRC/line_detection.py.

Host geometry (contours, rotation angle, reading-order sort: tens of contours per page) stays on
the CPU with cv2, as SURVEY.md 8(a)-10 scopes it; everything that touches page-sized pixel data
-- rotation (cv2.warpAffine), dilation, masking, row/column deletion and crop gathering -- runs
in libbdrc_b200.so on the GPU.  Function names, arguments and return values follow the reference.
"""
from __future__ import annotations

import math
from typing import Sequence
from uuid import uuid1

import cv2
import numpy as np
import torch

from .data import BBox, Line

# ------------------------------------------------------------------------------------------------
# small host helpers (BDRC/line_detection.py:20-59)
# ------------------------------------------------------------------------------------------------


def generate_guid(clock_seq: int):
    return uuid1(clock_seq=clock_seq)


def get_contours(image: np.ndarray):
    found, _ = cv2.findContours(image, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    return found


class RowCompaction:
    """The rows of a mask that hold any foreground, each run of them followed by one empty separator row.

    cv2.findContours scans every pixel, and on a text page most rows are empty.  Border following only looks at
    8-neighbours, runs of occupied rows stay separated by an empty row, and dropping empty rows keeps the raster
    order of the start points, so the contours of the compacted mask are the contours of the page -- same
    points, same list order -- once their y coordinates go back through `ymap`
    (tests/test_cpu_host_logic.py checks this against cv2 on the full mask).  `small` is None when compaction
    would not pay (most rows occupied)."""

    def __init__(self, mask: np.ndarray, worth: float = 0.6, occ: np.ndarray | None = None):
        """`occ` (optional): per-row "holds foreground" flags already computed on the GPU (bdrc_mask_census)."""
        self.small = self.ymap = None
        H = int(mask.shape[0])
        occ = (mask.max(axis=1) != 0) if occ is None or len(occ) != H else (np.asarray(occ) != 0)
        n_occ = int(np.count_nonzero(occ))
        self.empty = n_occ == 0
        if self.empty or H < 64:
            return
        d = np.diff(occ.astype(np.int8))
        starts = list(np.flatnonzero(d == 1) + 1)
        ends = list(np.flatnonzero(d == -1) + 1)
        if occ[0]:
            starts.insert(0, 0)
        if occ[-1]:
            ends.append(H)
        self.y0, self.y1 = int(starts[0]), int(ends[-1])          # first occupied row, one past the last
        if n_occ + len(starts) > worth * H:
            return
        small = np.zeros((n_occ + len(starts), mask.shape[1]), mask.dtype)
        ymap = np.zeros(small.shape[0], np.int32)
        o = 0
        for a, b in zip(starts, ends):
            small[o:o + b - a] = mask[a:b]
            ymap[o:o + b - a] = np.arange(a, b, dtype=np.int32)
            ymap[o + b - a] = b                                    # separator row (never on a contour)
            o += b - a + 1
        self.small, self.ymap = small, ymap

    def contours(self):
        found, _ = cv2.findContours(self.small, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
        if len(found) < 8:
            for c in found:
                c[:, 0, 1] = self.ymap[c[:, 0, 1]]
            return found
        # many small contours: one gather over all points, then views of it (no per-contour numpy round trip)
        pts = np.concatenate(found, axis=0)
        pts[:, 0, 1] = self.ymap[pts[:, 0, 1]]
        out, o = [], 0
        for c in found:
            out.append(pts[o:o + len(c)])
            o += len(c)
        return tuple(out)


def get_contours_compact(mask: np.ndarray, rows: "RowCompaction | None" = None):
    """get_contours(mask) through the row compaction (identical result, a fraction of the pixels scanned)."""
    if mask.ndim != 2 or mask.dtype != np.uint8:
        return get_contours(mask)
    rows = rows or RowCompaction(mask)
    if rows.empty:
        return ()
    return rows.contours() if rows.small is not None else get_contours(mask)


def optimize_countour(cnt, e: float = 0.001):
    return cv2.approxPolyDP(cnt, e * cv2.arcLength(cnt, True), True)


def _engine():
    from .inference import get_engine

    return get_engine()


# ------------------------------------------------------------------------------------------------
# a-4 rotation (BDRC/line_detection.py:62-76, 107-142, 212-244)
# ------------------------------------------------------------------------------------------------


def rotate_from_angle(image: np.ndarray, angle: float) -> np.ndarray:
    """cv2.warpAffine about the centre (bilinear, black border, same size) on the GPU."""
    rows, cols = image.shape[:2]
    M = cv2.getRotationMatrix2D((cols / 2, rows / 2), angle, 1)
    eng = _engine()
    src = torch.from_numpy(np.ascontiguousarray(image, dtype=np.uint8)).to(eng.device)
    return eng.warp_affine(src, M).cpu().numpy()


def get_rotation_angle_from_lines(line_mask: np.ndarray, max_angle: float = 5.0, debug_angles: bool = False,
                                  _contours=None, _areas=None) -> float:
    """Mean minAreaRect angle of the big blobs (> 0.1 % of the page), :107-142.  `_contours` (internal):
    the result of get_contours(line_mask) when the caller already has it."""
    found = get_contours(line_mask) if _contours is None else _contours
    areas = [cv2.contourArea(c) for c in found] if _areas is None else _areas
    big = line_mask.shape[0] * line_mask.shape[1] * 0.001
    blobs = [c for c, a in zip(found, areas) if a > big]
    angles = [cv2.minAreaRect(c)[2] for c in blobs]
    low = [a for a in angles if abs(a) != 0.0 and a < max_angle]
    high = [a for a in angles if abs(a) != 90.0 and a > (90 - max_angle)]
    if debug_angles:
        print(f"All Angles: {angles}")
    if len(low) > len(high) and len(low) > 0:
        return np.mean(low)
    if len(high) > 0:
        return -(90 - np.mean(high))
    return 0


class Mask3(np.ndarray):
    """The (H, W, 3) mask build_raw_line_data returns (:235,241 GRAY2RGB), remembering the 1-channel mask it
    replicates: get_line_threshold converts it straight back with BGR2GRAY (:375-376), which is the identity on
    replicated channels (tests/test_cpu_host_logic.py), so the 19 MB conversion is skipped.  Views and copies do
    not inherit the link."""

    gray = None
    rows = None          # RowCompaction of `gray`
    census = None        # parse_census() of the GPU's connected-component census of `gray` (or of this 2-D mask)

    def __array_finalize__(self, obj):
        self.gray = self.rows = self.census = None


def build_raw_line_data(image: np.ndarray, line_mask: np.ndarray, rot_threshold: float = 0.5):
    """:212-244 -> (rotated image, rotated 3-channel mask, contours with area > 10, angle)."""
    census = getattr(line_mask, "census", None)         # a 2-D mask that came with its GPU census (Detection.predict)
    if census is not None and census["overflow"]:
        census = None
    if line_mask.ndim == 3:
        line_mask, census = cv2.cvtColor(line_mask, cv2.COLOR_BGR2GRAY), None
    rows = RowCompaction(line_mask, occ=census["row_any"] if census else None) \
        if line_mask.ndim == 2 and line_mask.dtype == np.uint8 else None
    found = get_contours_compact(line_mask, rows)
    areas = [cv2.contourArea(c) for c in found]
    angle = get_rotation_angle_from_lines(line_mask, _contours=found, _areas=areas)
    if angle > rot_threshold:
        out_mask = rotate_from_angle(line_mask, angle)
        out_img = rotate_from_angle(image, angle)
        rows = RowCompaction(out_mask)
        found = get_contours_compact(out_mask, rows)
        areas = [cv2.contourArea(c) for c in found]
        census = None        # belongs to the unrotated mask
    else:
        # unrotated: the reference runs the identical findContours call a second time (:238); reuse it
        out_mask, out_img = line_mask, image
    contours = [c for c, a in zip(found, areas) if a > 10]
    mask3 = cv2.cvtColor(out_mask, cv2.COLOR_GRAY2RGB).view(Mask3)
    mask3.gray, mask3.rows, mask3.census = np.asarray(out_mask), rows, census
    return out_img, mask3, contours, angle


def page_geometry(line_mask: np.ndarray, merge_lines: bool = True, rot_threshold: float = 0.5, census: dict | None = None):
    """Everything between the line mask and the GPU line extraction that needs no pixels of the page and no GPU
    (build_raw_line_data without rotation, filter_line_contours, build_line_data, sort_lines_by_threshold2 and
    the filled-contour ROI masks of extract_line_images), as ONE picklable call for a geometry worker process
    (geometry_pool.py).  Returns
      ("rotate",)                      the page needs rotating (angle > rot_threshold): GPU work, caller's job
      ("failed", message)              the reference's stage-2 failure messages (inference.py:521-528)
      ("ok", angle, lines, rois, masks, hs)   lines = [(guid, contour, (x, y, w, h), center)] in reading order
    """
    if census is not None and census["overflow"]:
        census = None
    rows = RowCompaction(line_mask, occ=census["row_any"] if census else None)
    found = get_contours_compact(line_mask, rows)
    areas = [cv2.contourArea(c) for c in found]
    angle = get_rotation_angle_from_lines(line_mask, _contours=found, _areas=areas)
    if angle > rot_threshold:
        return ("rotate",)
    contours = [c for c, a in zip(found, areas) if a > 10]
    if len(contours) == 0:
        return ("failed", "No lines detected")
    filtered = filter_line_contours(line_mask, contours)
    if len(filtered) == 0:
        return ("failed", "No valid lines after filtering")
    line_data = [build_line_data(c, _guid=False) for c in filtered]
    linked = line_mask.view(Mask3)          # lets get_line_threshold reuse the row compaction
    linked.gray, linked.rows, linked.census = line_mask, rows, census
    sorted_lines, _ = sort_lines_by_threshold2(linked, line_data, group_lines=merge_lines)
    for ln in sorted_lines:                 # single-chunk rows keep their chunk line: it gets its guid now
        if ln.guid is None:
            ln.guid = generate_guid(clock_seq=23)
    H, W = line_mask.shape[:2]
    rois, masks, hs = line_rois(sorted_lines, H, W)
    lines = [(ln.guid, ln.contour, (ln.bbox.x, ln.bbox.y, ln.bbox.w, ln.bbox.h), ln.center) for ln in sorted_lines]
    return ("ok", angle, lines, rois, masks, hs)


def filter_line_contours(image: np.ndarray, line_contours, threshold: float = 0.01) -> list:
    """:247-264: keep contours wider than `threshold` of the page and taller than 10 px."""
    keep = []
    for c in line_contours:
        _, _, w, h = cv2.boundingRect(c)
        if w > image.shape[1] * threshold and h > 10:
            keep.append(c)
    return keep


def build_line_data(contour: np.ndarray, optimize: bool = True, _guid: bool = True) -> Line:
    """:188-209.  `_guid=False` (internal) leaves the guid to the caller: most chunk lines of a page end up merged
    into a row hull, which gets its own guid, and uuid1 costs as much as the polygon approximation."""
    if optimize:
        contour = optimize_countour(contour)
    x, y, w, h = cv2.boundingRect(contour)
    return Line(generate_guid(clock_seq=23) if _guid else None, contour, BBox(x, y, w, h), (x + (w // 2), y + (h // 2)))


# ------------------------------------------------------------------------------------------------
# a-10 reading order (BDRC/line_detection.py:362-634)
# ------------------------------------------------------------------------------------------------


def parse_census(buf: np.ndarray, H: int, max_centers: int = 4096) -> dict:
    """The buffer Engine.mask_census() fills (bdrc_mask_census: connected-component census of the mask on the GPU),
    as the host sees it.  No GPU needed: geometry worker processes call this."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    hdr = buf[:64].view(np.int32)
    n_c = int(hdr[9])
    return {"bbox": (int(hdr[0]), int(hdr[1]), int(hdr[2]), int(hdr[3])), "x_steps": int(hdr[4]), "best": int(hdr[7]),
            "n": int(hdr[8]), "centers": buf[64:64 + 4 * n_c].view(np.int32).copy(),
            "row_any": buf[64 + 4 * max_centers:64 + 4 * max_centers + H].copy(), "overflow": bool(hdr[10]),
            "slice_width": 20}


def threshold_from_census(census: dict) -> float:
    """get_line_threshold's result from the GPU census: the slices' contour counts and the fullest slice's
    bounding-box centres were computed by connected-component labelling on the device (csrc/geometry.cu)."""
    if census["x_steps"] == 0:
        return 0.0
    if census["n"] == 0 or len(census["centers"]) == 0:
        print("number of contours is 0")
        return 0.0
    return float(np.median(census["centers"]) // census["n"])


def get_line_threshold(line_prediction: np.ndarray, slice_width: int = 20) -> float:
    """:362-421: census of contours in vertical slices; median centre-y of the fullest slice // n."""
    census = getattr(line_prediction, "census", None)
    if census is not None and not census["overflow"] and census["slice_width"] == slice_width:
        return threshold_from_census(census)
    gray = getattr(line_prediction, "gray", None)
    rows = None
    if gray is not None and gray.shape == line_prediction.shape[:2]:
        rows = getattr(line_prediction, "rows", None)
        line_prediction = gray
    elif line_prediction.ndim == 3:
        line_prediction = cv2.cvtColor(line_prediction, cv2.COLOR_BGR2GRAY)
    if rows is None and line_prediction.ndim == 2 and line_prediction.dtype == np.uint8:
        rows = RowCompaction(line_prediction)
    if rows is not None and rows.small is not None:
        # the slices of the compacted mask hold the same contours (rows between y and y + h that are empty
        # carry none); only the top edge of each bounding box has to go back through ymap
        x, _, w, _ = cv2.boundingRect(rows.small)
        y = rows.y0
        src, ymap = rows.small, rows.ymap
    else:
        x, y, w, h = cv2.boundingRect(line_prediction)
        src, ymap = line_prediction[y:y + h], None
    x_steps = (w // slice_width) // 2
    if x_steps == 0:
        return 0.0
    # The reference runs findContours once per vertical slice (up to ~70 calls whose fixed cost dominates).  Here the
    # slices are laid side by side in ONE image, one empty column between neighbours: 8-connected blobs cannot
    # cross an empty column and the column plays the part of the image border, so every slice keeps exactly its
    # contours; a contour is attributed to its slice by the x of its bounding box.
    W_src = src.shape[1]
    pitch = slice_width + 1
    cols = np.full(x_steps * pitch, -1, dtype=np.int64)
    for step in range(1, x_steps + 1):
        x0 = x + x_steps * step
        n_cols = max(0, min(x0 + slice_width, W_src) - x0) if x0 < W_src else 0
        cols[(step - 1) * pitch:(step - 1) * pitch + n_cols] = np.arange(x0, x0 + n_cols)
    strip = np.zeros((src.shape[0], cols.size), dtype=src.dtype)
    valid = cols >= 0
    strip[:, valid] = src[:, cols[valid]]
    found, _ = cv2.findContours(strip, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    if len(found) == 0:
        print("number of contours is 0")
        return 0.0
    boxes = np.array([cv2.boundingRect(c) for c in found], dtype=np.int64)
    which = boxes[:, 0] // pitch
    counts = np.bincount(which, minlength=x_steps)
    best = int(np.argmax(counts))            # the first slice with the highest count (stable sort in the reference)
    n = int(counts[best])
    centers = []
    for _, cy, _, ch in boxes[which == best]:
        cy = int(cy)
        if ymap is not None:
            cy = int(ymap[cy]) - y
        centers.append(cy + (int(ch) // 2))
    return float(np.median(centers) // n) if centers else 0.0


def sort_bbox_centers(bbox_centers: list[tuple[int, int]], line_threshold: float = 20) -> list:
    """:424-486: greedy grouping by the running mean y of the current row; rows sorted by x;
    the row list is returned reversed."""
    if not bbox_centers:
        return []
    rows, cur, ysum = [], [], 0
    for c in bbox_centers:
        # running integer sum: ysum / len(cur) is the same float64 as np.mean of the ys (exact sums below 2**53)
        if cur and abs(ysum / len(cur) - c[1]) > line_threshold:
            cur.sort(key=lambda p: p[0])
            rows.append(cur.copy())
            cur.clear()
            ysum = 0
        cur.append(c)
        ysum += int(c[1])
    if cur:
        rows.append(cur)
    for r in rows:
        r.sort(key=lambda p: p[0])
    return list(reversed(rows))


def group_line_chunks(sorted_bbox_centers, lines: list[Line], adaptive_grouping: bool = True) -> list[Line]:
    """:489-548: rows with several chunks are merged into the convex hull of their contours."""
    out = []
    first = {}
    for ln in lines:              # the reference scans `lines` for the first one with this centre (:505-512)
        first.setdefault(ln.center, ln)
    for centers in sorted_bbox_centers:
        if len(centers) > 1:
            stack = [first[c].contour for c in centers if c in first]
            hull = cv2.convexHull(np.vstack(stack))
            x, y, w, h = cv2.boundingRect(hull)
            out.append(Line(generate_guid(clock_seq=23), hull, BBox(x, y, w, h), (x + (w // 2), y + (h // 2))))
        else:
            out.extend(first[c] for c in centers if c in first)
    return out


def sort_lines_by_threshold2(line_mask: np.ndarray, lines: list[Line], threshold: int = 20,
                             calculate_threshold: bool = True, group_lines: bool = True):
    """:594-634."""
    thr = get_line_threshold(line_mask) if calculate_threshold else threshold
    rows = sort_bbox_centers([ln.center for ln in lines], line_threshold=thr)
    if group_lines:
        return group_line_chunks(rows, lines), thr
    out = []
    for c in [p for r in rows for p in r]:
        for ln in lines:
            if c == ln.center:
                out.append(ln)
    return out, thr


sort_lines_by_threshold = sort_lines_by_threshold2   # :551-591 is the same algorithm


# ------------------------------------------------------------------------------------------------
# a-5 per-line masked crops (BDRC/line_detection.py:79-104, 267-359) on the GPU
# ------------------------------------------------------------------------------------------------


class LineImages(list):
    """list[np.ndarray] of crops that also remembers the packed device copy, so that
    OCRPipeline.run_text_recognition does not upload them again."""

    device_pack = None


class DeviceCrop:
    """A line crop that stays on the GPU (run_ocr's internal fast path): has .shape like the ndarray the
    reference returns and turns into one (one D2H copy) when numpy asks for it."""

    def __init__(self, crops_dev: torch.Tensor, offset: int, shape: tuple[int, int, int]):
        self._dev, self._off, self.shape = crops_dev, offset, shape
        self.ndim, self.dtype = 3, np.dtype(np.uint8)

    def __array__(self, dtype=None, copy=None):
        h, w, c = self.shape
        a = self._dev[self._off:self._off + h * w * c].cpu().numpy().reshape(h, w, c)
        return a.astype(dtype) if dtype is not None else a


def _roi_mask(contour: np.ndarray, H: int, W: int):
    """Filled contour rasterised into its (page-clipped) bounding box, as cv2.drawContours with
    thickness=-1 does on the full page (line_detection.py:349-351)."""
    x, y, w, h = cv2.boundingRect(contour)
    x0, y0 = max(x, 0), max(y, 0)
    x1, y1 = min(x + w, W), min(y + h, H)
    if x1 <= x0 or y1 <= y0:
        return (x0, y0, 0, 0), np.zeros((0, 0), np.uint8), h
    m = np.zeros((y1 - y0, x1 - x0), np.uint8)
    cv2.drawContours(m, [contour], -1, 255, -1, offset=(-x0, -y0))
    return (x0, y0, x1 - x0, y1 - y0), m, h


def _se_for(bbox_h: int, k: float):
    """Structuring element of extract_line (:280-283); None where cv2 would raise."""
    ks = int(bbox_h * k)
    kh = int(ks * k)
    return (ks, kh) if ks >= 1 and kh >= 1 else None


def _extract_on_device(page_dev: torch.Tensor, rois, masks, bbox_hs, default_k: float, bbox_tolerance: float):
    """The shrink-k search of get_line_image/extract_line_images (:291-359) for all lines of a page.

    The candidate factors form ONE sequence k0, k0-0.1, ... shared by all lines (the adapted factor
    carries over to the next line), so crop heights are evaluated on the GPU for a window of
    candidates per line and the sequential decision is replayed on the host from those numbers.
    Returns (chosen (kw,kh) or None per line, crops ...).
    """
    eng = _engine()
    L = len(rois)
    H, W, C = int(page_dev.shape[0]), int(page_dev.shape[1]), int(page_dev.shape[2]) if page_dev.dim() == 3 else 1
    kseq = [default_k]
    while len(kseq) < 64:
        kseq.append(kseq[-1] - 0.1)
    valid = [i for i in range(L) if rois[i][2] > 0 and rois[i][3] > 0]
    pack = eng.pack_masks([masks[i] for i in valid], [rois[i] for i in valid]) if valid else None
    vpos = {i: p for p, i in enumerate(valid)}
    heights: dict[tuple[int, int], int] = {}
    NC = 11

    def fetch(from_line: int, cur: int):
        todo = [i for i in valid if i >= from_line]
        if not todo:
            return
        cand = []
        for i in todo:
            row = []
            for j in range(cur, cur + NC):
                se = _se_for(bbox_hs[i], kseq[j]) if j < len(kseq) else None
                row.append(se if se else (1, 1))
            cand.append(row)
        hs = eng.line_crop_heights(page_dev, pack, [vpos[i] for i in todo], cand)
        for r, i in enumerate(todo):
            for c, j in enumerate(range(cur, cur + NC)):
                heights[(i, j)] = int(hs[r, c])

    def height(i: int, j: int):
        """Crop height for line i at candidate j, or raises ValueError where cv2 would."""
        if _se_for(bbox_hs[i], kseq[j]) is None:
            raise ValueError("empty structuring element")
        if i not in vpos:
            return 0
        if (i, j) not in heights:
            fetch(i, j)
        return heights[(i, j)]

    chosen: list[tuple[int, int] | None] = [None] * L
    cur = 0
    for i in range(L):
        j = cur
        try:
            hgt = height(i, j)
            pick, attempts = j, 0
            while hgt > bbox_hs[i] * bbox_tolerance and attempts < 10:
                j += 1
                if kseq[j] <= 0.1:
                    break
                hgt = height(i, j)
                pick = j
                attempts += 1
            chosen[i] = _se_for(bbox_hs[i], kseq[pick])
            if kseq[cur] != kseq[j]:
                cur = j
        except ValueError as e:   # the reference prints and returns a black fallback image (:322-327)
            print(f"Error in get_line_image: {e}")
            chosen[i] = None
    real = [i for i in range(L) if chosen[i] is not None and i in vpos]
    crops_dev = hw = off = None
    if real:
        crops_dev, hw, off = eng.extract_lines(page_dev, pack, [vpos[i] for i in real], [chosen[i] for i in real])
    return chosen, real, crops_dev, hw, off, C


def line_rois(line_data: list[Line], H: int, W: int):
    """Filled-contour ROI masks of the lines (the host half of extract_line_images): (rois, masks, bbox heights)."""
    rois, masks, hs = [], [], []
    for ln in line_data:
        roi, m, h = _roi_mask(ln.contour, H, W)
        rois.append(roi)
        masks.append(m)
        hs.append(h)
    return rois, masks, hs


def extract_line_images(image: np.ndarray, line_data: list[Line], default_k: float = 1.7,
                        bbox_tolerance: float = 3, page_dev: torch.Tensor | None = None,
                        download: bool = True, _rois=None) -> LineImages:
    """:330-359.  `page_dev` (optional) is the same image already resident on the GPU; with
    download=False the crops stay on the device (DeviceCrop placeholders) for the recogniser.
    `_rois` (internal): line_rois(line_data, H, W) when a geometry worker already rasterised them."""
    eng = _engine()
    if image.ndim == 2:
        image = image[..., None]
    if page_dev is None:
        page_dev = torch.from_numpy(np.ascontiguousarray(image, dtype=np.uint8)).to(eng.device)
    H, W = image.shape[:2]
    rois, masks, hs = _rois if _rois is not None else line_rois(line_data, H, W)
    chosen, real, crops_dev, hw, off, C = _extract_on_device(page_dev, rois, masks, hs, default_k, bbox_tolerance)
    out = LineImages()
    host = {}
    if real:
        hw_h, off_h = hw.cpu().numpy(), off.cpu().numpy()
        total = int(off_h[-1] + ((int(hw_h[-1, 0]) * int(hw_h[-1, 1]) * C + 15) // 16) * 16)
        if download:
            buf = crops_dev[:max(total, 1)].cpu().numpy()
        for p, i in enumerate(real):
            h, w = int(hw_h[p, 0]), int(hw_h[p, 1])
            host[i] = (buf[off_h[p]:off_h[p] + h * w * C].reshape(h, w, C).copy() if download
                       else DeviceCrop(crops_dev, int(off_h[p]), (h, w, C)))
        out.device_pack = (crops_dev, hw, off, real)
    for i in range(len(line_data)):
        if i in host:
            out.append(host[i])
        elif chosen[i] is None:
            out.append(np.zeros((hs[i], hs[i] * 2, 3), np.uint8))     # fallback of :325-327
        else:
            out.append(np.zeros((0, 0, C), np.uint8))                 # nothing left after mask_n_crop
    return out


def mask_n_crop(image: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """:79-104 on the GPU: AND with the mask, delete rows/cols where any channel is all zero."""
    eng = _engine()
    image = image.astype(np.uint8)
    if image.ndim == 2:
        image = image[..., None]
    C = image.shape[2]
    ys, xs = np.nonzero(mask)
    if len(ys) == 0:
        return np.zeros((0, 0, C), np.uint8)
    y0, y1, x0, x1 = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
    page_dev = torch.from_numpy(np.ascontiguousarray(image)).to(eng.device)
    pack = eng.pack_masks([np.ascontiguousarray(mask[y0:y1, x0:x1]).astype(np.uint8)], [(x0, y0, x1 - x0, y1 - y0)])
    crops_dev, hw, off = eng.extract_lines(page_dev, pack, [0], [(1, 1)])
    h, w = (int(v) for v in hw.cpu().numpy()[0])
    return crops_dev[: h * w * C].cpu().numpy().reshape(h, w, C).copy()


def extract_line(image: np.ndarray, mask: np.ndarray, bbox_h: int, k_factor: float = 1.2):
    """:267-288 for one full-page mask."""
    se = _se_for(bbox_h, k_factor)
    if se is None:
        raise ValueError("empty structuring element")
    eng = _engine()
    if image.ndim == 2:
        image = image[..., None]
    C = image.shape[2]
    ys, xs = np.nonzero(mask)
    if len(ys) == 0:
        return np.zeros((0, 0, C), np.uint8)
    y0, y1, x0, x1 = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
    page_dev = torch.from_numpy(np.ascontiguousarray(image, dtype=np.uint8)).to(eng.device)
    pack = eng.pack_masks([np.ascontiguousarray(mask[y0:y1, x0:x1]).astype(np.uint8)], [(x0, y0, x1 - x0, y1 - y0)])
    crops_dev, hw, off = eng.extract_lines(page_dev, pack, [0], [se])
    h, w = (int(v) for v in hw.cpu().numpy()[0])
    return crops_dev[: h * w * C].cpu().numpy().reshape(h, w, C).copy()


def get_line_image(image: np.ndarray, mask: np.ndarray, bbox_h: int, bbox_tolerance: float = 2.5,
                   k_factor: float = 1.2):
    """:291-327 for one full-page mask -> (line image, adapted k)."""
    try:
        k = k_factor
        img = extract_line(image, mask, bbox_h, k)
        attempts = 0
        while img.shape[0] > bbox_h * bbox_tolerance and attempts < 10:
            k = k - 0.1
            if k <= 0.1:
                break
            img = extract_line(image, mask, bbox_h, k)
            attempts += 1
        return img, k
    except Exception as e:  # noqa: BLE001
        print(f"Error in get_line_image: {e}")
        return np.zeros((bbox_h, bbox_h * 2, 3), dtype=np.uint8), k_factor
