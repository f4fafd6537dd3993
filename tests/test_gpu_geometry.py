"""a-10, GPU half: the connected-component census of the line mask against the host code it replaces
-- cv2.boundingRect, the per-slice cv2.findContours census of get_line_threshold (oracle/geometry.line_threshold
restates BDRC/line_detection.py:362-421 slice by slice) and cv2.connectedComponentsWithStats."""
import cv2
import numpy as np
import pytest
import torch

from ocr_inference_b200 import line_detection as ld
from ocr_inference_b200 import synthetic as synth
from oracle import geometry as og

pytestmark = pytest.mark.gpu


def _masks():
    rng = np.random.default_rng(21)
    out = []
    page, boxes, _ = synth.text_page(3003)
    out.append(("bands", synth.band_mask(boxes, 2048, 3072, grow=2, thin=0.45)))
    # ragged blobs with holes, specks, diagonal links, blobs touching the image border
    m = np.zeros((900, 1400), np.uint8)
    for _ in range(40):
        c = (int(rng.integers(0, 1400)), int(rng.integers(0, 900)))
        cv2.ellipse(m, c, (int(rng.integers(5, 160)), int(rng.integers(3, 40))), float(rng.integers(0, 180)), 0, 360, 255, -1)
    for _ in range(60):
        c = (int(rng.integers(0, 1400)), int(rng.integers(0, 900)))
        cv2.circle(m, c, int(rng.integers(1, 9)), 0, -1)                       # holes
    m[(rng.random(m.shape) < 0.002)] = 255                                    # specks
    for i in range(200):                                                       # a diagonal chain: 8-connected only
        m[100 + i, 50 + i] = 255
    out.append(("blobs", m))
    noise = (cv2.GaussianBlur((rng.random((700, 1111)) * 255).astype(np.uint8), (0, 0), 2.5) > 131).astype(np.uint8) * 255
    out.append(("noise", noise))
    out.append(("empty", np.zeros((300, 500), np.uint8)))
    one = np.zeros((300, 500), np.uint8)
    one[100:120, 200:230] = 255                                                # narrower than two slices: x_steps = 0
    out.append(("narrow", one))
    full = np.full((256, 640), 255, np.uint8)
    full[50:60, 100:300] = 0                                                   # a hole in a mask that fills the frame
    out.append(("full", full))
    tall = np.zeros((2600, 800), np.uint8)
    for y in range(40, 2560, 90):
        cv2.line(tall, (30, y), (770, y + int(rng.integers(-8, 9))), 255, 14)
    out.append(("tall", tall))
    return out


@pytest.mark.parametrize("name,mask", _masks(), ids=[n for n, _ in _masks()])
def test_census_equals_host_census(engine, name, mask):
    buf = engine.mask_census(torch.from_numpy(mask).cuda()).cpu().numpy()
    cen = engine.parse_census(buf, mask.shape[0])
    assert not cen["overflow"]
    assert cen["bbox"] == tuple(cv2.boundingRect(mask)) or (cen["bbox"][2] == 0 and not mask.any())
    assert np.array_equal(cen["row_any"] != 0, mask.max(axis=1) != 0)
    want = og.line_threshold(mask)                     # the reference's slice-by-slice findContours census
    got = ld.threshold_from_census(cen)
    assert got == want, (name, got, want, cen["n"], cen["x_steps"])
    # and through the product's own host path (strip image, one findContours call)
    assert ld.get_line_threshold(mask) == want
    m3 = mask.view(ld.Mask3)
    m3.census = cen
    assert ld.get_line_threshold(m3) == want
    # the slice census itself: contour count and centres of the fullest slice
    x, y, w, h = cv2.boundingRect(mask)
    steps = (w // 20) // 2
    if steps:
        counts, cents = [], []
        for step in range(1, steps + 1):
            x0 = x + steps * step
            sl = np.ascontiguousarray(mask[y:y + h, x0:x0 + 20])
            found, _ = cv2.findContours(sl, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE) if sl.size else ((), None)
            counts.append(len(found))
            cents.append(sorted(cv2.boundingRect(c)[1] + cv2.boundingRect(c)[3] // 2 for c in found))
        best = int(np.argmax(counts))
        assert cen["best"] == best and cen["n"] == counts[best]
        assert sorted(cen["centers"].tolist()) == cents[best]


@pytest.mark.parametrize("name,mask", _masks()[:4], ids=[n for n, _ in _masks()[:4]])
def test_components_equal_cv2(engine, name, mask):
    got = engine.mask_components(torch.from_numpy(mask).cuda(), max_comp=1 << 16)
    n, _, stats, _ = cv2.connectedComponentsWithStats(mask, connectivity=8)
    want = sorted((int(s[0]), int(s[1]), int(s[0] + s[2] - 1), int(s[1] + s[3] - 1), int(s[4])) for s in stats[1:])
    assert sorted(tuple(int(v) for v in r[:5]) for r in got) == want
    assert len(got) == n - 1


def test_page_geometry_with_census_equals_without(engine):
    """The staged pipeline hands page_geometry the GPU census: same lines, boxes, order, ROI masks."""
    page, boxes, _ = synth.text_page(3004)
    rng = np.random.default_rng(2)
    mask = synth.band_mask(boxes, 2048, 3072, grow=2, thin=0.45)
    mask[(rng.random(mask.shape) < 0.0005)] = 255
    cen = engine.parse_census(engine.mask_census(torch.from_numpy(mask).cuda()).cpu().numpy(), mask.shape[0])
    a = ld.page_geometry(mask, True)
    b = ld.page_geometry(mask, True, census=cen)
    assert a[0] == b[0] == "ok" and a[1] == b[1]
    assert [(l[2], l[3]) for l in a[2]] == [(l[2], l[3]) for l in b[2]]
    assert all(np.array_equal(x[1], y[1]) for x, y in zip(a[2], b[2]))
    assert a[3] == b[3] and all(np.array_equal(x, y) for x, y in zip(a[4], b[4])) and a[5] == b[5]
