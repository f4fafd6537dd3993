"""Pins oracle/ against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import os

import cv2
import numpy as np
import pytest

from ocr_inference_b200 import synthetic
from oracle import geometry as og
from oracle import hostops as ho

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
PAGES = {"plain": (600, 700), "upsize": (300, 900), "tall": (2300, 1000), "wide": (700, 4300)}


def _logits(low):
    return np.stack([[cv2.resize(c, (512, 512), interpolation=cv2.INTER_CUBIC) for c in l] for l in low])


@pytest.mark.parametrize("name", list(PAGES))
@pytest.mark.parametrize("closed_form", [True, False])
def test_seg_preprocess_golden(name, closed_form):
    H, W = PAGES[name]
    page, _ = synthetic.page(100 + H, H, W)
    tiles, meta = ho.seg_preprocess(page, closed_form=closed_form)
    H0, W0, n, ys, px, py = G[f"pre_{name}_meta"]
    assert (tiles.shape[0], meta["y_steps"], meta["pad_x"], meta["pad_y"]) == (n, ys, px, py)
    ref = np.unpackbits(G[f"pre_{name}_tiles"])[: tiles.size].reshape(tiles.shape)
    assert np.array_equal(tiles, ref)


@pytest.mark.parametrize("name", list(PAGES))
@pytest.mark.parametrize("closed_form", [True, False])
def test_mask_postprocess_golden(name, closed_form):
    H, W = PAGES[name]
    page, _ = synthetic.page(100 + H, H, W)
    _, meta = ho.seg_preprocess(page, closed_form=False)
    mask = ho.line_postprocess(_logits(G[f"post_{name}_low"]), meta, 0.9, closed_form=closed_form)
    ref = np.unpackbits(G[f"post_{name}_mask"])[: mask.size].reshape(mask.shape) * 255
    assert mask.shape == (H, W) and np.array_equal(mask, ref)
    m5 = ho.layout_postprocess(_logits(G[f"lay_{name}_low"]), meta, 0.8, closed_form=closed_form)
    ref5 = np.unpackbits(G[f"lay_{name}_mask"])[: m5.size].reshape(m5.shape) * 255
    assert m5.shape == (H, W, 5) and np.array_equal(m5, ref5)


@pytest.mark.parametrize("i", range(5))
@pytest.mark.parametrize("closed_form", [True, False])
def test_ocr_line_prep_golden(i, closed_form):
    h, w = G[f"prep_{i}_hw"]
    crop = synthetic.line_crop(500 + i, int(h), int(w))
    for key, pre in (("plain", False), ("prepad", True)):
        x = ho.prepare_ocr_line(crop, 80, 2000, pre_pad=pre, closed_form=closed_form)
        ref = np.unpackbits(G[f"prep_{i}_{key}"])[: 80 * 2000].reshape(80, 2000).astype(np.float32) * 2 - 1
        assert x.shape == (1, 80, 2000) and x.dtype == np.float32
        assert np.array_equal(x[0], ref)


def test_line_geometry_and_crops_golden():
    page, boxes = synthetic.page(700, 800, 1200)
    mask = synthetic.band_mask(boxes, 800, 1200)
    rot_mask, lines, imgs, angle = og.page_to_lines(page, mask, 2.0, 2.5)
    assert angle == G["lines_angle_thr"][0]
    assert og.line_threshold(rot_mask) == G["lines_angle_thr"][1]
    assert np.array_equal(np.array([ln.center for ln in lines]), G["lines_centers"])
    assert np.array_equal(np.array([ln.bbox for ln in lines]), G["lines_bboxes"])
    assert np.array_equal(np.array([im.shape for im in imgs]), G["lines_shapes"])
    assert np.array_equal(np.array([int(im.astype(np.int64).sum()) for im in imgs]), G["lines_sums"])
    assert np.array_equal(imgs[0], G["lines_first"])


def test_ctc_conventions_golden():
    ids = ho.ctc_greedy_ids(G["ctc_logits"], 77)
    text = "".join(synthetic.CHARSET_76[i - 1] for i in ids)
    assert text == str(G["ctc_text"])
    assert [synthetic.CHARSET_76.index(c) + 1 for c in "abc§Z9"] == G["ctc_encode"].tolist()


def test_sigmoid_threshold_in_logit_domain():
    t = ho.sigmoid_logit_threshold(0.9)
    assert t == np.float32(2.1972249)
    below = np.nextafter(t, np.float32(-np.inf))
    assert not ho.sigmoid_ref(below) > np.float32(0.9) and ho.sigmoid_ref(t) > np.float32(0.9)
