"""Whole-page parity: the GPU pipeline with its OWN line mask (no injection) against the CPU oracle's
page -> text results (tests/golden/page_oracle.json, produced by tests/golden/make_page_oracle.py), on 32
synthetic text pages at BASELINE size (2048x3072), for the reference's default run_ocr parameters and the
Demo-OCR notebook's.  Target (BASELINE.json): >= 95 % identical decoded lines.

What may differ, and why: the GPU network runs bf16 operands with fp32 accumulation, so mask pixels whose fp32
logit lies within the bf16 tolerance of the threshold can flip; a flipped pixel on a blob's outline moves a
contour vertex, hence a line box, by a pixel or two, and the crop with it.  Line COUNT and reading order do not
depend on such pixels, and a recogniser that reads its input is not sensitive to a one-pixel shift: that is what
the >= 95 % bar measures.  Easter2's architecture is our restatement (parity unpinned against BDRC's ONNX)."""
import json
import os

import numpy as np
import pytest

from ocr_inference_b200 import synthetic as synth
from ocr_inference_b200.data import CharsetEncoder, LineDetectionConfig, OCRModelConfig, OpStatus
from ocr_inference_b200.synthetic_weights import easter2_weights_fitted, seg_weights_fitted

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "page_oracle.json")


def _fixture():
    with open(FIXTURE) as f:
        return json.load(f)


def _pipe(engine):
    from ocr_inference_b200.inference import OCRPipeline

    return OCRPipeline(OCRModelConfig(easter2_weights_fitted(), "Easter2", 2000, 80, "input", "output", True, False,
                                      CharsetEncoder.STACK, synth.CHARSET_76, True, "standin"),
                       LineDetectionConfig("", seg_weights_fitted(11), "deeplabv3plus-r34", 512, ["line"]), None,
                       engine=engine)


def _compare(results, pages, key):
    n_lines = same = same_box = max_box = count_ok = truth_ok = 0
    for (st, res), rec in zip(results, pages):
        want = rec[key]
        assert st == OpStatus.SUCCESS and "failed" not in want, (rec["seed"], res if st != OpStatus.SUCCESS else want)
        got_txt = [l.text for l in res[2]]
        got_box = [[l.bbox.x, l.bbox.y, l.bbox.w, l.bbox.h] for l in res[1]]
        n_lines += len(want["texts"])
        if len(got_txt) != len(want["texts"]):
            continue
        count_ok += 1
        same += sum(a == b for a, b in zip(got_txt, want["texts"]))
        truth_ok += sum(a == b for a, b in zip(got_txt, rec["truth"])) if len(rec["truth"]) == len(got_txt) else 0
        for a, b in zip(got_box, want["boxes"]):
            same_box += int(a == b)
            max_box = max(max_box, max(abs(p - q) for p, q in zip(a, b)))
    return dict(lines=n_lines, identical=same, identical_boxes=same_box, max_box_diff=max_box, pages_equal_count=count_ok,
                equal_to_truth=truth_ok)


@pytest.mark.gpu
def test_whole_pages_with_the_gpu_mask_match_the_cpu_oracle(engine):
    fx = _fixture()
    pages_fx = fx["pages"]
    assert len(pages_fx) >= 32
    pages = [synth.text_page(rec["seed"])[0] for rec in pages_fx]
    pipe = _pipe(engine)
    try:
        for key, (kf, tol) in fx["settings"].items():
            res = pipe.run_ocr_batch(pages, k_factor=kf, bbox_tolerance=tol)
            s = _compare(res, pages_fx, key)
            print(f"[{key}] {s}")
            assert s["pages_equal_count"] == len(pages_fx), "line count / reading order must not depend on tie pixels"
            assert s["identical"] >= 0.95 * s["lines"], s
            assert s["max_box_diff"] <= 8, s
            assert s["equal_to_truth"] >= 0.9 * s["lines"], s
        # the page-by-page entry point gives the batch's results
        kf, tol = fx["settings"]["default"]
        batch = pipe.run_ocr_batch(pages[:4], k_factor=kf, bbox_tolerance=tol)
        for page, b in zip(pages[:4], batch):
            st, r = pipe.run_ocr(page, k_factor=kf, bbox_tolerance=tol)
            assert st == OpStatus.SUCCESS
            assert [l.text for l in r[2]] == [l.text for l in b[1][2]]
            assert [(l.bbox.x, l.bbox.y, l.bbox.w, l.bbox.h) for l in r[1]] == \
                   [(l.bbox.x, l.bbox.y, l.bbox.w, l.bbox.h) for l in b[1][1]]
    finally:
        pipe.close()


@pytest.mark.gpu
def test_gpu_mask_equals_oracle_mask_away_from_the_threshold(engine):
    """Same pages, stage a-1..a-3 only: the GPU mask and the fp32 oracle's agree on every pixel whose oracle logit
    is outside the bf16 tolerance of the threshold; the white-pixel count is within 1 % of the fixture's."""
    from ocr_inference_b200.inference import LineDetection
    from oracle import hostops as ho
    from oracle import nets

    fx = _fixture()["pages"]
    det = LineDetection(LineDetectionConfig("", seg_weights_fitted(11), "deeplabv3plus-r34", 512, ["line"]), engine=engine)
    pages = [synth.text_page(rec["seed"])[0] for rec in fx[:8]]
    masks = det.predict_batch(pages)
    for rec, m in zip(fx[:8], masks):
        assert abs(int((m == 255).sum()) - rec["mask_white"]) <= 0.01 * rec["mask_white"], rec["seed"]
    # live oracle on one page: pixel-level statement
    page = pages[0]
    ws = seg_weights_fitted(11)
    tiles, meta = ho.seg_preprocess(page, closed_form=False)
    x = np.repeat(tiles[:, None].astype(np.float32), 3, axis=1)
    logits = np.concatenate([nets.seg_forward(ws, x[i:i + 8]) for i in range(0, len(x), 8)])
    omask = ho.line_postprocess(logits, meta, 0.9, closed_form=False)
    assert int((omask == 255).sum()) == fx[0]["mask_white"], "the fixture is not what the oracle produces here"
    thr = float(ho.sigmoid_logit_threshold(0.9))
    tol = 0.08 * logits.std() + 0.05 * np.abs(logits)
    decided = (np.abs(logits - thr) > tol)[:, 0]
    dec = np.vstack([np.hstack(list(r)) for r in np.split(decided, meta["y_steps"])])[:page.shape[0], :page.shape[1]]
    assert np.array_equal(masks[0][dec], omask[dec])
    differ = int((masks[0] != omask).sum())
    print(f"mask pixels differing from the fp32 oracle: {differ} of {omask.size} ({100.0 * differ / omask.size:.4f} %), "
          f"undecided band {100.0 * (1 - dec.mean()):.3f} %")
    assert differ <= 0.002 * omask.size


def test_page_oracle_fixture_is_what_the_oracle_produces():
    """CPU: two pages of the committed fixture recomputed live by the oracle (needs no GPU and no /root/reference)."""
    import sys

    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_page_oracle as mpo

    fx = _fixture()
    ws, we = seg_weights_fitted(11), easter2_weights_fitted()
    for rec in (fx["pages"][0], fx["pages"][-1]):
        live = mpo.oracle_page(rec["seed"], ws, we, {"default": tuple(fx["settings"]["default"])})
        assert live["mask_white"] == rec["mask_white"]
        assert live["default"]["boxes"] == rec["default"]["boxes"]
        assert live["default"]["texts"] == rec["default"]["texts"]
        assert live["truth"] == rec["truth"]
