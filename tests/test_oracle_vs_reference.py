"""Live pin of oracle/ against the UNMODIFIED reference (imported through oracle/ref_shim.py).
Skipped where /root/reference is absent (GPU box); the golden vectors cover that case."""
import numpy as np
import pytest

from ocr_inference_b200 import synthetic
from ocr_inference_b200.synthetic_weights import easter2_weights, seg_weights
from oracle import geometry as og
from oracle import hostops as ho
from oracle import nets, ref_shim

pytestmark = pytest.mark.needs_reference


@pytest.fixture(scope="module")
def ref():
    return ref_shim.import_reference()


def test_full_run_ocr_matches_oracle(ref):
    """OCRPipeline.run_ocr of the reference (fake ORT sessions = torch-CPU oracle nets, greedy
    decoder, identity pyewts) vs the oracle's own page -> text restatement."""
    V = 77
    we = easter2_weights(7, V, 80)
    page, boxes = synthetic.page(31, 1024, 1536)
    band = synthetic.band_mask(boxes, 1024, 1536)
    # segmentation session: returns logits whose sigmoid>0.9 decision equals the band mask
    def seg_session(x):
        assert x.shape[1:] == (3, 512, 512) and x.dtype == np.float32
        n = x.shape[0]
        tiles = band.reshape(2, 512, 3, 512).transpose(0, 2, 1, 3).reshape(n, 1, 512, 512)
        return np.where(tiles > 0, 8.0, -8.0).astype(np.float32)
    calls = []
    def ocr_session(x):
        calls.append(x.shape)
        return nets.easter2_forward(we, x)
    ref_shim.SESSION_REGISTRY["seg.onnx"] = seg_session
    ref_shim.SESSION_REGISTRY["ocr.onnx"] = ocr_session
    D = ref.data
    pipe = ref.inference.OCRPipeline(
        D.OCRModelConfig("ocr.onnx", "Easter2", 2000, 80, "input", "output", True, False, D.CharsetEncoder.STACK,
                         synthetic.CHARSET_76, True, "t"),
        D.LineDetectionConfig("", "seg.onnx", "dlv3p", 512, ["line"]), None)
    status, result = pipe.run_ocr(page, k_factor=2.0, bbox_tolerance=2.5)
    assert status == D.OpStatus.SUCCESS, result
    rot_mask, lines, ocr_lines, angle = result
    assert all(c == (1, 80, 2000) for c in calls) and len(calls) == len(lines)
    o_mask, o_lines, o_imgs, o_angle = og.page_to_lines(page, band, 2.0, 2.5)
    assert np.array_equal(rot_mask, o_mask) and angle == o_angle
    assert [ln.center for ln in lines] == [ln.center for ln in o_lines]
    texts, _ = og.recognise(o_imgs, we, synthetic.CHARSET_76)
    assert [l.text for l in ocr_lines] == texts


def test_layout_predict_matches_oracle(ref):
    page, _ = synthetic.page(5, 700, 900)
    rng = np.random.default_rng(5)
    tiles, meta = ho.seg_preprocess(page, closed_form=False)
    logits = rng.normal(0, 2, (tiles.shape[0], 5, 512, 512)).astype(np.float32)
    ref_shim.SESSION_REGISTRY["lay.onnx"] = lambda x: logits
    D = ref.data
    lay = ref.inference.LayoutDetection(D.LayoutDetectionConfig("", "lay.onnx", "x", 512, ["background", "image", "line", "caption", "margin"]))
    got = lay.predict(page)
    assert np.array_equal(got, ho.layout_postprocess(logits, meta, 0.8, closed_form=True))


def test_torch_path_collate_matches_oracle(ref):
    import torch
    for (H, W) in [(700, 900), (400, 1500), (2300, 1000)]:
        page, _ = synthetic.page(H, H, W, rgb=True)
        chw = np.ascontiguousarray(page.transpose(2, 0, 1))
        all_tiles, ranges, metas = ref.utils.multi_image_collate_fn([(torch.from_numpy(chw), {"image_name": "a"})])
        x, meta = ho.torch_path_pre(chw)
        assert np.array_equal(all_tiles.numpy(), x)
        assert (metas[0]["pad_x"], metas[0]["pad_y"], metas[0]["x_steps"], metas[0]["y_steps"]) == \
               (meta["pad_x"], meta["pad_y"], meta["x_steps"], meta["y_steps"])


def test_check_line_tps_matches_reference_on_narrow_and_wide_lines(ref):
    """image_dewarping.check_line_tps works on the line's ROI; the reference slices a page-sized canvas.  Lines
    narrower than 4 slice widths put slices 2 and 5 partly outside the bounding box (negative ROI starts): the
    control points must still be the reference's (BDRC/image_dewarping.py:118-193)."""
    import cv2

    from ocr_inference_b200 import image_dewarping as idw

    ref_idw = __import__("BDRC.image_dewarping", fromlist=["check_line_tps"])
    rng = np.random.default_rng(5)
    checked, narrow = 0, 0
    for w in (30, 60, 100, 150, 159, 160, 400, 1200):
        for x0 in (3, 50, 700):
            H, W = 400, 2000
            canvas = np.zeros((H, W), np.uint8)
            # a steep slanted band (plus a ripple): the slices' centres spread by more than a slice height
            xs = np.arange(x0, x0 + w)
            yc = 120 + (170.0 * (xs - x0) / w + 6 * np.sin(xs / 9.0)).astype(int)
            cv2.polylines(canvas, [np.stack([xs, yc], axis=1).astype(np.int32)], False, 255, int(rng.integers(9, 15)))
            cnts, _ = cv2.findContours(canvas, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
            cnt = max(cnts, key=cv2.contourArea)
            image = np.zeros((H, W, 3), np.uint8)
            want = ref_idw.check_line_tps(image, cnt)
            got = idw.check_line_tps(image, cnt)
            assert got[0] == want[0], (w, x0)
            if want[0]:
                assert np.array_equal(np.asarray(got[1]), np.asarray(want[1])), (w, x0, got[1], want[1])
                assert np.array_equal(np.asarray(got[2]), np.asarray(want[2])), (w, x0)
                assert got[3] == want[3]
                checked += 1
                narrow += int(w < 160)
    assert checked >= 8 and narrow >= 2, (checked, narrow)


def test_exporters_write_the_reference_files(ref, tmp_path):
    """exporter.py (PageXML / text / JSON) against BDRC/exporter.py on the same lines: identical files (the XML
    apart from its time stamp), for contours and for bounding boxes, with and without rotation."""
    import copy
    import re
    from uuid import uuid1

    import cv2

    from ocr_inference_b200 import exporter as mine
    from ocr_inference_b200.data import BBox, Line, OCRLine

    ref_exp = __import__("BDRC.exporter", fromlist=["PageXMLExporter"])
    ref_data = __import__("BDRC.data", fromlist=["Line"])
    rng = np.random.default_rng(4)
    image = np.zeros((900, 1400, 3), np.uint8)
    lines, texts = [], []
    for i in range(5):
        canvas = np.zeros((900, 1400), np.uint8)
        pts = np.stack([np.arange(100, 1300, 40), 100 + 150 * i + rng.integers(-8, 9, 30)], axis=1).astype(np.int32)
        cv2.polylines(canvas, [pts], False, 255, 22)
        cnt = max(cv2.findContours(canvas, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)[0], key=cv2.contourArea)
        x, y, w, h = cv2.boundingRect(cnt)
        lines.append((uuid1(), cnt, (x, y, w, h), (x + w // 2, y + h // 2)))
        texts.append((lines[-1][0], f"line {i} ཀ་", "UNICODE"))
    for kw in (dict(), dict(bbox=True), dict(angle=1.5, optimize=False)):
        for cls in ("PageXMLExporter", "JsonExporter", "TextExporter"):
            outs = []
            for mod, dmod, sub in ((ref_exp, ref_data, "ref"), (mine, None, "mine")):
                d = tmp_path / f"{sub}_{cls}_{len(kw)}"
                d.mkdir(exist_ok=True)
                L = (dmod.Line if dmod else Line)
                B = (dmod.BBox if dmod else BBox)
                O = (dmod.OCRLine if dmod else OCRLine)
                ls = [L(g, c.copy(), B(*bb), ctr) for g, c, bb, ctr in lines]
                ts = [O(g, t, e, None, None, None) for g, t, e in texts]
                ex = getattr(mod, cls)(str(d))
                if cls == "TextExporter":
                    ex.export_lines(image, "page_01", ls, ts)
                    ex.export_text("page_02", ts)
                else:
                    ex.export_lines(image, "page_01", ls, ts, **kw)
                files = sorted(p for p in d.iterdir())
                outs.append([re.sub(r"<Created>[^<]*</Created>", "<Created/>", p.read_text(encoding="UTF-8")) for p in files])
                assert files
            assert outs[0] == outs[1], (cls, kw)


def test_parquet_writers_match_the_reference(ref, tmp_path):
    import pyarrow.parquet as pq

    from ocr_inference_b200 import dataio
    from ocr_inference_b200.data import OCRLine

    ref_utils = __import__("BDRC.utils", fromlist=["write_result_parquet"])
    result = {"image_name": "p1.jpg", "image_width": 300, "image_height": 200, "num_contours": 2,
              "contours": [[(1, 2), (30, 2), (30, 20)], [(5, 50), (90, 52), (88, 70), (4, 66)]],
              "bboxes": [(1, 2, 30, 19), (4, 50, 87, 21)]}
    ref_utils.write_result_parquet(result, tmp_path / "ref")
    dataio.write_result_parquet(result, tmp_path / "mine")
    a, b = pq.read_table(tmp_path / "ref" / "p1.parquet"), pq.read_table(tmp_path / "mine" / "p1.parquet")
    assert a.schema == b.schema and a.to_pylist() == b.to_pylist()
    lines = [OCRLine("g1", "abc", "UNICODE", 0.9, [-1.0, -2.5], None), OCRLine("g2", "", "UNICODE", 0.5, [-3.0], [0.1, 0.2])]
    ref_utils.save_ocr_lines_parquet(lines, tmp_path / "ref.parquet")
    dataio.save_ocr_lines_parquet(lines, tmp_path / "mine.parquet")
    a, b = pq.read_table(tmp_path / "ref.parquet"), pq.read_table(tmp_path / "mine.parquet")
    assert a.schema == b.schema and a.to_pylist() == b.to_pylist()
