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
