"""Generates the committed golden vectors by running the UNMODIFIED reference
(/root/reference/BDRC/*.py, imported through oracle/ref_shim.py) on small seeded inputs.

    python tests/golden/make_golden.py        # build container only (needs /root/reference)

The .npz files are what tests/test_oracle_golden.py checks oracle/ against on any machine.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from ocr_inference_b200 import synthetic  # noqa: E402
from oracle import ref_shim  # noqa: E402


def main():
    ref = ref_shim.import_reference()
    inf, ld = ref.inference, ref.line_detection

    class Det(inf.LineDetection):
        def __init__(self):
            self._patch_size = 512
            self.session_out = None

        def _predict(self, batch):
            return self.session_out

    class Lay(inf.LayoutDetection):
        def __init__(self):
            self._patch_size = 512
            self.session_out = None

        def _predict(self, batch):
            return self.session_out

    out = {}
    # a-1: Detection._preprocess_image on pages hitting every clamp branch
    for name, (H, W) in {"plain": (600, 700), "upsize": (300, 900), "tall": (2300, 1000), "wide": (700, 4300)}.items():
        page, _ = synthetic.page(100 + H, H, W)
        det = Det()
        _, tiles, ys, px, py = det._preprocess_image(page, 512)
        assert tiles.dtype == np.float32 and (tiles[..., 0] == tiles[..., 1]).all()
        out[f"pre_{name}_tiles"] = np.packbits(tiles[..., 0].astype(np.uint8), axis=None)
        out[f"pre_{name}_meta"] = np.array([H, W, tiles.shape[0], ys, px, py])
        # a-3: LineDetection.predict post-processing on seeded logits
        rng = np.random.default_rng(H)
        low = rng.normal(1.5, 3.0, (tiles.shape[0], 1, 32, 32)).astype(np.float32)
        logits = np.stack([cv2.resize(l[0], (512, 512), interpolation=cv2.INTER_CUBIC) for l in low])[:, None]
        det.session_out = logits
        mask = det.predict(page)
        out[f"post_{name}_low"] = low
        out[f"post_{name}_mask"] = np.packbits(mask // 255, axis=None)
        lay = Lay()
        low5 = rng.normal(0, 3.0, (tiles.shape[0], 5, 16, 16)).astype(np.float32)
        logits5 = np.stack([[cv2.resize(c, (512, 512), interpolation=cv2.INTER_CUBIC) for c in l] for l in low5])
        lay.session_out = logits5
        m5 = lay.predict(page)
        out[f"lay_{name}_low"] = low5
        out[f"lay_{name}_mask"] = np.packbits(m5 // 255, axis=None)
    # a-7: OCRInference._prepare_ocr_line (+ _pre_pad)
    ocr = inf.OCRInference.__new__(inf.OCRInference)
    ocr._input_width, ocr._input_height = 2000, 80
    for i, (h, w) in enumerate([(60, 1500), (140, 2900), (80, 2000), (30, 400), (200, 900)]):
        crop = synthetic.line_crop(500 + i, h, w)
        a = ocr._prepare_ocr_line(crop)
        b = ocr._prepare_ocr_line(ocr._pre_pad(crop))
        assert set(np.unique(a)) <= {-1.0, 1.0}
        out[f"prep_{i}_hw"] = np.array([h, w])
        out[f"prep_{i}_plain"] = np.packbits((a[0] > 0).astype(np.uint8), axis=None)
        out[f"prep_{i}_prepad"] = np.packbits((b[0] > 0).astype(np.uint8), axis=None)
    # a-5 / a-10: sort + extract on a page with an injected band mask
    page, boxes = synthetic.page(700, 800, 1200)
    mask = synthetic.band_mask(boxes, 800, 1200)
    rot_img, rot_mask, contours, angle = ld.build_raw_line_data(page, mask)
    filt = ld.filter_line_contours(rot_mask, contours)
    lines = [ld.build_line_data(c) for c in filt]
    sorted_lines, thr = ld.sort_lines_by_threshold2(rot_mask, lines, group_lines=True)
    imgs = ld.extract_line_images(rot_img, sorted_lines, 2.0, 2.5)
    out["lines_angle_thr"] = np.array([float(angle), float(thr)])
    out["lines_centers"] = np.array([ln.center for ln in sorted_lines])
    out["lines_bboxes"] = np.array([[ln.bbox.x, ln.bbox.y, ln.bbox.w, ln.bbox.h] for ln in sorted_lines])
    out["lines_shapes"] = np.array([im.shape for im in imgs])
    out["lines_sums"] = np.array([int(im.astype(np.int64).sum()) for im in imgs])
    out["lines_first"] = imgs[0]
    # a-9: CTCDecoder index conventions
    dec = inf.CTCDecoder(synthetic.CHARSET_76, True, None)
    rng = np.random.default_rng(9)
    lg = rng.normal(0, 1, (120, 77)).astype(np.float32)
    lg[:, 0] += 1.0
    out["ctc_logits"] = lg
    out["ctc_text"] = np.array(dec.ctc_decode(lg))
    out["ctc_encode"] = np.array(dec.encode("abc§Z9"))
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), os.path.getsize(os.path.join(HERE, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
