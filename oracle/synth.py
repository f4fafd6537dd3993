"""TEST INFRASTRUCTURE ONLY -- re-export of the deterministic synthetic page / line generators
(they live in ocr_inference_b200/synthetic.py because bench.py's product arm needs them too)."""
from ocr_inference_b200.synthetic import *  # noqa: F401,F403
from ocr_inference_b200.synthetic import CHARSET_76, band_mask, line_crop, line_crops, page  # noqa: F401
