"""cProfile of OCRPipeline.run_ocr on synthetic text pages (B200): python tools/pipeline_cprofile.py [pages]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ocr_inference_b200 import synthetic  # noqa: E402
from ocr_inference_b200.inference import get_engine  # noqa: E402
from tools.pipeline_profile import build_pipeline  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
pipe = build_pipeline(get_engine(0))
pages = [synthetic.text_page(3003 + i)[0] for i in range(n)]
pipe.run_ocr(pages[0])
pr = cProfile.Profile()
pr.enable()
for p in pages:
    pipe.run_ocr(p)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
