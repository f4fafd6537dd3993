"""One synthetic text page through the GPU pipeline and through the CPU oracle, stage by stage:
python tools/page_parity.py [seed]   (B200)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from ocr_inference_b200 import synthetic  # noqa: E402
from ocr_inference_b200.data import OpStatus  # noqa: E402
from ocr_inference_b200.inference import get_engine  # noqa: E402
from ocr_inference_b200.synthetic_weights import easter2_weights_fitted, seg_weights_fitted  # noqa: E402
from oracle import geometry as og  # noqa: E402
from oracle import hostops as ho  # noqa: E402
from oracle import nets  # noqa: E402
from tools.pipeline_profile import build_pipeline  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3003
page = synthetic.text_page(seed)[0]
pipe = build_pipeline(get_engine(0))
kw = dict(k_factor=2.0, bbox_tolerance=2.5)
st, res = pipe.run_ocr(page, **kw)
assert st == OpStatus.SUCCESS, res
gmask = pipe.line_inference.predict(page)
ws, we = seg_weights_fitted(11), easter2_weights_fitted()
tiles, meta = ho.seg_preprocess(page, closed_form=False)
logits = np.concatenate([nets.seg_forward(ws, np.repeat(tiles[i:i + 8, None].astype(np.float32), 3, axis=1))
                         for i in range(0, len(tiles), 8)])
omask = ho.line_postprocess(logits, meta, 0.9, closed_form=False)
print("mask pixels differing:", int((gmask != omask).sum()), "of", gmask.size)
r = og.page_to_lines(page, omask, 2.0, 2.5)
texts = og.recognise(r[2], we, synthetic.CHARSET_76)[0]
gb = [(l.bbox.x, l.bbox.y, l.bbox.w, l.bbox.h) for l in res[1]]
ob = [tuple(l.bbox) for l in r[1]]
print("GPU boxes   ", gb)
print("oracle boxes", ob)
# the GPU pipeline fed with the ORACLE's mask: stages 2-5 in isolation
pipe.line_inference.predict = lambda image, class_threshold=None: omask
st2, res2 = pipe.run_ocr(page, **kw)
for i, t in enumerate(texts):
    g = res[2][i].text if i < len(res[2]) else None
    g2 = res2[2][i].text if i < len(res2[2]) else None
    t = t.strip().replace("§", " ")
    print(f"line {i}: own mask {'==' if g == t else '!='} oracle ({bench._edit_distance(g or '', t)} edits of {len(t)}); "
          f"oracle mask {'==' if g2 == t else '!='} oracle ({bench._edit_distance(g2 or '', t)} edits)")
