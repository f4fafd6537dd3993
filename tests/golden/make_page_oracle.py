"""TEST-FIXTURE GENERATOR -- the CPU oracle's page -> text results for the whole-page parity test.

    python tests/golden/make_page_oracle.py        (CPU, ~10 s per page; writes tests/golden/page_oracle.json)

For each of N synthetic text pages (synthetic.text_page(seed), 2048x3072) it runs the ORACLE end to end -- cv2
tile/binarise, torch-CPU fp32 DeepLabV3+ (tests/golden/fitted_heads.npz decoder), mask threshold, the reference's
cv2/numpy line geometry + extraction restated in oracle/geometry.py, line prep, torch-CPU fp32 Easter2, greedy CTC
-- with the reference's default run_ocr parameters (k_factor 2.5, bbox_tolerance 4.0) and with the Demo-OCR
notebook's (2.0, 2.5), and stores per page: line boxes in reading order, decoded strings, the white-pixel count of
the mask.  tests/test_gpu_pages.py compares the GPU pipeline (its OWN mask, no injection) with this file at
BASELINE size and re-runs the oracle live on a few pages to show the file is what the oracle produces.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from ocr_inference_b200 import synthetic  # noqa: E402
from ocr_inference_b200.synthetic_weights import easter2_weights_fitted, seg_weights_fitted  # noqa: E402
from oracle import geometry as og  # noqa: E402
from oracle import hostops as ho  # noqa: E402
from oracle import nets  # noqa: E402

OUT = os.path.join(HERE, "page_oracle.json")
SEEDS = list(range(5000, 5032))
SETTINGS = {"default": (2.5, 4.0), "demo": (2.0, 2.5)}


def oracle_page(seed: int, ws, we, settings=SETTINGS):
    page, _, texts = synthetic.text_page(seed)
    tiles, meta = ho.seg_preprocess(page, closed_form=False)
    x = np.repeat(tiles[:, None].astype(np.float32), 3, axis=1)
    logits = np.concatenate([nets.seg_forward(ws, x[i:i + 8]) for i in range(0, len(x), 8)])
    mask = ho.line_postprocess(logits, meta, 0.9, closed_form=False)
    rec = {"seed": seed, "mask_white": int((mask == 255).sum()),
           "truth": ["".join(synthetic.CHARSET_76[i] for i in t).strip().replace("§", " ") for t in texts]}
    for name, (kf, tol) in settings.items():
        r = og.page_to_lines(page, mask, kf, tol)
        if isinstance(r, str):
            rec[name] = {"failed": r}
            continue
        txt = og.recognise(r[2], we, synthetic.CHARSET_76)[0]
        rec[name] = {"boxes": [[int(v) for v in ln.bbox] for ln in r[1]],
                     "texts": [t.strip().replace("§", " ") for t in txt],
                     "crop_shapes": [[int(v) for v in im.shape] for im in r[2]]}
    return rec


def main():
    import torch

    torch.set_num_threads(len(os.sched_getaffinity(0)))
    ws, we = seg_weights_fitted(11), easter2_weights_fitted()
    pages, t0 = [], time.time()
    for s in SEEDS:
        pages.append(oracle_page(s, ws, we))
        d = pages[-1]["default"]
        ok = sum(a == b for a, b in zip(d.get("texts", []), pages[-1]["truth"]))
        print(f"seed {s}: {len(d.get('texts', []))} lines, {ok} equal to the truth, {time.time() - t0:.0f}s", flush=True)
    with open(OUT, "w") as f:
        json.dump({"settings": SETTINGS, "pages": pages}, f, indent=0)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
