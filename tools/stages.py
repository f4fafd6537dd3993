"""The HBM-class stage rooflines of bench.py alone (B200): python tools/stages.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench  # noqa: E402
from ocr_inference_b200.engine import Engine  # noqa: E402

if __name__ == "__main__":
    st = bench.roofline_stages(Engine(0))
    for k, v in st.items():
        if isinstance(v, dict):
            print(f"{k:9s} kernel {v['us']:9.1f} us  call {v['us_call']:9.1f} us  {v['achieved']:8.1f} GB/s  frac {v['frac']:.3f}  [{v['timed'][:6]}]")
    print(json.dumps(st)[:200])
