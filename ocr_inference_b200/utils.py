"""Host glue on the reference side of the boundary: the model-config JSON readers of BDRC/utils.py, so that
a model directory laid out like the reference's (model_config.json next to the weights) can be pointed at the
B200 path.  Same keys, same return types; the only difference is that the "onnx-model" entry names a BDRCW1
weight container (see INTEGRATION.md for the one-off ONNX -> BDRCW1 conversion).  If a file with the same stem
and the extension `.bdrcw` exists next to the named `.onnx`, it is used.
"""
from __future__ import annotations

import json
import os

from .data import CharsetEncoder, LayoutDetectionConfig, LineDetectionConfig, OCRModelConfig

CHARSETENCODER = {"wylie": CharsetEncoder.WYLIE, "stack": CharsetEncoder.STACK}      # Config.py:69


def _weights_path(model_dir: str, name: str) -> str:
    path = f"{model_dir}/{name}"
    alt = os.path.splitext(path)[0] + ".bdrcw"
    return alt if os.path.exists(alt) else path


def get_charset(charset) -> list[str]:
    """BDRC/utils.py:210-217."""
    return [x for x in charset]


def read_line_model_config(config_file: str) -> LineDetectionConfig:
    """BDRC/utils.py:162-180."""
    model_dir = os.path.dirname(config_file)
    with open(config_file, encoding="utf-8") as f:
        js = json.loads(f.read())
    return LineDetectionConfig(f"{model_dir}/{js['checkpoint']}", _weights_path(model_dir, js["onnx-model"]),
                               js["architecture"], int(js["patch_size"]), js["classes"])


def read_layout_model_config(config_file: str) -> LayoutDetectionConfig:
    """BDRC/utils.py:183-207 (the reference prefixes the architecture with the model directory, :196)."""
    model_dir = os.path.dirname(config_file)
    with open(config_file, encoding="utf-8") as f:
        js = json.loads(f.read())
    return LayoutDetectionConfig(f"{model_dir}/{js['checkpoint']}", _weights_path(model_dir, js["onnx-model"]),
                                 f"{model_dir}/{js['architecture']}", int(js["patch_size"]), js["classes"])


def read_ocr_model_config(config_file: str) -> OCRModelConfig:
    """BDRC/utils.py:349-393: "yes"/"no" strings become booleans, the encoder name an enum."""
    model_dir = os.path.dirname(config_file)
    with open(config_file, encoding="utf-8") as f:
        js = json.loads(f.read())
    return OCRModelConfig(
        _weights_path(model_dir, js["onnx-model"]), js["architecture"], js["input_width"], js["input_height"],
        js["input_layer"], js["output_layer"], js["squeeze_channel_dim"] == "yes", js["swap_hw"] == "yes",
        encoder=CHARSETENCODER[js["encoder"]], charset=js["charset"], add_blank=js["add_blank"] == "yes",
        version=js["version"])
