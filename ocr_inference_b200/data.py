"""Data types of the drop-in API.  Field names and order are the constructor contract of the
reference (BDRC/data.py:17-21, 24-35, 103-139, 170-217, 285-289); only the types the hot path
touches are provided."""
from __future__ import annotations

import enum
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Optional
from uuid import UUID

import numpy as np


class OpStatus(enum.Enum):          # BDRC/data.py:17-21
    SUCCESS = 0
    FAILED = 1


class Encoding(enum.Enum):          # BDRC/data.py:24-28
    UNICODE = 0
    WYLIE = 1


class CharsetEncoder(enum.Enum):    # BDRC/data.py:31-35
    WYLIE = 0
    STACK = 1


@dataclass
class BBox:                         # BDRC/data.py:103-110
    x: int
    y: int
    w: int
    h: int


@dataclass
class Line:                         # BDRC/data.py:120-127
    guid: UUID
    contour: np.ndarray
    bbox: BBox
    center: tuple[int, int]


@dataclass
class OCRLine:                      # BDRC/data.py:130-139
    guid: UUID
    text: str
    encoding: str
    ctc_conf: Optional[float]
    logits: Optional[list[float]]
    lm_scores: Optional[list[float]]


@dataclass
class DewarpingResult:              # BDRC/data.py:170-181
    work_img: np.ndarray
    work_mask: np.ndarray
    filtered_contours: list
    page_angle: float
    applied: bool
    tps_ratio: Optional[float] = None
    dewarped_img: Optional[np.ndarray] = None
    dewarped_mask: Optional[np.ndarray] = None


@dataclass
class LineDetectionConfig:          # BDRC/data.py:183-190
    checkpoint: str
    onnx_file: Any                  # path of a BDRCW1 container, its bytes, or a dict name -> ndarray
    architecture: str
    patch_size: int
    classes: list[str]


@dataclass
class LayoutDetectionConfig:        # BDRC/data.py:192-199
    checkpoint: str
    onnx_file: Any
    architecture: str
    patch_size: int
    classes: list[str]


@dataclass
class OCRModelConfig:               # BDRC/data.py:202-217
    model_file: Any                 # path of a BDRCW1 container, its bytes, or a dict name -> ndarray
    architecture: str
    input_width: int
    input_height: int
    input_layer: str
    output_layer: str
    squeeze_channel: bool
    swap_hw: bool
    encoder: CharsetEncoder
    charset: list[str]
    add_blank: bool
    version: str


@dataclass
class KenLMConfig:                  # BDRC/data.py:285-289
    kenlm_file: str | Path
    arpa_file: str | Path
    unigrams: list[str]
