"""I/O edges of the page -> text path: image decode -> device, parquet / JSON result writers.

  ImageInferenceDataset   BDRC/inference.py:829-857 -- same constructor, len, get_item / __getitem__ -> (CHW u8
                          tensor, {"image_name", "orig_shape"}); decoded with cv2 (libjpeg-turbo / libpng, the
                          decoders torchvision's read_image wraps), so the pixels are the reference's.
  PageLoader              addition: the dataset read ahead by a few decode threads (cv2.imdecode releases the GIL)
                          into a bounded queue of HWC pages, ready for OCRPipeline.run_ocr_batch /
                          LineDetection.predict_batch, which move them to the GPU through their pinned slot rings.
  write_result_parquet    BDRC/utils.py:1111-1139 (line-detection result, Config.LINE_DETECTION_SCHEMA)
  save_ocr_lines_parquet  BDRC/utils.py:1240-1292 (OCRLine list)
  line_detection_result   addition: run_ocr's payload -> the dict write_result_parquet takes
"""
from __future__ import annotations

import os
import queue
import threading
from glob import glob
from pathlib import Path
from typing import Iterator, Sequence

import cv2
import numpy as np
import torch

IMAGE_SUFFIXES = (".jpg", ".png", ".jpeg")


def _decode(path: str, mode: str) -> np.ndarray:
    """HWC u8: RGB for mode 'rgb' (ImageReadMode.RGB), else the file's own channels (gray stays 1 channel)."""
    data = np.fromfile(path, dtype=np.uint8)
    if mode == "rgb":
        img = cv2.imdecode(data, cv2.IMREAD_COLOR)
        if img is None:
            raise OSError(f"cannot decode {path}")
        return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    img = cv2.imdecode(data, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise OSError(f"cannot decode {path}")
    if img.ndim == 2:
        return img[:, :, None]
    if img.shape[2] == 4:
        return cv2.cvtColor(img, cv2.COLOR_BGRA2RGBA)
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


class ImageInferenceDataset(torch.utils.data.Dataset):
    def __init__(self, root_dir: str, mode: str = "rgb"):
        self._mode = mode
        self.paths = sorted(p for p in glob(os.path.join(root_dir, "*")) if p.lower().endswith(IMAGE_SUFFIXES))

    def __len__(self):
        return len(self.paths)

    def get_item(self, idx):
        return self.__getitem__(idx)

    def __getitem__(self, idx):
        hwc = _decode(self.paths[idx], self._mode)
        img = torch.from_numpy(np.ascontiguousarray(hwc.transpose(2, 0, 1)))
        meta = {"image_name": os.path.basename(self.paths[idx]), "orig_shape": (img.shape[1], img.shape[2])}
        return img, meta


class PageLoader:
    """Iterates (HWC u8 RGB page, meta) over a directory (or a list of paths), decoding `workers` files ahead."""

    def __init__(self, source, workers: int = 4, prefetch: int = 8, mode: str = "rgb"):
        if isinstance(source, (str, os.PathLike)):
            self.paths = sorted(p for p in glob(os.path.join(str(source), "*")) if p.lower().endswith(IMAGE_SUFFIXES))
        else:
            self.paths = [str(p) for p in source]
        self.workers, self.prefetch, self.mode = max(1, workers), max(1, prefetch), mode

    def __len__(self):
        return len(self.paths)

    def __iter__(self) -> Iterator[tuple[np.ndarray, dict]]:
        n = len(self.paths)
        slots: dict[int, object] = {}
        cv = threading.Condition()
        nxt = [0]          # next index to hand to a decode thread
        head = [0]         # next index the consumer wants
        stop = [False]

        def work():
            while True:
                with cv:
                    while not stop[0] and (nxt[0] >= n or nxt[0] - head[0] >= self.prefetch):
                        if nxt[0] >= n:
                            return
                        cv.wait(0.1)
                    if stop[0]:
                        return
                    i = nxt[0]
                    nxt[0] += 1
                try:
                    item: object = _decode(self.paths[i], self.mode)
                except Exception as e:  # noqa: BLE001
                    item = e
                with cv:
                    slots[i] = item
                    cv.notify_all()

        threads = [threading.Thread(target=work, daemon=True, name="bdrc-decode") for _ in range(min(self.workers, max(1, n)))]
        for t in threads:
            t.start()
        try:
            for i in range(n):
                with cv:
                    while i not in slots:
                        cv.wait(0.1)
                    item = slots.pop(i)
                    head[0] = i + 1
                    cv.notify_all()
                if isinstance(item, Exception):
                    raise item
                yield item, {"image_name": os.path.basename(self.paths[i]), "orig_shape": item.shape[:2]}
        finally:
            with cv:
                stop[0] = True
                cv.notify_all()

    def batches(self, size: int):
        pages, metas = [], []
        for p, m in self:
            pages.append(p)
            metas.append(m)
            if len(pages) == size:
                yield pages, metas
                pages, metas = [], []
        if pages:
            yield pages, metas


# ---- parquet ---------------------------------------------------------------------------------------------------
def _line_detection_schema():
    import pyarrow as pa

    pt = pa.struct([("x", pa.int32()), ("y", pa.int32())])
    bb = pa.struct([("x", pa.int32()), ("y", pa.int32()), ("w", pa.int32()), ("h", pa.int32())])
    return pa.schema([("image_name", pa.string()), ("image_width", pa.int32()), ("image_height", pa.int32()),
                      ("num_contours", pa.int32()), ("contours", pa.list_(pa.list_(pt))), ("bboxes", pa.list_(bb))])


def bboxes_to_pyarrow(bboxes):
    return [{"x": x, "y": y, "w": w, "h": h} for (x, y, w, h) in bboxes]


def contours_to_arrow(contours):
    return [[{"x": x, "y": y} for x, y in contour] for contour in contours]


def line_detection_result(image_name: str, image_shape: Sequence[int], lines) -> dict:
    """run_ocr's sorted Line list -> the record write_result_parquet stores (contours as (x, y) point lists)."""
    return {"image_name": image_name, "image_width": int(image_shape[1]), "image_height": int(image_shape[0]),
            "num_contours": len(lines),
            "contours": [[(int(p[0][0]), int(p[0][1])) for p in ln.contour] for ln in lines],
            "bboxes": [(int(ln.bbox.x), int(ln.bbox.y), int(ln.bbox.w), int(ln.bbox.h)) for ln in lines]}


def write_result_parquet(result: dict, out_dir: str | Path):
    import pyarrow as pa
    import pyarrow.parquet as pq

    os.makedirs(out_dir, exist_ok=True)
    base_name, _ = os.path.splitext(result["image_name"])
    row = {k: result[k] for k in ("image_name", "image_width", "image_height", "num_contours")}
    row["contours"] = contours_to_arrow(result["contours"])
    row["bboxes"] = bboxes_to_pyarrow(result["bboxes"])
    table = pa.Table.from_pylist([row], schema=_line_detection_schema())
    pq.write_table(table, os.path.join(out_dir, f"{base_name}.parquet"), compression="zstd")


def save_ocr_lines_parquet(ocr_lines, out_path):
    import pyarrow as pa
    import pyarrow.parquet as pq

    cols = {"guid": [str(l.guid) for l in ocr_lines], "text": [l.text for l in ocr_lines],
            "encoding": [l.encoding for l in ocr_lines], "ctc_conf": [float(l.ctc_conf) for l in ocr_lines],
            "norm_logp": [float(getattr(l, "norm_logp", 0.0)) for l in ocr_lines],
            "n_beams": [len(l.logits) for l in ocr_lines], "logits": [l.logits for l in ocr_lines],
            "lm_scores": [l.lm_scores for l in ocr_lines]}
    schema = pa.schema([("guid", pa.string()), ("text", pa.string()), ("encoding", pa.string()), ("ctc_conf", pa.float32()),
                        ("norm_logp", pa.float32()), ("n_beams", pa.int16()), ("logits", pa.list_(pa.float32())),
                        ("lm_scores", pa.list_(pa.float32()))])
    pq.write_table(pa.Table.from_pydict(cols, schema=schema), out_path, compression="zstd", compression_level=7)
