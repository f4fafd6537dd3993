"""Result exporters: the file formats of BDRC/exporter.py:24-407 -- plain text, PageXML (2013-07-15
schema, Transkribus-style), JSON -- written from the Line / OCRLine lists run_ocr returns, plus `export_batch`,
which feeds a whole run_ocr_batch result list to any exporter.  Same class names, method names and arguments."""
from __future__ import annotations

import json
import logging
import xml.etree.ElementTree as etree  # nosec B405
from datetime import datetime
from typing import List, Sequence
from xml.dom import minidom

import numpy as np

from .data import BBox, Line, OCRLine, OpStatus
from .line_detection import optimize_countour

PAGE_NS = "http://schema.primaresearch.org/PAGE/gts/pagecontent/2013-07-15"


def get_utc_time() -> str:
    """BDRC/utils.py:92-102 ((sic) local time, ISO format to the second)."""
    return datetime.now().strftime("%Y-%m-%dT%H:%M:%S")


def get_text_bbox(lines: list[Line]) -> BBox:
    """BDRC/utils.py:611-621 ((sic): widest line's width and the LAST line's bottom edge, not a union)."""
    boxes = [ln.bbox for ln in lines]
    return BBox(min(b.x for b in boxes), min(b.y for b in boxes), max(b.w for b in boxes), boxes[-1].y + boxes[-1].h)


def rotate_contour(cnt, center: tuple[int, int], angle: float):
    """BDRC/utils.py:636-658: rotation about `center` in polar form, truncated to int32."""
    rel = cnt - [center[0], center[1]]
    xs, ys = rel[:, 0, 0], rel[:, 0, 1]
    theta = np.deg2rad((np.rad2deg(np.arctan2(ys, xs)) + angle) % 360)
    rho = np.hypot(xs, ys)
    rel[:, 0, 0] = rho * np.cos(theta)
    rel[:, 0, 1] = rho * np.sin(theta)
    return (rel + [center[0], center[1]]).astype(np.int32)


class Exporter:
    def __init__(self, output_dir: str):
        self.output_dir = output_dir
        try:
            import pyewts  # type: ignore

            self.converter = pyewts.pyewts()
        except Exception:  # noqa: BLE001  (third-party transliteration, absent here: SURVEY 8c)
            self.converter = None
        logging.info("Init Exporter")

    def export_text(self, image_name: str, text_lines: List[OCRLine]):
        raise NotImplementedError

    def export_lines(self, image, image_name: str, lines: List[Line], text_lines):
        raise NotImplementedError

    @staticmethod
    def get_bbox(bbox: BBox) -> tuple[int, int, int, int]:
        return bbox.x, bbox.y, bbox.w, bbox.h

    @staticmethod
    def get_text_points(contour) -> str:
        return "".join(f"{p[0][0]},{p[0][1]} " for p in contour)

    @staticmethod
    def get_bbox_points(bbox: BBox) -> str:
        x0, y0, x1, y1 = bbox.x, bbox.y, bbox.x + bbox.w, bbox.y + bbox.h
        return f"{x0},{y0} {x1},{y0} {x1},{y1} {x0},{y1}"

    # shared by the XML and JSON writers: optional rotation back, polygon simplification, coordinate strings
    def _plain(self, image, lines: List[Line], optimize: bool, bbox: bool, angle: float):
        if angle != abs(0):
            centre = (image.shape[1] // 2, image.shape[0] // 2)
            for ln in lines:
                ln.contour = rotate_contour(ln.contour, centre, angle)
        if optimize:
            for ln in lines:
                ln.contour = optimize_countour(ln.contour)
        plain = [self.get_bbox(ln.bbox) for ln in lines] if bbox else [self.get_text_points(ln.contour) for ln in lines]
        return plain, self.get_bbox_points(get_text_bbox(lines))


class PageXMLExporter(Exporter):
    def __init__(self, output_dir: str) -> None:
        super().__init__(output_dir)
        logging.info("Init XML Exporter")

    def get_text_line_block(self, coordinate, index: int, unicode_text: str):
        el = etree.Element("TextLine")
        el.attrib["id"] = f"line_9874_{index}"
        el.attrib["custom"] = f"readingOrder {{index: {index};}}"
        etree.SubElement(el, "Coords").attrib["points"] = coordinate
        etree.SubElement(etree.SubElement(el, "TextEquiv"), "Unicode").text = unicode_text
        return el

    def build_xml_document(self, image, image_name: str, text_bbox: str, lines, text_lines):
        root = etree.Element("PcGts")
        root.attrib["xmlns"] = PAGE_NS
        root.attrib["xmlns:xsi"] = "http://www.w3.org/2001/XMLSchema-instance"
        root.attrib["xsi:schemaLocation"] = f"{PAGE_NS} {PAGE_NS}/pagecontent.xsd"
        meta = etree.SubElement(root, "Metadata")
        etree.SubElement(meta, "Creator").text = "Transkribus"
        etree.SubElement(meta, "Created").text = get_utc_time()
        page = etree.SubElement(root, "Page")
        page.attrib["imageFilename"] = image_name
        page.attrib["imageWidth"] = f"{image.shape[1]}"
        page.attrib["imageHeight"] = f"{image.shape[0]}"
        order = etree.SubElement(page, "ReadingOrder")
        group = etree.SubElement(order, "OrderedGroup")
        group.attrib["id"] = "1234_0"
        group.attrib["caption"] = "Regions reading order"
        ref = etree.SubElement(order, "RegionRefIndexed")        # (sic) a sibling of the group, as the reference writes it
        ref.attrib["index"] = "0"
        ref.attrib["regionRef"] = "region_main"
        region = etree.SubElement(page, "TextRegion")
        region.attrib["id"] = "region_main"
        region.attrib["custom"] = "readingOrder {index:0;}"
        etree.SubElement(region, "Coords").attrib["points"] = text_bbox
        have_text = text_lines is not None and len(text_lines) > 0
        for i, coord in enumerate(lines):
            region.append(self.get_text_line_block(coord, i, text_lines[i].text if have_text else ""))
        return minidom.parseString(etree.tostring(root)).toprettyxml()

    def export_lines(self, image, image_name: str, lines: List[Line], text_lines: List[OCRLine], optimize: bool = True,
                     bbox: bool = False, angle: float = 0.0):
        plain, box = self._plain(image, lines, optimize, bbox, angle)
        doc = self.build_xml_document(image, image_name, text_bbox=box, lines=plain, text_lines=text_lines)
        with open(f"{self.output_dir}/{image_name}.xml", "w", encoding="UTF-8") as f:
            f.write(doc)


class TextExporter(Exporter):
    def __init__(self, output_dir: str) -> None:
        super().__init__(output_dir)
        logging.info("Init Text Exporter")

    def export_lines(self, image, image_name: str, lines: List[Line], text_lines: list[OCRLine], **_):
        self.export_text(image_name, text_lines)

    def export_text(self, image_name: str, text_lines: List[OCRLine]):
        with open(f"{self.output_dir}/{image_name}.txt", "w", encoding="UTF-8") as f:
            f.writelines(f"{ln.text}\n" for ln in text_lines)


class JsonExporter(Exporter):
    def __init__(self, output_dir: str) -> None:
        super().__init__(output_dir)
        logging.info("Init JSON Exporter")

    def export_lines(self, image, image_name: str, lines: List[Line], text_lines: list[OCRLine], optimize: bool = True,
                     bbox: bool = False, angle: float = 0.0):
        plain, box = self._plain(image, lines, optimize, bbox, angle)
        record = {"image": image_name, "textbox": box, "lines": plain, "text": [t.text for t in text_lines]}
        with open(f"{self.output_dir}/{image_name}.jsonl", "w", encoding="UTF-8") as f:
            json.dump(record, f, ensure_ascii=False, indent=1)


def export_batch(exporters: Sequence[Exporter], images, image_names: Sequence[str], results, **kw) -> int:
    """Feeds OCRPipeline.run_ocr_batch's result list to the exporters (addition).  Pages whose status is FAILED are
    skipped; returns the number of exported pages.  `kw` goes to export_lines (optimize / bbox / angle)."""
    import copy

    done = 0
    for image, name, (status, payload) in zip(images, image_names, results):
        if status != OpStatus.SUCCESS:
            continue
        _, lines, ocr_lines, angle = payload
        for ex in exporters:
            ex.export_lines(image, name, copy.deepcopy(lines), ocr_lines, **({"angle": 0.0} if isinstance(ex, TextExporter) else kw))
        done += 1
    return done
