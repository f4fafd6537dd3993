"""Deterministic synthetic pages / line crops: inputs for bench.py and the tests.

All generators are numpy.random.default_rng(seed) driven, u8, 3 identical channels unless noted.
"""
from __future__ import annotations

import cv2
import numpy as np

CHARSET_76 = list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789'+-./_~()[]|§&")
assert len(CHARSET_76) == 76 and len(set(CHARSET_76)) == 76


def _strokes(rng, canvas, y0, y1, x0, x1, n, dark=(20, 90)):
    for _ in range(n):
        cx = int(rng.integers(x0, max(x0 + 1, x1)))
        cy = int(rng.integers(y0, max(y0 + 1, y1)))
        w = int(rng.integers(3, 18))
        h = int(rng.integers(3, max(4, (y1 - y0) // 2)))
        col = int(rng.integers(*dark))
        if rng.random() < 0.5:
            cv2.ellipse(canvas, (cx, cy), (w, h), float(rng.integers(0, 180)), 0, 360, col, int(rng.integers(1, 4)))
        else:
            cv2.line(canvas, (cx - w, cy - h // 2), (cx + w, cy + h // 2), col, int(rng.integers(1, 4)))


def page(seed: int, H: int = 2048, W: int = 3072, n_lines: tuple[int, int] = (14, 22), rgb: bool = False):
    """Synthetic page with horizontal text bands. Returns (u8 [H,W,3], list of band boxes (x,y,w,h))."""
    rng = np.random.default_rng(seed)
    g = np.clip(rng.normal(215, 12, (H, W)), 0, 255).astype(np.uint8)
    n = int(rng.integers(n_lines[0], n_lines[1] + 1))
    band_h = max(12, min(70, int(H / (n * 2.2))))
    ys = np.linspace(band_h, H - 2 * band_h, n).astype(int)
    boxes = []
    for y in ys:
        h = int(rng.integers(max(10, int(band_h * 0.6)), band_h + 1))
        x0 = int(rng.integers(W // 40, W // 10))
        x1 = W - int(rng.integers(W // 40, W // 10))
        _strokes(rng, g, y, y + h, x0, x1, n=max(20, (x1 - x0) // 12))
        boxes.append((x0, int(y), x1 - x0, h))
    g = cv2.GaussianBlur(g, (3, 3), 0)
    if rgb:
        img = np.stack([g, np.clip(g.astype(int) - 6, 0, 255).astype(np.uint8),
                        np.clip(g.astype(int) - 14, 0, 255).astype(np.uint8)], axis=-1)
    else:
        img = np.stack([g, g, g], axis=-1)
    return np.ascontiguousarray(img), boxes


def band_mask(boxes, H: int, W: int, grow: int = 4, thin: float = 1.0) -> np.ndarray:
    """u8 {0,255} mask with one blob per text band (a trained line model's output shape).  thin < 1 keeps
    only the central fraction of each band's height: line models mark the core of a text line, and the
    reference's extraction (dilate by ~2.4x the blob height) expects that."""
    m = np.zeros((H, W), np.uint8)
    for (x, y, w, h) in boxes:
        cut = int(round(h * (1.0 - thin) / 2))
        cv2.rectangle(m, (x, max(0, y + cut - grow)), (x + w, min(H - 1, y + h - cut + grow)), 255, -1)
    return m


def line_crop(seed: int, h: int | None = None, w: int | None = None) -> np.ndarray:
    """One ragged line crop u8 [h,w,3]: bright background, dark strokes, sigma=1 blur (C1 inputs)."""
    rng = np.random.default_rng(seed)
    h = int(rng.integers(60, 141)) if h is None else h
    w = int(rng.integers(900, 2901)) if w is None else w
    g = np.clip(rng.normal(225, 10, (h, w)), 0, 255).astype(np.uint8)
    _strokes(rng, g, h // 6, h - h // 6, 10, w - 10, n=int(rng.integers(20, 61)) * max(1, w // 900), dark=(10, 90))
    g = cv2.GaussianBlur(g, (0, 0), 1.0)
    return np.ascontiguousarray(np.stack([g, g, g], axis=-1))


def line_crops(seed: int, n: int, fixed: tuple[int, int] | None = None) -> list[np.ndarray]:
    return [line_crop(seed * 100003 + i, *(fixed if fixed else (None, None))) for i in range(n)]


# ------------------------------------------------------------------------------------------
# "glyph" lines: text-like line crops with a known transcription (76 fixed stroke glyphs, one per
# charset symbol).  Used to fit the recogniser's output layer so that its logits are peaky like a
# trained model's (tests/golden/make_heads.py) and as recognisable bench / parity inputs.
# ------------------------------------------------------------------------------------------
GLYPH_H, GLYPH_W = 44, 20
_GLYPHS = None


def glyph_bitmaps() -> np.ndarray:
    """[76, GLYPH_H, GLYPH_W] u8 {0,1}: 1 = ink.  Fixed seed: part of the synthetic 'script'."""
    global _GLYPHS
    if _GLYPHS is None:
        rng = np.random.default_rng(760076)
        g = np.zeros((76, GLYPH_H, GLYPH_W), np.uint8)
        for i in range(76):
            c = np.zeros((GLYPH_H, GLYPH_W), np.uint8)
            for _ in range(int(rng.integers(3, 6))):
                p0 = (int(rng.integers(1, GLYPH_W - 1)), int(rng.integers(1, GLYPH_H - 1)))
                p1 = (int(rng.integers(1, GLYPH_W - 1)), int(rng.integers(1, GLYPH_H - 1)))
                if rng.random() < 0.65:
                    cv2.line(c, p0, p1, 1, int(rng.integers(2, 4)))
                else:
                    cv2.ellipse(c, p0, (int(rng.integers(3, 8)), int(rng.integers(4, 12))), 0, 0, 360, 1, 2)
            g[i] = c
        _GLYPHS = g
    return _GLYPHS


def glyph_line(seed: int, n_chars: int | None = None, h: int | None = None):
    """One line crop u8 [h,w,3] of glyphs + its transcription (indices into CHARSET_76)."""
    rng = np.random.default_rng(seed)
    glyphs = glyph_bitmaps()
    h = int(rng.integers(64, 97)) if h is None else h
    n = int(rng.integers(24, 61)) if n_chars is None else n_chars
    ids = rng.integers(0, 76, n)
    gaps = rng.integers(5, 12, n + 1)
    w = int(gaps.sum() + n * GLYPH_W + 16)
    ink = np.zeros((h, w), np.uint8)
    x = 8 + int(gaps[0])
    y0 = (h - GLYPH_H) // 2
    for i, g in enumerate(ids):
        dy = int(rng.integers(-3, 4))
        ys = min(max(0, y0 + dy), h - GLYPH_H)
        ink[ys:ys + GLYPH_H, x:x + GLYPH_W] |= glyphs[g]
        x += GLYPH_W + int(gaps[i + 1])
    bg = np.clip(rng.normal(225, 8, (h, w)), 0, 255)
    fg = np.clip(rng.normal(40, 12, (h, w)), 0, 255)
    img = np.where(ink > 0, fg, bg).astype(np.uint8)
    img = cv2.GaussianBlur(img, (0, 0), 0.8)
    return np.ascontiguousarray(np.stack([img, img, img], axis=-1)), [int(i) for i in ids]


def glyph_lines(seed: int, n: int, **kw):
    out = [glyph_line(seed * 100003 + i, **kw) for i in range(n)]
    return [o[0] for o in out], [o[1] for o in out]


def text_page(seed: int, H: int = 2048, W: int = 3072, n_lines: tuple[int, int] = (6, 7), glyph_scale: float = 1.5):
    """Synthetic page whose text bands are glyph lines with a known transcription (config C3/C5).
    Glyphs are drawn `glyph_scale` x larger than in glyph_line(): the reference's line extraction returns
    crops about twice as tall as the text (dilated contour, BDRC/line_detection.py:267-288), so after the
    resize to 80 px the glyphs arrive at the size the recogniser was trained on.
    Returns (u8 [H,W,3], band boxes (x,y,w,h), transcriptions as CHARSET_76 index lists)."""
    rng = np.random.default_rng(seed)
    gh, gw = int(round(GLYPH_H * glyph_scale)), int(round(GLYPH_W * glyph_scale))
    glyphs = [cv2.resize(b * 255, (gw, gh), interpolation=cv2.INTER_LINEAR) > 127 for b in glyph_bitmaps()]
    g = np.clip(rng.normal(215, 12, (H, W)), 0, 255).astype(np.uint8)
    n = int(rng.integers(n_lines[0], n_lines[1] + 1))
    pitch = H / (n + 1)
    boxes, texts = [], []
    for li in range(n):
        y = int(pitch * (li + 0.5) + rng.integers(-4, 5))
        x0 = int(rng.integers(W // 40, W // 10))
        x1 = W - int(rng.integers(W // 40, W // 10))
        x, ids = x0, []
        while x + gw < x1:
            gid = int(rng.integers(0, 76))
            dy = int(rng.integers(-2, 3))
            ys = min(max(0, y + dy), H - gh)
            ink = glyphs[gid]
            patch = g[ys:ys + gh, x:x + gw]
            patch[ink] = np.clip(rng.normal(40, 12, int(ink.sum())), 0, 255).astype(np.uint8)
            ids.append(gid)
            x += gw + int(rng.integers(int(5 * glyph_scale), int(12 * glyph_scale)))
        boxes.append((x0, y, x - x0, gh))
        texts.append(ids)
    g = cv2.GaussianBlur(g, (3, 3), 0)
    return np.ascontiguousarray(np.stack([g, g, g], axis=-1)), boxes, texts
