"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the beam-search decode spec of
ocr_inference_b200/ctc_beam.py, from the raw logits (its own log-softmax and candidate selection).

PARITY UNPINNED: the reference's decoder is an un-vendored, unpinned pyctcdecode fork (BDRC/requirements.txt:13);
this follows the published pyctcdecode algorithm without a language model (beam_width 100, beam_prune_logp -10,
token_min_logp -5, at most 16 candidates per frame) with "<blk>" (index 0) as the CTC blank.  It is written
differently from the product on purpose: frame-synchronous lists and np.logaddexp, no dictionaries keyed like
the product's.
"""
from __future__ import annotations

import numpy as np


def log_softmax(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.float32)
    m = x.max(axis=-1, keepdims=True)
    return (x - m) - np.log(np.exp(x - m).sum(axis=-1, keepdims=True, dtype=np.float32))


def frame_candidates(row_logp: np.ndarray, min_logp: float = -5.0, cap: int = 16):
    order = np.lexsort((np.arange(len(row_logp)), -row_logp))          # descending value, ties -> lower index
    keep = [int(order[0])]
    for i in order[1:cap]:
        if row_logp[i] >= min_logp:
            keep.append(int(i))
        else:
            break
    return keep


def decode_beams(logits_tv: np.ndarray, blank: int = 0, beam_width: int = 100, prune: float = -10.0,
                 min_logp: float = -5.0):
    """logits [T, V] -> [(tuple of token ids, score)] best first."""
    lp = log_softmax(logits_tv)
    texts, lasts, scores = [()], [-1], [0.0]
    for t in range(lp.shape[0]):
        cand = frame_candidates(lp[t], min_logp)
        nt, nl, ns = [], [], []
        for text, last, sc in zip(texts, lasts, scores):
            for c in cand:
                nt.append(text if (c == blank or c == last) else text + (c,))
                nl.append(c)
                ns.append(sc + float(lp[t, c]))
        # merge identical (text, last)
        merged = {}
        for text, last, sc in zip(nt, nl, ns):
            k = (text, last)
            merged[k] = np.logaddexp(merged[k], sc) if k in merged else sc
        items = sorted(merged.items(), key=lambda kv: (-kv[1], kv[0]))
        top = items[0][1]
        items = [it for it in items if it[1] >= top + prune][:beam_width]
        texts, lasts, scores = [k[0] for k, _ in items], [k[1] for k, _ in items], [float(v) for _, v in items]
    final = {}
    for text, sc in zip(texts, scores):
        final[text] = np.logaddexp(final[text], sc) if text in final else sc
    items = sorted(final.items(), key=lambda kv: (-kv[1], kv[0]))
    top = items[0][1]
    return [(k, float(v)) for k, v in items if v >= top + prune][:beam_width]
