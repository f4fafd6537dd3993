"""CTC beam-search decode: the host half (prefix merge) of CTCDecoder.ctc_beam_decode.

The reference decodes through `pyctcdecode.build_ctcdecoder(vocab).decode_beams(logits)` from an UNPINNED fork
(BDRC/requirements.txt:13, BDRC/inference.py:75-118) that is absent from /root/reference: parity with it is
unpinned.  This module implements the published pyctcdecode algorithm without a language model, which is what
the reference runs when no KenLM file is configured:

  * per frame, log-softmax of the logits; only tokens with log p >= token_min_logp (-5) take part, the arg-max
    always does (the GPU front end bdrc_ctc_topk also caps them at 16 per frame, in descending order);
  * a beam is (emitted token ids, last frame's token); a frame extends every beam with every candidate:
    blank -> same text, a repeat of the last token -> same text, anything else -> text + token;
    beams with the same (text, last token) merge by log-sum-exp, are sorted by score and cut to `beam_width`
    (100) and to `beam_prune_logp` (-10) below the best;
  * at the end beams with the same text merge, and the result is sorted by score.

Blank: CTCDecoder puts "<blk>" at index 0 when add_blank is set (BDRC/inference.py:88-91) -- that index is the
CTC blank here.  OutputBeam mirrors the fields the reference reads (text, logit_score; lm_score equals logit_score
without a language model).  ctc_conf of run_text_recognition_eval is exp(logit_score / max(len(text), 1)).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

BEAM_WIDTH = 100
BEAM_PRUNE_LOGP = -10.0
TOKEN_MIN_LOGP = -5.0


@dataclass
class OutputBeam:
    text: str
    last_lm_state: object
    text_frames: list = field(default_factory=list)
    logit_score: float = 0.0
    lm_score: float = 0.0


def _lse(a: float, b: float) -> float:
    if a < b:
        a, b = b, a
    return a + math.log1p(math.exp(b - a)) if b > -math.inf else a


def merge_frames(tokens: np.ndarray, logps: np.ndarray, counts: np.ndarray, blank: int = 0,
                 beam_width: int = BEAM_WIDTH, beam_prune_logp: float = BEAM_PRUNE_LOGP) -> list[tuple[tuple[int, ...], float]]:
    """tokens / logps [T, K], counts [T] (one line's candidates, descending per frame) -> [(ids, score)] sorted."""
    beams: dict[tuple[tuple[int, ...], int], float] = {((), -1): 0.0}
    T = int(counts.shape[0])
    tok_l, lp_l, cnt_l = tokens.tolist(), logps.tolist(), counts.tolist()
    for t in range(T):
        n = cnt_l[t]
        cand = list(zip(tok_l[t][:n], lp_l[t][:n]))
        new: dict[tuple[tuple[int, ...], int], float] = {}
        for (text, last), score in beams.items():
            for tk, lp in cand:
                if tk == blank or tk == last:
                    key = (text, tk)
                else:
                    key = (text + (tk,), tk)
                s = score + lp
                old = new.get(key)
                new[key] = s if old is None else _lse(old, s)
        best = max(new.values())
        kept = sorted(((k, v) for k, v in new.items() if v >= best + beam_prune_logp), key=lambda kv: (-kv[1], kv[0]))
        beams = dict(kept[:beam_width])
    final: dict[tuple[int, ...], float] = {}
    for (text, _), score in beams.items():
        old = final.get(text)
        final[text] = score if old is None else _lse(old, score)
    best = max(final.values())
    out = sorted(((k, v) for k, v in final.items() if v >= best + beam_prune_logp), key=lambda kv: (-kv[1], kv[0]))
    return out[:beam_width]


def beams_to_output(beams, vocab: list[str], blank: int = 0) -> list[OutputBeam]:
    out = []
    for ids, score in beams:
        text = "".join(vocab[i] for i in ids if i != blank)
        out.append(OutputBeam(text, None, [], float(score), float(score)))
    return out
