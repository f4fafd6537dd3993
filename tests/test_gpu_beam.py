"""Beam-search CTC decode and eval mode: the GPU candidate kernel + host prefix merge against the
numpy restatement of the same spec (oracle/ctc_beam.py).  PARITY UNPINNED against the reference's decoder (an
un-vendored pyctcdecode fork): the spec is the published pyctcdecode algorithm without a language model."""
import numpy as np
import pytest
import torch

from ocr_inference_b200 import synthetic as synth
from ocr_inference_b200.data import CharsetEncoder, LineDetectionConfig, OCRModelConfig, OpStatus
from ocr_inference_b200.engine import LAYOUT_TV, LAYOUT_VT
from ocr_inference_b200.synthetic_weights import easter2_weights_fitted, seg_weights_fitted
from oracle import ctc_beam as ob
from oracle import hostops as ho

pytestmark = pytest.mark.gpu


def _cfg(swap_hw=False, squeeze=True):
    return OCRModelConfig(easter2_weights_fitted(), "Easter2", 2000, 80, "input", "output", squeeze, swap_hw,
                          CharsetEncoder.STACK, synth.CHARSET_76, True, "standin")


@pytest.mark.parametrize("layout", [LAYOUT_TV, LAYOUT_VT])
def test_ctc_topk_candidates_match_numpy(engine, layout):
    rng = np.random.default_rng(3 + layout)
    L, T, V = 5, 137, 77
    logits = rng.normal(0, 2.5, (L, T, V)).astype(np.float32)
    logits[:, ::3, :] *= 0.05                         # flat rows: many candidates above -5 (the cap of 16 bites)
    logits[0, 5, 10] = logits[0, 5, 20] = logits[0, 5].max() + 1.0      # an exact tie at the top: lower index first
    logits[1, 7, :] = 0.0                             # all equal
    x = logits if layout == LAYOUT_TV else np.ascontiguousarray(logits.transpose(0, 2, 1))
    tok, lp, cnt = (t.cpu().numpy() for t in engine.ctc_topk(torch.from_numpy(x).cuda(), layout=layout))
    ref_lp = ob.log_softmax(logits)
    for l in range(L):
        for t in range(T):
            want = ob.frame_candidates(ref_lp[l, t])
            n = int(cnt[l, t])
            # a candidate whose log p sits within float rounding of -5 may be in one list and not in the other
            edge = np.abs(ref_lp[l, t] - (-5.0)) < 1e-4
            if n != len(want):
                assert edge.any(), (l, t, n, len(want))
                n = min(n, len(want))
            assert tok[l, t, :n].tolist() == want[:n], (l, t)
            assert np.allclose(lp[l, t, :n], ref_lp[l, t, want[:n]], atol=2e-5)
    assert tok[0, 5, 0] == 10 and tok[0, 5, 1] == 20
    assert cnt[1, 7] == 16 and tok[1, 7].tolist() == list(range(16))


def test_beam_decode_matches_the_spec_and_greedy_on_clean_lines(engine):
    from ocr_inference_b200.inference import OCRInference

    ocr = OCRInference(_cfg(), None, engine=engine)
    crops, truth = synth.glyph_lines(777, 12)
    beams = ocr.run_beam_batch(crops, pre_pad=False)
    greedy = ocr.run_batch(crops, pre_pad=False)
    dev, hw, off = engine.pack_crops(crops)
    lines, _ = engine.ocr_prepare_lines(dev, hw, off, len(crops), 80, 2000, False)
    logits = engine.ocr_forward(ocr.ocr_session, lines).cpu().numpy()                # [L, T, V]
    for i in range(len(crops)):
        want = ob.decode_beams(logits[i])
        got = beams[i]
        assert got[0].text == "".join(ocr.ctc_decoder.ctc_vocab[k] for k in want[0][0])
        assert abs(got[0].logit_score - want[0][1]) < 1e-2
        assert [b.text for b in got[:5]] == ["".join(ocr.ctc_decoder.ctc_vocab[k] for k in w[0]) for w in want[:5]]
        assert np.allclose([b.logit_score for b in got[:5]], [w[1] for w in want[:5]], atol=1e-2)
        assert got[0].text == greedy[i]                 # a peaky model: best beam = best path
        assert all(a.logit_score >= b.logit_score for a, b in zip(got, got[1:]))
    right = sum(b[0].text == "".join(synth.CHARSET_76[k] for k in t) for b, t in zip(beams, truth))
    assert right >= 11
    # the noisy case: soft logits where merging paths matters -- product vs spec on synthetic logits
    rng = np.random.default_rng(9)
    soft = rng.normal(0, 1.2, (120, 77)).astype(np.float32)
    soft[np.arange(120), rng.integers(0, 77, 120)] += 2.5
    got = ocr.ctc_decoder.ctc_beam_decode(soft)
    want = ob.decode_beams(soft)
    assert [b.text for b in got[:3]] == ["".join(ocr.ctc_decoder.ctc_vocab[k] for k in w[0]) for w in want[:3]]
    assert np.allclose([b.logit_score for b in got[:3]], [w[1] for w in want[:3]], atol=1e-2)
    got_vt = ocr.ctc_decoder.ctc_beam_decode(np.ascontiguousarray(soft.T))            # [V, T] input is transposed
    assert [b.text for b in got_vt[:3]] == [b.text for b in got[:3]]


def test_input_layout_switches_and_vt_logits(engine):
    """swap_hw / squeeze_channel (BDRC/inference.py:422-426) and the [V,T] -> [T,V] rule (:373-376)."""
    from ocr_inference_b200.inference import OCRInference

    crop = synth.glyph_lines(31, 1)[0][0]
    x = ho.prepare_ocr_line(crop, 80, 2000, pre_pad=False, closed_form=False)           # [1, 80, 2000]
    base = OCRInference(_cfg(), None, engine=engine)
    ref = base._predict(x)
    assert ref.shape == (500, 77)
    swapped = OCRInference(_cfg(swap_hw=True), None, engine=engine)._predict(np.transpose(x, (0, 2, 1)))
    chan = OCRInference(_cfg(squeeze=False), None, engine=engine)._predict(x[:, None])
    assert np.array_equal(ref, swapped) and np.array_equal(ref, chan)
    with pytest.raises(ValueError):
        OCRInference(_cfg(squeeze=False), None, engine=engine)._predict(x)
    text = base._decode(ref)
    assert text == base._decode(np.ascontiguousarray(ref.T)) == base.run(crop, pre_pad=False)
    assert base._decode_beams(ref)[0].text == text


def test_run_ocr_eval_mode(engine):
    from ocr_inference_b200.inference import OCRPipeline

    pipe = OCRPipeline(_cfg(), LineDetectionConfig("", seg_weights_fitted(11), "deeplabv3plus-r34", 512, ["line"]), None,
                       engine=engine)
    page, _, truth = synth.text_page(5100, 1024, 1536, n_lines=(3, 3))
    st, res = pipe.run_ocr(page)
    st2, res2 = pipe.run_ocr(page, eval_mode=True)
    assert st == st2 == OpStatus.SUCCESS
    assert len(res2[2]) == len(res[2]) == 3
    for a, b in zip(res[2], res2[2]):
        assert b.text == a.text.replace(" ", "") or b.text == a.text
        assert 0.0 < b.ctc_conf <= 1.0 + 1e-6 and b.ctc_conf > 0.5
        assert 1 <= len(b.logits) <= 10 and all(x >= y for x, y in zip(b.logits, b.logits[1:]))
        assert a.ctc_conf is None and a.logits is None
    batch = pipe.run_ocr_batch([page, page], eval_mode=True)
    assert [l.text for l in batch[1][1][2]] == [l.text for l in res2[2]]
