"""TEST-FIXTURE GENERATOR -- trains the two stand-in models whose tensors live in tests/golden/fitted_heads.npz.

    gpurun -- python tests/golden/train_standins.py --out gpurun_out/fitted_heads.npz      (one B200, ~15 min)
    python tests/golden/train_standins.py --dry                                            (CPU, tiny, checks the script)

Why it exists.  BASELINE.json's parity target is ">= 95 % identical decoded lines vs the reference CPU path" on
whole pages.  The reference ships no checkpoints and there is no network, so the two networks run on synthetic
weights; whole-page identity only means something when those weights behave like trained ones: a line model
whose mask is one clean blob per text line, and a recogniser that actually reads the crops the pipeline hands
it.  This script produces such weights with STOCK torch (cuDNN autograd) on the GPU box:

  1. segmentation: the ResNet-34 body keeps the seeded random init (synthetic_weights.seg_weights(11)); the whole
     DeepLabV3+ decoder (ASPP, block1, block2, head: 1.5 M parameters) is trained against the thinned text-band
     mask of synthetic.text_page() / synthetic.page() pages;
  2. line crops: the pipeline's own extraction (detect_lines -> build_lines -> extract_lines with the new mask
     model, two (k_factor, bbox_tolerance) settings) cuts the lines of a few thousand text pages; reading order
     gives each crop its transcription;
  3. recogniser: Easter2 (6.2 M parameters) is trained with CTC on those crops plus synthetic.glyph_line() crops
     (config C1), with +-3 px shift and salt noise augmentation so that a crop moved by one pixel reads the same.

Nothing here is product code or a parity check: the product path and the oracle both LOAD the tensors this
script writes (synthetic_weights.seg_weights_fitted / easter2_weights_fitted).  Weights are rounded to fp16
before the final evaluation, i.e. the numbers printed at the end belong to the committed tensors.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from multiprocessing import get_context

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from ocr_inference_b200 import synthetic  # noqa: E402
from ocr_inference_b200.synthetic_weights import (R34_LAYERS, easter2_weights, easter2_weights_fitted,  # noqa: E402
                                                  seg_weights)

SEG_SEED, OCR_SEED, V = 11, 7, 77
SEG_SIZES = [(2048, 3072), (1024, 1536), (1536, 2048), (1024, 3072), (2048, 2048)]
T0 = time.time()


def log(*a):
    print(f"[{time.time() - T0:7.1f}s]", *a, flush=True)


# ------------------------------------------------------------------------------------------------
# page generators (worker processes)
# ------------------------------------------------------------------------------------------------
def _gen_seg_page(job):
    """(seed, H, W) -> (gray u8 [H,W], target u8 {0,1} [H,W]).  Even seeds: glyph text pages with varied line
    counts / glyph sizes, odd seeds: the stroke-band pages of config C2."""
    seed, H, W = job
    rng = np.random.default_rng(seed + 77)
    if seed % 2 == 0:
        scale = float(rng.uniform(1.2, 1.9))
        nmax = max(3, int(H / (synthetic.GLYPH_H * scale * 2.6)))
        n = int(rng.integers(max(2, nmax // 2), nmax + 1))
        page, boxes, _ = synthetic.text_page(seed, H, W, n_lines=(n, n), glyph_scale=scale)
    else:
        page, boxes = synthetic.page(seed, H, W)
    m = synthetic.band_mask(boxes, H, W, grow=2, thin=0.45)
    return page[:, :, 0].copy(), (m > 0).astype(np.uint8)


def _gen_text_page(seed):
    page, boxes, texts = synthetic.text_page(seed)
    return seed, page[:, :, 0].copy(), boxes, texts


def _gen_glyph_line(job):
    seed, tall = job
    rng = np.random.default_rng(seed + 5)
    if tall:
        img, ids = synthetic.glyph_line(seed, h=int(rng.integers(64, 191)))
    else:
        img, ids = synthetic.glyph_line(seed)
    return img, ids


_POOL = None


def _worker_init():
    import cv2

    cv2.setNumThreads(1)


def start_pool(procs: int):
    """One pool for the whole run, forked BEFORE the parent touches CUDA or cv2's thread pool."""
    global _POOL
    if procs > 1:
        _POOL = get_context("fork").Pool(procs, initializer=_worker_init)


def pool_map(fn, jobs, procs=None):
    if _POOL is None:
        return [fn(j) for j in jobs]
    return _POOL.map(fn, jobs, chunksize=max(1, len(jobs) // 64))


def rgb(gray):
    return np.ascontiguousarray(np.repeat(gray[:, :, None], 3, axis=2))


# ------------------------------------------------------------------------------------------------
# differentiable restatements of the two graphs (same graphs as oracle/nets.py, checked in --dry)
# ------------------------------------------------------------------------------------------------
def _c2(P, name, x, stride=1, pad=0, dil=1, groups=1):
    return F.conv2d(x, P[name + ".weight"], P.get(name + ".bias"), stride=stride, padding=pad, dilation=dil, groups=groups)


@torch.no_grad()
def seg_encoder(P, x):
    """x fp32 [N,3,512,512] in {0,1} -> (layer-1 skip [N,64,128,128], layer-4 output [N,512,32,32])."""
    x = F.relu(_c2(P, "encoder.conv1", x, stride=2, pad=3))
    x = F.max_pool2d(x, 3, 2, 1)
    skip = None
    for li, (ch, blocks, stride, dil) in enumerate(R34_LAYERS, start=1):
        for b in range(blocks):
            p = f"encoder.layer{li}.{b}"
            s = stride if b == 0 else 1
            y = F.relu(_c2(P, p + ".conv1", x, stride=s, pad=dil, dil=dil))
            y = _c2(P, p + ".conv2", y, stride=1, pad=dil, dil=dil)
            idt = _c2(P, p + ".downsample.0", x, stride=s) if (p + ".downsample.0.weight") in P else x
            x = F.relu(y + idt)
        if li == 1:
            skip = x
    return skip, x


def seg_decoder(P, skip, x, lowres=False):
    br = [F.relu(_c2(P, "decoder.aspp.0.convs.0.0", x))]
    for i, r in zip((1, 2, 3), (12, 24, 36)):
        d = _c2(P, f"decoder.aspp.0.convs.{i}.0.0", x, pad=r, dil=r, groups=512)
        br.append(F.relu(_c2(P, f"decoder.aspp.0.convs.{i}.0.1", d)))
    g = x.mean(dim=(2, 3), keepdim=True)
    g = F.relu(_c2(P, "decoder.aspp.0.convs.4.1", g))
    br.append(F.interpolate(g, size=x.shape[2:], mode="bilinear", align_corners=False))
    x = F.relu(_c2(P, "decoder.aspp.0.project.0", torch.cat(br, dim=1)))
    x = _c2(P, "decoder.aspp.1.0", x, pad=1, groups=256)
    x = F.relu(_c2(P, "decoder.aspp.1.1", x))
    x = F.interpolate(x, scale_factor=4, mode="bilinear", align_corners=True)
    hi = F.relu(_c2(P, "decoder.block1.0", skip))
    x = torch.cat([x, hi], dim=1)
    x = _c2(P, "decoder.block2.0.0", x, pad=1, groups=304)
    x = F.relu(_c2(P, "decoder.block2.0.1", x))
    x = _c2(P, "segmentation_head.0", x)
    if lowres:
        return x
    return F.interpolate(x, scale_factor=4, mode="bilinear", align_corners=True)


def _c1(P, name, x, stride=1, pad=0, dil=1):
    return F.conv1d(x, P[name + ".weight"], P[name + ".bias"], stride=stride, padding=pad, dilation=dil)


def easter2(P, x):
    """x [B,80,W] in {-1,+1} -> logits [B,V,W/4]."""
    x = F.relu(_c1(P, "conv1", x, 2, 1))
    x = F.relu(_c1(P, "conv2", x, 2, 1))
    old = data = x
    for u, k in zip((1, 2, 3), (5, 7, 9)):
        p = f"unit{u}"
        res = _c1(P, p + ".res_old", old) + _c1(P, p + ".res_data", data)
        d = F.relu(_c1(P, p + ".conv_a", data, pad=k // 2))
        d = F.relu(_c1(P, p + ".conv_b", d, pad=k // 2))
        d = _c1(P, p + ".conv_c", d, pad=k // 2)
        z = F.relu(F.linear(d.mean(dim=2), P[p + ".se_fc1.weight"], P[p + ".se_fc1.bias"]))
        g = torch.sigmoid(F.linear(z, P[p + ".se_fc2.weight"], P[p + ".se_fc2.bias"]))
        data = F.relu(res + d * g.unsqueeze(-1))
        old = res
    x = F.relu(_c1(P, "conv3", data, pad=10, dil=2))
    return _c1(P, "head", F.relu(_c1(P, "conv4", x)))


def greedy_ids(logits_bvt: torch.Tensor) -> list[list[int]]:
    am = logits_bvt.argmax(dim=1).cpu().numpy()
    out = []
    for row in am:
        keep = np.ones(len(row), bool)
        keep[1:] = row[1:] != row[:-1]
        r = row[keep]
        out.append([int(v) for v in r[r != 0]])
    return out


# ------------------------------------------------------------------------------------------------
# backends for the pipeline stages around the networks: product kernels on the GPU box, oracle in --dry
# ------------------------------------------------------------------------------------------------
class Backend:
    def __init__(self, dry: bool):
        self.dry = dry
        if not dry:
            from ocr_inference_b200.inference import get_engine

            self.eng = get_engine(0)

    def tiles(self, gray: np.ndarray) -> np.ndarray:
        """binarised u8 {0,1} tiles [N,512,512] of a page whose sides are multiples of 512."""
        if self.dry:
            from oracle import hostops as ho

            return ho.seg_preprocess(rgb(gray), closed_form=False)[0]
        t, _ = self.eng.seg_preprocess(torch.from_numpy(rgb(gray)).cuda())
        torch.cuda.synchronize()
        return t.cpu().numpy()

    def pipeline(self, seg_w: dict, ocr_w: dict):
        from ocr_inference_b200.data import CharsetEncoder, LineDetectionConfig, OCRModelConfig
        from ocr_inference_b200.inference import OCRPipeline

        ocr_cfg = OCRModelConfig(ocr_w, "Easter2", 2000, 80, "input", "output", True, False, CharsetEncoder.WYLIE,
                                 synthetic.CHARSET_76, True, "standin")
        return OCRPipeline(ocr_cfg, LineDetectionConfig("", seg_w, "deeplabv3plus-r34", 512, ["line"]), None,
                           engine=self.eng)

    def page_crops(self, pipe, seg_w, gray, k_factor, bbox_tolerance):
        """-> (line boxes (x,y,w,h) in reading order, crops u8 HWC) or None."""
        page = rgb(gray)
        if self.dry:
            from oracle import geometry as og
            from oracle import hostops as ho
            from oracle import nets

            tiles, meta = ho.seg_preprocess(page, closed_form=False)
            lg = np.concatenate([nets.seg_forward(seg_w, np.repeat(tiles[i:i + 8, None].astype(np.float32), 3, axis=1))
                                 for i in range(0, len(tiles), 8)])
            mask = ho.line_postprocess(lg, meta, 0.9, closed_form=False)
            r = og.page_to_lines(page, mask, k_factor, bbox_tolerance)
            if isinstance(r, str):
                return None
            return [tuple(l.bbox) for l in r[1]], r[2]
        from ocr_inference_b200.data import OpStatus

        st, mask = pipe.detect_lines(page)
        st, r = pipe.build_lines(page, mask)
        if st != OpStatus.SUCCESS:
            return None
        rot_img, rot_mask, _, filt, _ = r
        st, r = pipe.extract_lines(rot_img, rot_mask, filt, True, k_factor, bbox_tolerance)
        if st != OpStatus.SUCCESS:
            return None
        lines, imgs = r
        return [(l.bbox.x, l.bbox.y, l.bbox.w, l.bbox.h) for l in lines], [np.asarray(i) for i in imgs]

    def prep(self, crops: list[np.ndarray], pre_pad: bool) -> np.ndarray:
        """a-7 on a list of crops -> int8 [L,80,2000] in {-1,+1}."""
        if self.dry:
            from oracle import hostops as ho

            return np.stack([ho.prepare_ocr_line(c, 80, 2000, pre_pad=pre_pad, closed_form=False)[0] for c in crops]
                            ).astype(np.int8)
        out = []
        for i in range(0, len(crops), 256):
            dev, hw, off = self.eng.pack_crops(crops[i:i + 256], key="train_crops")
            _, f32 = self.eng.ocr_prepare_lines(dev, hw, off, len(crops[i:i + 256]), 80, 2000, pre_pad, want_f32=True)
            torch.cuda.synchronize()
            out.append(f32.to(torch.int8).cpu().numpy())
        return np.concatenate(out)


def assign_truth(line_boxes, band_boxes, texts):
    """Reading-order lines -> transcriptions by nearest band centre; None unless it is a bijection."""
    if len(line_boxes) != len(band_boxes):
        return None
    cy = np.array([y + h / 2 for (_, y, _, h) in band_boxes])
    pick = [int(np.argmin(np.abs(cy - (y + h / 2)))) for (_, y, _, h) in line_boxes]
    if sorted(pick) != list(range(len(band_boxes))):
        return None
    for (x, y, w, h), j in zip(line_boxes, pick):
        bx, _, bw, _ = band_boxes[j]
        if abs(x - bx) > 60 or abs((x + w) - (bx + bw)) > 60:      # a line that lost an end is not a clean sample
            return None
    return [texts[j] for j in pick]


# ------------------------------------------------------------------------------------------------
# stage 1: segmentation decoder
# ------------------------------------------------------------------------------------------------
SEG_TRAINABLE = ("decoder.", "segmentation_head.")


def round_fp16(P, keys):
    with torch.no_grad():
        for k in keys:
            if k.endswith(".weight"):
                P[k].copy_(P[k].half().float())


def train_seg(be: Backend, dev, n_pages: int, minutes: float, procs: int):
    w0 = seg_weights(SEG_SEED, 1, 3)
    P = {k: torch.from_numpy(v.copy()).to(dev) for k, v in w0.items()}
    tkeys = [k for k in P if k.startswith(SEG_TRAINABLE)]
    with torch.no_grad():
        P["segmentation_head.0.weight"].mul_(0.05)
    for k in tkeys:
        P[k].requires_grad_(True)
    rs = np.random.default_rng(99)
    jobs = [(50_000 + i, *SEG_SIZES[int(rs.integers(0, len(SEG_SIZES)))]) for i in range(n_pages)]
    if be.dry:
        jobs = [(50_000 + i, 512, 1024) for i in range(n_pages)]
    log(f"seg: generating {len(jobs)} pages with {procs} processes")
    pages = pool_map(_gen_seg_page, jobs, procs)
    skips, feats, tgts = [], [], []
    for gray, tgt in pages:
        tiles = be.tiles(gray)
        H, W = gray.shape
        t = tgt.reshape(H // 512, 512, W // 512, 512).transpose(0, 2, 1, 3).reshape(-1, 512, 512)
        assert t.shape[0] == tiles.shape[0]
        for i in range(0, len(tiles), 32):
            x = torch.from_numpy(tiles[i:i + 32]).to(dev).float()[:, None].expand(-1, 3, -1, -1).contiguous()
            s, f = seg_encoder(P, x)
            skips.append(s.half())
            feats.append(f.half())
        tgts.append(torch.from_numpy(t).to(dev))
    skips, feats, tgts = torch.cat(skips), torch.cat(feats), torch.cat(tgts)
    n = skips.shape[0]
    n_val = max(2, n // 20)
    log(f"seg: {n} tiles cached ({(skips.numel() + feats.numel()) * 2 / 1e9:.1f} GB fp16), {n_val} held out, "
        f"positives {tgts.float().mean().item():.3f}")
    opt = torch.optim.Adam([P[k] for k in tkeys], lr=1e-3)
    bs = 4 if be.dry else 24
    t_end = time.time() + minutes * 60
    t_start = time.time()
    it = 0
    g = torch.Generator(device="cpu").manual_seed(1)
    while time.time() < t_end:
        frac = (time.time() - t_start) / (minutes * 60)
        for grp in opt.param_groups:
            grp["lr"] = 1e-3 * (0.02 + 0.98 * 0.5 * (1 + np.cos(np.pi * min(1.0, frac))))
        idx = torch.randint(0, n - n_val, (bs,), generator=g).to(dev)
        z = seg_decoder(P, skips[idx].float(), feats[idx].float())[:, 0]
        y = tgts[idx].float()
        loss = F.binary_cross_entropy_with_logits(z, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([P[k] for k in tkeys], 5.0)
        opt.step()
        if it % 200 == 0:
            log(f"seg it {it} bce {loss.item():.4f} lr {opt.param_groups[0]['lr']:.2e}")
        it += 1
        if be.dry and it >= 6:
            break
    round_fp16(P, tkeys)
    inter = union = 0
    with torch.no_grad():
        for i in range(n - n_val, n, 16):
            z = seg_decoder(P, skips[i:i + 16].float(), feats[i:i + 16].float())[:, 0]
            m = z > 2.1972246
            y = tgts[i:i + 16] > 0
            inter += (m & y).sum().item()
            union += (m | y).sum().item()
    log(f"seg: {it} steps; held-out IoU at sigmoid > 0.9: {inter / max(1, union):.4f}")
    out = {k: P[k].detach().cpu().numpy() for k in tkeys}
    full = dict(w0)
    full.update(out)
    return out, full, inter / max(1, union)


# ------------------------------------------------------------------------------------------------
# stage 2: line crops through the pipeline's own extraction
# ------------------------------------------------------------------------------------------------
def collect_lines(be: Backend, pipe, seg_w, seeds, procs, settings=((2.5, 4.0), (2.0, 2.5))):
    xs, ys, stats = [], [], {"pages": 0, "pages_clean": 0, "lines": 0, "contours": 0}
    for c0 in range(0, len(seeds), 96):
        chunk = pool_map(_gen_text_page, seeds[c0:c0 + 96], procs)
        crops, labels = [], []
        for seed, gray, boxes, texts in chunk:
            stats["pages"] += 1
            clean = True
            for kf, tol in settings:
                r = be.page_crops(pipe, seg_w, gray, kf, tol)
                tr = assign_truth(r[0], boxes, texts) if r is not None else None
                if tr is None:
                    clean = False
                    continue
                for c, t in zip(r[1], tr):
                    if c.shape[0] > 0 and c.shape[1] > 0:
                        crops.append(np.ascontiguousarray(c))
                        labels.append([k + 1 for k in t])
            stats["pages_clean"] += int(clean)
        if crops:
            xs.append(be.prep(crops, pre_pad=False))
            ys.extend(labels)
            stats["lines"] += len(labels)
        log(f"crops: {stats}")
    return (np.concatenate(xs) if xs else np.zeros((0, 80, 2000), np.int8)), ys, stats


# ------------------------------------------------------------------------------------------------
# stage 3: recogniser
# ------------------------------------------------------------------------------------------------
def train_ocr(be: Backend, dev, X: np.ndarray, Y: list, Xv: np.ndarray, Yv: list, minutes: float, ckpt=None):
    try:
        w0 = easter2_weights_fitted()
    except Exception:  # noqa: BLE001
        w0 = easter2_weights(OCR_SEED, V, 80)
    P = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(dev).requires_grad_(True) for k, v in w0.items()}
    keys = list(P)
    Xd = torch.from_numpy(X).to(dev)
    Xvd = torch.from_numpy(Xv).to(dev)
    n = Xd.shape[0]
    bs = 4 if be.dry else 48
    opt = torch.optim.Adam([P[k] for k in keys], lr=6e-4)
    g = torch.Generator(device="cpu").manual_seed(2)
    rs = np.random.default_rng(3)
    t_start = time.time()
    t_end = t_start + minutes * 60
    it, last_ck = 0, time.time()

    def evaluate(Xe, Ye):
        ok = 0
        with torch.no_grad():
            for i in range(0, Xe.shape[0], 64):
                ids = greedy_ids(easter2(P, Xe[i:i + 64].float()))
                ok += sum(a == b for a, b in zip(ids, Ye[i:i + 64]))
        return ok / max(1, Xe.shape[0])

    while time.time() < t_end:
        frac = (time.time() - t_start) / (minutes * 60)
        for grp in opt.param_groups:
            grp["lr"] = 6e-4 * (0.02 + 0.98 * 0.5 * (1 + np.cos(np.pi * min(1.0, frac))))
        idx = torch.randint(0, n, (bs,), generator=g)
        x = Xd[idx.to(dev)].float()
        dy, dx = int(rs.integers(-3, 4)), int(rs.integers(-8, 9))
        if dy or dx:
            x = torch.roll(x, shifts=(dy, dx), dims=(1, 2))
        if rs.random() < 0.7:
            flip = torch.rand(x.shape, device=dev) < 0.004
            x = torch.where(flip, -x, x)
        lp = F.log_softmax(easter2(P, x), dim=1).permute(2, 0, 1)       # [T,B,V]
        tg = torch.tensor([t for j in idx.tolist() for t in Y[j]], dtype=torch.long)
        tl = torch.tensor([len(Y[j]) for j in idx.tolist()], dtype=torch.long)
        loss = F.ctc_loss(lp, tg, torch.full((bs,), lp.shape[0], dtype=torch.long), tl, blank=0, zero_infinity=True)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([P[k] for k in keys], 5.0)
        opt.step()
        if it % 250 == 0:
            log(f"ocr it {it} ctc {loss.item():.4f} lr {opt.param_groups[0]['lr']:.2e}"
                + (f" held-out exact lines {evaluate(Xvd, Yv):.3f}" if it % 1000 == 0 else ""))
        if ckpt and time.time() - last_ck > 150:
            np.savez_compressed(ckpt, **{"ocr." + k: P[k].detach().cpu().numpy().astype(np.float16) for k in keys})
            last_ck = time.time()
        it += 1
        if be.dry and it >= 4:
            break
    with torch.no_grad():
        for k in keys:
            P[k].copy_(P[k].half().float())
    acc = evaluate(Xvd, Yv)
    with torch.no_grad():
        lo = easter2(P, Xvd[:64].float())
        srt = torch.sort(lo, dim=1).values
        margin = (srt[:, -1] - srt[:, -2]).flatten()
    log(f"ocr: {it} steps; held-out exact lines {acc:.4f} of {Xvd.shape[0]}; logit std {lo.std().item():.2f}; "
        f"top-2 margin median {margin.median().item():.2f}, 1st pct {torch.quantile(margin[:100000], 0.01).item():.3f}")
    return {k: P[k].detach().cpu().numpy() for k in keys}, acc


# ------------------------------------------------------------------------------------------------
def save(path, seg_out, ocr_out):
    blob = {}
    for k, v in (seg_out or {}).items():
        blob["seg." + k] = v.astype(np.float16) if k.endswith(".weight") else v.astype(np.float32)
    for k, v in (ocr_out or {}).items():
        blob["ocr." + k] = v.astype(np.float16)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez_compressed(path, **blob)
    log("wrote", path, os.path.getsize(path), "bytes")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fitted_heads.npz"))
    ap.add_argument("--dry", action="store_true")
    ap.add_argument("--seg-pages", type=int, default=420)
    ap.add_argument("--seg-minutes", type=float, default=4.0)
    ap.add_argument("--text-pages", type=int, default=1400)
    ap.add_argument("--glyph-lines", type=int, default=6000)
    ap.add_argument("--ocr-minutes", type=float, default=8.0)
    ap.add_argument("--eval-pages", type=int, default=48)
    ap.add_argument("--seg-from", default=None, help="npz holding seg.* tensors: skip stage 1")
    a = ap.parse_args()
    dry = a.dry
    if dry:
        a.out = "/tmp/fitted_heads_dry.npz"
        a.seg_pages, a.seg_minutes, a.text_pages, a.glyph_lines, a.ocr_minutes, a.eval_pages = 3, 0.2, 2, 6, 0.2, 1
    dev = torch.device("cpu" if dry else "cuda:0")
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    procs = 1 if dry else max(1, min(24, len(os.sched_getaffinity(0)) - 2))
    start_pool(procs)
    be = Backend(dry)
    report = {}

    # 1 ---- segmentation decoder
    if a.seg_from:
        z = np.load(a.seg_from)
        seg_out = {k[4:]: z[k].astype(np.float32) for k in z.files if k.startswith("seg.")}
        seg_full = seg_weights(SEG_SEED, 1, 3)
        seg_full.update(seg_out)
    else:
        seg_out, seg_full, iou = train_seg(be, dev, a.seg_pages, a.seg_minutes, procs)
        report["seg_heldout_iou"] = iou
    save(a.out, seg_out, None)
    if dry:   # the differentiable graphs must be the oracle's graphs
        from oracle import nets

        x = (np.random.default_rng(0).random((1, 3, 512, 512)) > 0.8).astype(np.float32)
        Pt = {k: torch.from_numpy(v) for k, v in seg_full.items()}
        with torch.no_grad():
            mine = seg_decoder(Pt, *seg_encoder(Pt, torch.from_numpy(x))).numpy()
        assert np.allclose(mine, nets.seg_forward(seg_full, x), atol=1e-4), "seg graph differs from oracle/nets.py"
        we = easter2_weights(OCR_SEED, V, 80)
        xe = np.sign(np.random.default_rng(1).standard_normal((2, 80, 400))).astype(np.float32)
        with torch.no_grad():
            mine = easter2({k: torch.from_numpy(v) for k, v in we.items()}, torch.from_numpy(xe)).numpy()
        assert np.allclose(mine, nets.easter2_forward(we, xe), atol=1e-4), "Easter2 graph differs from oracle/nets.py"
        log("graphs equal oracle/nets.py")

    # 2 ---- crops
    try:
        ocr_w0 = easter2_weights_fitted()
    except Exception:  # noqa: BLE001
        ocr_w0 = easter2_weights(OCR_SEED, V, 80)
    pipe = None if dry else be.pipeline(seg_full, ocr_w0)
    X, Y, st = collect_lines(be, pipe, seg_full, [100_000 + i for i in range(a.text_pages)], procs,
                             settings=((2.5, 4.0),) if dry else ((2.5, 4.0), (2.0, 2.5)))
    report["train_crops"] = st
    Xv, Yv, stv = collect_lines(be, pipe, seg_full, [3003 + i for i in range(a.eval_pages)], procs, settings=((2.5, 4.0),))
    report["heldout_crops"] = stv
    # mask quality on the held-out pages: contours per page against lines per page
    log(f"glyph lines: {a.glyph_lines}")
    gl = pool_map(_gen_glyph_line, [(700_000 + i, i % 3 == 0) for i in range(a.glyph_lines)], procs)
    Xg = be.prep([g_[0] for g_ in gl], pre_pad=False)
    Yg = [[k + 1 for k in g_[1]] for g_ in gl]
    glv = pool_map(_gen_glyph_line, [(1001 * 100003 + i, False) for i in range(64 if not dry else 2)], procs)
    Xgv = be.prep([g_[0] for g_ in glv], pre_pad=False)
    Ygv = [[k + 1 for k in g_[1]] for g_ in glv]
    if X.shape[0] == 0:
        log("no clean page crops: recogniser trains on glyph lines only")
    Xall = np.concatenate([X, Xg])
    Yall = Y + Yg
    Xval = np.concatenate([Xv, Xgv])
    Yval = Yv + Ygv
    log(f"recogniser data: {X.shape[0]} page lines + {Xg.shape[0]} glyph lines; held out {Xv.shape[0]} + {Xgv.shape[0]}")

    # 3 ---- recogniser
    ocr_out, acc = train_ocr(be, dev, Xall, Yall, Xval, Yval, a.ocr_minutes, ckpt=a.out.replace(".npz", "_ocr_ckpt.npz"))
    report["ocr_heldout_exact_lines"] = acc
    save(a.out, seg_out, ocr_out)

    # 4 ---- the product pipeline with the tensors just written, whole pages against the truth
    if not dry:
        from ocr_inference_b200.data import OpStatus

        z = np.load(a.out)
        ocr_w = {k[4:]: z[k].astype(np.float32) for k in z.files if k.startswith("ocr.")}
        pipe = be.pipeline(seg_full, ocr_w)
        right = total = found = 0
        for seed in range(3003, 3003 + a.eval_pages):
            page, boxes, texts = synthetic.text_page(seed)
            want = {"".join(synthetic.CHARSET_76[i] for i in t).strip().replace("§", " ") for t in texts}
            st_, res = pipe.run_ocr(page)
            total += len(texts)
            if st_ == OpStatus.SUCCESS:
                found += len(res[2])
                right += sum(l.text in want for l in res[2])
        report["pages_lines_equal_to_truth"] = right / max(1, total)
        report["pages_lines_found"] = found
        report["pages_lines_truth"] = total
        log(f"product pipeline on {a.eval_pages} held-out pages: {right}/{total} lines equal to the truth, {found} found")
    log("REPORT " + json.dumps(report))
    with open(a.out.replace(".npz", "_report.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
