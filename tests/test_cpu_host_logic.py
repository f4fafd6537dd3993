"""Host-side logic of the product package on CPU: geometry mirror vs the oracle, size clamps,
page sharding over a world_size-2 gloo group."""
import os

import cv2
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ocr_inference_b200 import engine as E
from ocr_inference_b200 import line_detection as ld
from ocr_inference_b200 import sharding, synthetic
from oracle import geometry as og
from oracle import hostops as ho


@pytest.mark.parametrize("hw", [(2048, 3072), (472, 2550), (3300, 2550), (700, 4600), (400, 1500), (512, 512), (600, 4096)])
def test_clamp_sizes_match_oracle(hw):
    assert E.clamp_size(*hw) == ho.clamp_size(*hw)
    chw = np.zeros((3, *hw), np.uint8)
    _, meta = ho.torch_path_pre(chw)
    assert E.clamp_size_torch(*hw) == (meta["Hc"], meta["Wc"])


def test_host_geometry_mirror_matches_oracle():
    page, boxes = synthetic.page(700, 800, 1200)
    mask = synthetic.band_mask(boxes, 800, 1200)
    mask[300:340, 580:620] = 0          # split one band into two chunks -> exercises the hull merge
    img, rot_mask, contours, angle = ld.build_raw_line_data(page, mask)
    o_img, o_mask, o_contours, o_angle = og.build_raw_line_data(page, mask)
    assert angle == o_angle and np.array_equal(rot_mask, o_mask) and len(contours) == len(o_contours)
    filt = ld.filter_line_contours(rot_mask, contours)
    assert len(filt) == len(og.filter_line_contours(o_mask, o_contours))
    lines = [ld.build_line_data(c) for c in filt]
    got, thr = ld.sort_lines_by_threshold2(rot_mask, lines, group_lines=True)
    want, othr = og.sort_lines(o_mask, [og.build_line(c) for c in filt], True)
    assert thr == othr and len(got) == len(want) < len(lines)
    assert [g.center for g in got] == [w.center for w in want]
    assert [(g.bbox.x, g.bbox.y, g.bbox.w, g.bbox.h) for g in got] == [w.bbox for w in want]
    assert all(np.array_equal(g.contour, w.contour) for g, w in zip(got, want))
    flat, _ = ld.sort_lines_by_threshold2(rot_mask, lines, group_lines=False)
    oflat, _ = og.sort_lines(o_mask, [og.build_line(c) for c in filt], False)
    assert [g.center for g in flat] == [w.center for w in oflat]


def test_sort_bbox_centers_edge_cases():
    assert ld.sort_bbox_centers([]) == []
    c = [(50, 10), (10, 12), (30, 60), (5, 58), (7, 120)]
    assert ld.sort_bbox_centers(c, 20) == og.sort_centers(c, 20) == [[(7, 120)], [(5, 58), (30, 60)], [(10, 12), (50, 10)]]


def test_sort_and_group_randomised_against_oracle():
    """The O(1) running mean and the first-match dictionary give what the reference's np.mean / list scans give,
    duplicates of a centre included."""
    rng = np.random.default_rng(5)
    for trial in range(40):
        n = int(rng.integers(1, 60))
        cs = [(int(rng.integers(0, 3000)), int(rng.integers(0, 2000) // 7 * 7)) for _ in range(n)]
        if n > 3:
            cs[n // 2] = cs[0]                                   # a repeated centre
        thr = float(rng.integers(0, 60))
        assert ld.sort_bbox_centers(list(cs), thr) == og.sort_centers(list(cs), thr)
        contours = [np.array([[[x - 5, y - 3]], [[x + 5, y - 3]], [[x + 5, y + 3 + i % 3]], [[x - 5, y + 3]]], np.int32)
                    for i, (x, y) in enumerate(cs)]
        lines = [ld.Line(ld.generate_guid(23), c, ld.BBox(*cv2.boundingRect(c)), ctr) for c, ctr in zip(contours, cs)]
        rows = ld.sort_bbox_centers(list(cs), thr)
        got = ld.group_line_chunks(rows, lines)
        # restatement of line_detection.py:489-548 with the reference's list scans
        want = []
        for centers in rows:
            if len(centers) > 1:
                stack = []
                for c in centers:
                    for ln in lines:
                        if c == ln.center:
                            stack.append(ln.contour)
                            break
                want.append(cv2.convexHull(np.vstack(stack)))
            else:
                for c in centers:
                    for ln in lines:
                        if c == ln.center:
                            want.append(ln.contour)
                            break
        assert len(got) == len(want) and all(np.array_equal(g.contour, w) for g, w in zip(got, want))


def _random_band_mask(rng, H, W):
    """Blobs with holes, ragged edges and specks in a few row bands; sometimes touching the image border."""
    m = np.zeros((H, W), np.uint8)
    for _ in range(int(rng.integers(1, 7))):
        y = int(rng.integers(0, H - 20))
        h = int(rng.integers(3, 40))
        x0, x1 = sorted(int(v) for v in rng.integers(0, W, 2))
        band = (cv2.GaussianBlur(rng.normal(0, 1, (h, max(x1 - x0, 1))).astype(np.float32), (0, 0), 2.0) > -0.05)
        m[y:y + h, x0:x0 + band.shape[1]] |= band[: H - y].astype(np.uint8) * 255
    for _ in range(int(rng.integers(0, 12))):
        m[int(rng.integers(0, H)), int(rng.integers(0, W))] = 255
    if rng.random() < 0.3:
        m[0, : W // 3] = 255
    if rng.random() < 0.3:
        m[H - 1, W // 2:] = 255
    if rng.random() < 0.2:
        m[:, 0] = 255                                            # every row occupied: no compaction
    return m


def test_row_compaction_gives_cv2_contours_and_census():
    """findContours / get_line_threshold through RowCompaction equal cv2 on the full mask: same contours, same
    points, same list order, same threshold."""
    rng = np.random.default_rng(12)
    n_compacted = 0
    for trial in range(60):
        H, W = int(rng.integers(64, 500)), int(rng.integers(40, 700))
        m = _random_band_mask(rng, H, W) if trial else np.zeros((H, W), np.uint8)
        want, _ = cv2.findContours(m, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
        rows = ld.RowCompaction(m)
        n_compacted += rows.small is not None
        got = ld.get_contours_compact(m, rows)
        assert len(got) == len(want)
        assert all(g.dtype == w.dtype and np.array_equal(g, w) for g, w in zip(got, want))
        assert ld.get_line_threshold(m) == og.line_threshold(m)
        _, m3, cs, ang = ld.build_raw_line_data(np.zeros((H, W, 3), np.uint8), m)
        _, om, ocs, oang = og.build_raw_line_data(np.zeros((H, W, 3), np.uint8), m)
        assert ang == oang and len(cs) == len(ocs) and all(np.array_equal(a, b) for a, b in zip(cs, ocs))
        assert ld.get_line_threshold(m3) == og.line_threshold(om)
    assert n_compacted > 30


def test_page_geometry_and_geometry_pool_match_stage_functions(monkeypatch):
    """line_detection.page_geometry (the CUDA-free part of stages 2-4 as one call) gives what the stage functions
    give one after the other, directly and through GeometryPool worker processes (mask via shared memory)."""
    from ocr_inference_b200.geometry_pool import GeometryPool

    page, boxes = synthetic.page(702, 800, 1200)
    mask = synthetic.band_mask(boxes, 800, 1200)
    mask[300:340, 580:620] = 0
    _, rot_mask, contours, angle = ld.build_raw_line_data(page, mask)
    filt = ld.filter_line_contours(rot_mask, contours)
    want, _ = ld.sort_lines_by_threshold2(rot_mask, [ld.build_line_data(c) for c in filt], group_lines=True)
    rois_w, masks_w, hs_w = ld.line_rois(want, 800, 1200)

    def check(res):
        assert res[0] == "ok" and res[1] == angle and len(res[2]) == len(want)
        assert all(g is not None for g, _, _, _ in res[2]) and len({g for g, _, _, _ in res[2]}) == len(res[2])
        for (guid, contour, bb, ctr), w in zip(res[2], want):
            assert np.array_equal(contour, w.contour) and bb == (w.bbox.x, w.bbox.y, w.bbox.w, w.bbox.h) and ctr == w.center
        assert res[3] == rois_w and res[5] == hs_w and all(np.array_equal(a, b) for a, b in zip(res[4], masks_w))

    check(ld.page_geometry(mask))
    assert ld.page_geometry(np.zeros((300, 400), np.uint8)) == ("failed", "No lines detected")
    speck = np.zeros((300, 400), np.uint8)
    speck[10:16, 10:16] = 255
    assert ld.page_geometry(speck) == ("failed", "No valid lines after filtering")
    # (with cv2 4.13 minAreaRect reports negative angles for tilted bands, so the reference's angle is never
    #  > 0.5 on such masks; the hand-over of a page that needs rotating is checked with a patched angle)
    with monkeypatch.context() as mp_:
        mp_.setattr(ld, "get_rotation_angle_from_lines", lambda *a, **k: 2.0)
        assert ld.page_geometry(mask) == ("rotate",)
    with GeometryPool(2, slots=3) as pool:
        futs = [pool.submit(m) for m in (mask, speck, mask, mask[:300, :400] * 0, mask, mask)]   # more pages than slots
        res = [f.result(timeout=120) for f in futs]
        for k in (0, 2, 4, 5):
            check(res[k])
        assert res[1] == ("failed", "No valid lines after filtering") and res[3] == ("failed", "No lines detected")
        with pytest.raises(ValueError):
            pool.submit(np.zeros((4, 4, 3), np.uint8))


def test_line_threshold_census_on_wide_pages_matches_oracle():
    """get_line_threshold counts the contours of all vertical slices in ONE findContours call on a strip image; on
    page-wide masks the reference's slice origins x + x_steps * step run past the right edge (empty slices), and a
    column of foreground switches the row compaction off: same threshold as the oracle's slice-by-slice census."""
    rng = np.random.default_rng(7)
    for t in range(6):
        H, W = int(rng.integers(500, 1200)), int(rng.integers(2300, 3400))
        _, boxes = synthetic.page(900 + t, H, W)
        m = synthetic.band_mask(boxes, H, W, thin=float(rng.uniform(0.3, 1.0)))
        for _ in range(150):
            m[int(rng.integers(0, H)), int(rng.integers(0, W))] = 255
        if t % 3 == 0:
            m[:, :5] = 255
        assert ld.get_line_threshold(m) == og.line_threshold(m)
        assert ld.get_line_threshold(cv2.cvtColor(m, cv2.COLOR_GRAY2RGB)) == og.line_threshold(m)


def test_mask3_link_is_the_identity_conversion():
    """BGR2GRAY of a channel-replicated image is the channel (all 256 values), so get_line_threshold may read the
    1-channel mask build_raw_line_data remembers; views and copies drop the link."""
    v = np.arange(256, dtype=np.uint8).reshape(16, 16)
    assert np.array_equal(cv2.cvtColor(cv2.cvtColor(v, cv2.COLOR_GRAY2RGB), cv2.COLOR_BGR2GRAY), v)
    page, boxes = synthetic.page(701, 600, 900)
    mask = synthetic.band_mask(boxes, 600, 900)
    _, m3, _, _ = ld.build_raw_line_data(page, mask)
    assert isinstance(m3, np.ndarray) and m3.shape == (600, 900, 3) and m3.gray is mask
    assert m3[:300].gray is None and m3.copy().gray is None and True
    assert ld.get_line_threshold(m3) == ld.get_line_threshold(np.array(m3)) == og.line_threshold(np.array(m3))


def test_roi_mask_equals_full_page_fill():
    page, boxes = synthetic.page(9, 400, 600)
    mask = synthetic.band_mask(boxes, 400, 600)
    cs, _ = cv2.findContours(mask, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    for c in cs:
        c = ld.optimize_countour(c)
        full = np.zeros((400, 600), np.uint8)
        cv2.drawContours(full, [c], -1, (255, 255, 255), -1)
        (x, y, w, h), roi, bh = ld._roi_mask(c, 400, 600)
        assert bh == cv2.boundingRect(c)[3]
        assert np.array_equal(full[y:y + h, x:x + w], roi)
        assert full.sum() == roi.sum()


def test_se_for_matches_reference_arithmetic():
    for h in (11, 37, 60, 140):
        for k in (2.5, 2.4, 1.7, 0.3, 0.1):
            ks = int(h * k)
            se = ld._se_for(h, k)
            assert se == ((ks, int(ks * k)) if ks >= 1 and int(ks * k) >= 1 else None)


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 16, 10000):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in sharding.shard_range(n, world, r)]
            assert got == list(range(n))
            sizes = [len(sharding.shard_range(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = sharding.run_sharded(11, lambda i: i * 10, lambda pages: [(p, rank) for p in pages], chunk=4)
        if rank == 0:
            q.put(res)
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


def test_page_sharding_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [i * 10 for i in range(11)]
    assert [r[1] for r in res] == [0] * 5 + [1] * 6


def test_pipeline_pool_plumbing_without_gpu():
    """PipelinePool: pages through shared memory to worker processes, results back in input order, FAILED
    tuples and masks-on-request preserved (a stand-in pipeline: the real one needs a B200)."""
    import numpy as np

    from ocr_inference_b200.data import OpStatus
    from ocr_inference_b200.pool import PipelinePool, _echo_factory

    pages = [np.full((4 + i, 6, 3), i, np.uint8) for i in range(7)] + [np.zeros((1, 5, 3), np.uint8)]
    with PipelinePool(devices=(0, 1), procs_per_device=2, factory=_echo_factory, start_timeout=120) as pool:
        assert len(pool) == 4
        out = pool.run_ocr_many(pages, k_factor=2.0)
        out2 = pool.run_ocr_many(pages[:2])
    assert [o[0] for o in out] == [OpStatus.SUCCESS] * 7 + [OpStatus.FAILED]
    assert out[-1][1] == "No lines detected"
    for i, (st, payload) in enumerate(out[:7]):
        assert payload[0] is None and payload[2] == [f"sum={int(pages[i].sum())} k=2.0"] and payload[3] in (0.0, 1.0)
    assert out2[1][1][2] == [f"sum={int(pages[1].sum())} k=None"]
    with PipelinePool(devices=(0,), procs_per_device=1, factory=_echo_factory, return_masks=True, start_timeout=120) as pool:
        st, payload = pool.run_ocr_many(pages[3:4])[0]
        assert np.array_equal(payload[0], pages[3][..., 0])


def test_model_config_readers_match_reference_layout(tmp_path):
    """read_*_model_config: the reference's model_config.json keys -> the config dataclasses (utils.py:162-207,
    349-393); a .bdrcw container next to the named .onnx is preferred."""
    import json

    from ocr_inference_b200 import utils
    from ocr_inference_b200.data import CharsetEncoder

    d = tmp_path / "Woodblock"
    d.mkdir()
    (d / "model_config.json").write_text(json.dumps({
        "onnx-model": "OCRModel.onnx", "architecture": "Easter2", "version": "1.0", "input_width": 2000,
        "input_height": 80, "input_layer": "input", "output_layer": "output", "encoder": "stack",
        "squeeze_channel_dim": "yes", "swap_hw": "no", "charset": "abc§", "add_blank": "yes"}), encoding="utf-8")
    cfg = utils.read_ocr_model_config(str(d / "model_config.json"))
    assert cfg.model_file == f"{d}/OCRModel.onnx" and cfg.squeeze_channel and not cfg.swap_hw and cfg.add_blank
    assert cfg.encoder == CharsetEncoder.STACK and utils.get_charset(cfg.charset) == ["a", "b", "c", "§"]
    (d / "OCRModel.bdrcw").write_bytes(b"x")
    assert utils.read_ocr_model_config(str(d / "model_config.json")).model_file == f"{d}/OCRModel.bdrcw"
    (d / "line.json").write_text(json.dumps({"checkpoint": "c.pth", "onnx-model": "lines.onnx", "architecture": "deeplabv3",
                                              "patch_size": "512", "classes": ["line"]}))
    lc = utils.read_line_model_config(str(d / "line.json"))
    assert (lc.onnx_file, lc.patch_size, lc.classes) == (f"{d}/lines.onnx", 512, ["line"])
    lay = utils.read_layout_model_config(str(d / "line.json"))
    assert lay.architecture == f"{d}/deeplabv3" and lay.patch_size == 512


def test_fold_batchnorm_state_dict_equals_eval_mode_modules():
    """weights.fold_batchnorm_state_dict: conv+BN (plain, separable, with and without conv bias, downsample
    naming) folded into conv weights/bias named like the exported ONNX initializers."""
    import torch
    from torch import nn

    from ocr_inference_b200.weights import fold_batchnorm_state_dict

    torch.manual_seed(0)

    class Sep(nn.Sequential):
        def __init__(self, cin, cout):
            super().__init__(nn.Conv2d(cin, cin, 3, padding=1, groups=cin, bias=False), nn.Conv2d(cin, cout, 1, bias=False))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(3, 8, 3, padding=1, bias=False)
            self.bn1 = nn.BatchNorm2d(8)
            self.block = nn.Sequential(Sep(8, 16), nn.BatchNorm2d(16), nn.ReLU())
            self.downsample = nn.Sequential(nn.Conv2d(16, 16, 1, bias=True), nn.BatchNorm2d(16))
            self.head = nn.Conv2d(16, 2, 1)

        def forward(self, x):
            return self.head(self.downsample(self.block(torch.relu(self.bn1(self.conv1(x))))))

    net = Net()
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.normal_()
            m.running_var.uniform_(0.5, 2.0)
            m.weight.data.normal_(1.0, 0.3)
            m.bias.data.normal_()
    net.eval()
    x = torch.randn(2, 3, 12, 12)
    want = net(x)
    w = {k: torch.from_numpy(v) for k, v in fold_batchnorm_state_dict(net.state_dict()).items()}
    assert set(w) == {"conv1.weight", "conv1.bias", "block.0.0.weight", "block.0.1.weight", "block.0.1.bias",
                      "downsample.0.weight", "downsample.0.bias", "head.weight", "head.bias"}
    F = torch.nn.functional
    y = torch.relu(F.conv2d(x, w["conv1.weight"], w["conv1.bias"], padding=1))
    y = F.conv2d(y, w["block.0.0.weight"], None, padding=1, groups=8)
    y = torch.relu(F.conv2d(y, w["block.0.1.weight"], w["block.0.1.bias"]))
    y = F.conv2d(y, w["downsample.0.weight"], w["downsample.0.bias"])
    y = F.conv2d(y, w["head.weight"], w["head.bias"])
    assert torch.allclose(y, want, atol=1e-5)


def test_modern_layout_weight_source_prefers_a_container_over_the_reference_pth(tmp_path):
    """A model directory laid out like the reference's (checkpoint = torch .pth, onnx-model next to a .bdrcw):
    the container wins; with only the .pth present the checkpoint is named (and folded at load)."""
    from ocr_inference_b200.data import LayoutDetectionConfig
    from ocr_inference_b200.inference import ModernBookFormatLayoutDetection as M
    from ocr_inference_b200.weights import write_blob

    pth, blob = tmp_path / "model.pth", tmp_path / "model.bdrcw"
    pth.write_bytes(b"PK\x03\x04 not a container")
    blob.write_bytes(write_blob({"a": np.zeros((1,), np.float32)}))
    classes = ["background", "line"]
    assert M._weights_source(LayoutDetectionConfig(str(pth), str(blob), "x", 512, classes)) == str(blob)
    assert M._weights_source(LayoutDetectionConfig(str(pth), str(tmp_path / "missing.onnx"), "x", 512, classes)) == str(pth)
    assert M._weights_source(LayoutDetectionConfig(str(blob), "", "x", 512, classes)) == str(blob)
    w = {"k": np.zeros(1, np.float32)}
    assert M._weights_source(LayoutDetectionConfig(w, "", "x", 512, classes)) is w
    with pytest.raises(FileNotFoundError):
        M._weights_source(LayoutDetectionConfig(str(tmp_path / "no.pth"), str(tmp_path / "no.onnx"), "x", 512, classes))


def test_image_dataset_and_page_loader(tmp_path):
    """dataio.ImageInferenceDataset (BDRC/inference.py:829-857) and PageLoader: sorted file order, CHW u8 RGB
    tensors with the decoder's pixels, meta fields, read-ahead that keeps order and propagates decode errors."""
    import cv2

    from ocr_inference_b200.dataio import ImageInferenceDataset, PageLoader

    rng = np.random.default_rng(1)
    imgs = {}
    for name, shape in (("b.png", (40, 60, 3)), ("a.jpg", (50, 30, 3)), ("c.jpeg", (20, 20, 3)), ("g.png", (16, 24))):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        cv2.imwrite(str(tmp_path / name), img)
        imgs[name] = img
    (tmp_path / "notes.txt").write_text("not an image")
    ds = ImageInferenceDataset(str(tmp_path))
    assert [os.path.basename(p) for p in ds.paths] == ["a.jpg", "b.png", "c.jpeg", "g.png"] and len(ds) == 4
    t, meta = ds[1]
    assert meta == {"image_name": "b.png", "orig_shape": (40, 60)} and t.dtype == torch.uint8 and tuple(t.shape) == (3, 40, 60)
    assert np.array_equal(t.numpy().transpose(1, 2, 0), imgs["b.png"][..., ::-1])        # PNG is lossless; BGR -> RGB
    tg, _ = ds.get_item(3)
    assert tuple(tg.shape) == (3, 16, 24)                                                  # gray file, mode "rgb"
    tg1, _ = ImageInferenceDataset(str(tmp_path), mode="unchanged")[3]
    assert tuple(tg1.shape) == (1, 16, 24) and np.array_equal(tg1[0].numpy(), imgs["g.png"])
    seen = [(m["image_name"], p.shape) for p, m in PageLoader(str(tmp_path), workers=3, prefetch=2)]
    assert seen == [("a.jpg", (50, 30, 3)), ("b.png", (40, 60, 3)), ("c.jpeg", (20, 20, 3)), ("g.png", (16, 24, 3))]
    assert [len(b[0]) for b in PageLoader(str(tmp_path)).batches(3)] == [3, 1]
    (tmp_path / "d.png").write_bytes(b"broken")
    with pytest.raises(OSError):
        list(PageLoader(str(tmp_path)))


def test_bench_deadline_fires_and_cancels():
    """bench.py arms a deadline around the multi-rank C5 block: past it the callback runs (rank 0 prints the
    headline line) and the process leaves with exit code 0; cancel() disarms it; seconds=None arms nothing."""
    import importlib.util
    import os
    import threading

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    fired, left = [], threading.Event()
    bench.Deadline(0.05, lambda: fired.append("line"), _exit=lambda code: (fired.append(code), left.set()))
    assert left.wait(5.0) and fired == ["line", 0]
    fired.clear()
    boom = bench.Deadline(0.05, lambda: 1 / 0, _exit=lambda code: (fired.append(code), left.set()))
    left.clear()
    assert left.wait(5.0) and fired == [0]                 # a failing callback still ends the process with 0
    del boom
    fired.clear()
    d = bench.Deadline(0.2, lambda: fired.append("late"), _exit=lambda code: fired.append(code))
    d.cancel()
    time_ = __import__("time")
    time_.sleep(0.4)
    assert fired == []
    bench.Deadline(None).cancel()


def test_geometry_workers_leave_when_their_owner_is_gone():
    """A GeometryPool worker must not outlive an owner that went away without close() (os._exit, SIGKILL): an
    orphan would keep the job's stdout / stderr pipes open."""
    import subprocess
    import sys
    import time

    code = ("import os, sys\n"
            "from ocr_inference_b200.geometry_pool import GeometryPool\n"
            "if __name__ == '__main__':\n"
            "    p = GeometryPool(1)\n"
            "    print(p._procs[0].pid, flush=True)\n"
            "    os._exit(0)\n")
    root = os.path.join(os.path.dirname(__file__), "..")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    pid = int(r.stdout.strip().splitlines()[-1])

    def alive():
        try:
            with open(f"/proc/{pid}/status") as f:
                return "State:\tZ" not in f.read()
        except OSError:
            return False

    t0 = time.time()
    while alive() and time.time() - t0 < 30:
        time.sleep(0.5)
    assert not alive(), "the orphaned geometry worker is still running"


def _check_bench_line(d, reference: bool):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "pages_per_sec" and d["unit"] == "pages/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    if reference:
        assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["value"] == d["value"]
    else:
        r = d["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, k
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert d["gpu_launches"] > 0 and d["simt_conv_launches"] == 0
        assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
        assert d["e2e"]["h2d_bytes_per_step"] == 16 * 2048 * 3072 * 3 and d["e2e"]["d2h_bytes_per_step"] == 16 * 2048 * 3072
        for k in ("a-1", "a-3", "a-5", "a-7", "a-9"):
            assert 0 < d["roofline_stages"][k]["frac"] < 1, k
        assert d["c1"]["roofline"]["bound"] == "tensor" and d["c3"]["cpu_baseline"]["kind"] == "port"
        assert d["c5"]["collectives_on_data_path"] == 0 and d["c5"]["scaling"] == "weak"


def test_bench_lines_keep_the_contract():
    """The JSON line of `bench.py --impl reference` (run here: the oracle on the host cores) and the committed GPU
    lines of the final build carry every key the bench contract names."""
    import json
    import subprocess
    import sys

    root = os.path.join(os.path.dirname(__file__), "..")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                    # ONE JSON line on stdout
    _check_bench_line(json.loads(lines[0]), reference=True)
    with open(os.path.join(root, "profiles", "r02e_bench.json")) as f:
        _check_bench_line(json.loads(f.read()), reference=False)
    with open(os.path.join(root, "profiles", "r02e_bench_2gpu.json")) as f:
        d2 = json.loads(f.read())
    assert d2["n_gpus"] == 2 and d2["c5"]["n_gpus"] == 2 and d2["c5"]["pages_ok"] == d2["c5"]["pages_per_step"]
    with open(os.path.join(root, "profiles", "r02e_bench_2gpu_c5_fault.json")) as f:
        lines = [ln for ln in f.read().splitlines() if ln.strip()]
    assert len(lines) == 1 and "error" in json.loads(lines[0])["c5"]
