"""CPU: the oracle is pinned against (a) the installed Pillow / transformers run live, (b) the golden fixtures
generated from the reference's own code (tests/golden/make_golden.py), (c) the known answers of SURVEY.md 8c."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import geometry as OG, processor as OP, resample as OR, tower as OT

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def geo():
    return json.load(open(os.path.join(GOLD, "geometry.json")))


def test_cut_box_vs_reference_cut_image(geo):
    for c in geo["cut_image"]:
        assert list(OG.cut_box(c["w"], c["h"], c["bbox"])) == c["box"], c


def test_resize_dims_vs_reference_resize_image(geo):
    for c in geo["resize_image"]:
        w, h, _ = OG.resize_dims(c["w"], c["h"], c["max_size"])
        assert [w, h] == c["size"], c


def test_smart_resize_vs_hf(geo):
    for c in geo["smart_resize"]:
        if c["out"] == "ValueError":
            with pytest.raises(ValueError, match="absolute aspect ratio must be smaller than 200"):
                OG.smart_resize(c["h"], c["w"], 28, 3136, c["max_pixels"])
        else:
            assert list(OG.smart_resize(c["h"], c["w"], 28, 3136, c["max_pixels"])) == c["out"], c


def test_extract_bbox_vs_reference(geo):
    for c in geo["extract_bbox"]:
        # demo.py parses ints only; infer.py parses floats - both are offered
        from zoomearth_b200 import extract_bbox
        assert extract_bbox(c["text"], c["scale"], integer_only=True) == c["out_int"], c
        for b in c["out_int"]:                       # every int-parsable box is also found by the float parser
            assert b in OG.extract_bbox(c["text"], c["scale"])
            assert b in extract_bbox(c["text"], c["scale"])
    assert OG.extract_bbox('"bbox_2d" : [1.5,2.5, 3.5 ,4.5]', 2.0) == [[3.0, 5.0, 7.0, 9.0]]


def test_survey_known_answers():
    img = np.random.default_rng(0).integers(0, 256, (5000, 5000, 3), dtype=np.uint8)
    assert hashlib.sha256(img.tobytes()).hexdigest()[:16] == "7e792df7cf394adb"
    box, pv, grid = OP.zoom_step_u8(img, (1000, 1200, 2300, 2100), max_pixels=1003520)
    assert box == (1000, 1200, 2300, 2100) and grid.tolist() == [[1, 58, 84]] and pv.shape == (4872, 1176)
    assert hashlib.sha256(pv.tobytes()).hexdigest()[:16] == "fded43574812b143"
    assert abs(float(pv.sum(dtype=np.float64)) - 1070490.5948) < 1e-3
    np.testing.assert_allclose(pv[0, :3], [-1.00394750, 0.30991095, 0.10553299], rtol=0, atol=1e-7)
    _, pv2, grid2 = OP.zoom_step_u8(img, (2000, 2000, 2512, 2512), max_pixels=12845056)
    assert grid2.tolist() == [[1, 36, 36]] and hashlib.sha256(pv2.tobytes()).hexdigest()[:16] == "0f13345ba19e8f91"
    w, cu = OT.window_index(np.array([[1, 26, 36]]))
    assert w[:20].tolist() == [0, 1, 2, 3, 18, 19, 20, 21, 36, 37, 38, 39, 54, 55, 56, 57, 4, 5, 6, 7]
    assert OT.unique_consecutive(cu)[:8] == [0, 64, 128, 192, 256, 288, 352, 416]


def test_pixels_golden():
    z = np.load(os.path.join(GOLD, "pixels.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    img = np.random.default_rng(101).integers(0, 256, (420, 560, 3), dtype=np.uint8)
    for i, m in enumerate(meta[:-1]):
        crop = OR.crop_u8(img, m["box"])
        pv, grid, resized = OP.preprocess_u8([crop], m["min_pixels"], m["max_pixels"])
        assert grid[0].tolist() == m["grid"]
        assert hashlib.sha256(resized[0].tobytes()).hexdigest() == m["resized_sha256"]
        if f"resized_{i}" in z:
            assert np.array_equal(resized[0], z[f"resized_{i}"])
        assert hashlib.sha256(pv.tobytes()).hexdigest() == m["pv_sha256"], m
    # the reference-faithful two-resample flow of demo.py: cut_image -> resize_image(1024) -> processor
    m = meta[-1]
    big = np.random.default_rng(102).integers(0, 256, (1300, 1700, 3), dtype=np.uint8)
    crop = OR.crop_u8(big, OG.cut_box(1700, 1300, m["bbox"]))
    w, h, _ = OG.resize_dims(crop.shape[1], crop.shape[0], 1024)
    two = OR.resize_u8(crop, w, h)
    assert [w, h] == m["two_step_size"] and hashlib.sha256(two.tobytes()).hexdigest() == m["two_step_sha256"]
    pv, grid, _ = OP.preprocess_u8([two], 3136, 200704)
    assert grid[0].tolist() == m["grid"] and hashlib.sha256(pv.tobytes()).hexdigest() == m["pv_sha256"]


def test_tower_golden():
    z = np.load(os.path.join(GOLD, "tower.npz"))
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(5, cfg)
    out = OT.forward(sd, torch.from_numpy(z["pixel_values"]), z["grid"], cfg)
    assert torch.allclose(out, torch.from_numpy(z["out"]), rtol=0, atol=2e-4), (out - torch.from_numpy(z["out"])).abs().max()
    w, cu = OT.window_index(z["grid"])
    assert np.array_equal(w, z["window_index"]) and cu == z["cu_window_seqlens"].tolist()


def test_resample_vs_live_pillow():
    from PIL import Image
    img = np.random.default_rng(7).integers(0, 256, (700, 900, 3), dtype=np.uint8)
    for box, ow, oh in [((0, 0, 900, 700), 448, 336), ((10, 20, 110, 90), 300, 200), ((0, 0, 512, 512), 504, 504),
                        ((0, 0, 600, 600), 600, 300), ((-30, -20, 200, 150), 140, 112), ((5, 5, 33, 33), 56, 56)]:
        a = OR.resize_u8(OR.crop_u8(img, box), ow, oh)
        b = np.asarray(Image.fromarray(img).crop(box).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(a, b), (box, ow, oh)
    with pytest.raises(ValueError, match="Coordinate 'right' is less than 'left'"):
        OR.crop_u8(img, (50, 0, 40, 10))


def test_processor_vs_live_hf():
    from oracle import hf_live
    imgs = [np.random.default_rng(s).integers(0, 256, hw + (3,), dtype=np.uint8) for s, hw in [(1, (300, 420)), (2, (60, 44))]]
    pv, grid, _ = OP.preprocess_u8(imgs, 3136, 100352)
    hpv, hgrid = hf_live.hf_preprocess(imgs, 3136, 100352)
    assert hgrid.tolist() == grid.tolist() and np.array_equal(hpv.numpy(), pv)


def test_tower_vs_live_hf():
    from oracle import hf_live
    cfg = OT.small_cfg(depth=2, fullatt=(0,))
    sd = OT.make_weights(9, cfg)
    grid = np.array([[1, 10, 6], [1, 4, 4]])
    pv = torch.randn(int((grid[:, 1] * grid[:, 2]).sum()), 1176, generator=torch.Generator().manual_seed(1))
    ref = hf_live.hf_tower_forward(hf_live.hf_tower(sd, cfg), pv, torch.from_numpy(grid))
    out = OT.forward(sd, pv, grid, cfg)
    assert torch.allclose(out, ref, rtol=0, atol=2e-4)
