"""CPU: the C-ABI library loads, exports every symbol include/zoomvit.h declares, its host functions are bit-exact
against the oracle, and its device entry points fail loudly (no fallback) when there is no GPU."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import geometry as OG, processor as OP, resample as OR, tower as OT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol(lib):
    from zoomearth_b200 import _lib
    header = open(os.path.join(ROOT, "include", "zoomvit.h")).read()
    declared = set(re.findall(r"ZV_API\s+[\w\s\*]+?\b(zv_\w+)\s*\(", header))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"libzoomvit.so does not export {name}"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert b"sm_100a" in lib.zv_version()


def test_cfg_struct_layout_matches_defaults(lib):
    from zoomearth_b200 import _lib
    cfg = _lib.default_cfg()
    assert (cfg.patch, cfg.merge, cfg.temporal, cfg.window, cfg.min_size) == (14, 2, 2, 112, 512)
    assert (cfg.depth, cfg.hidden, cfg.heads, cfg.inter, cfg.out_hidden) == (32, 1280, 16, 3420, 2048)
    assert cfg.min_pixels == 3136 and cfg.max_pixels == 1003520 and abs(cfg.rescale - 1 / 255) < 1e-18
    assert [(cfg.fullatt_mask_lo >> i) & 1 for i in (7, 15, 23, 31)] == [1, 1, 1, 1]
    assert abs(cfg.eps - 1e-6) < 1e-12 and abs(cfg.mean[0] - 0.48145466) < 1e-7


def test_geometry_vs_golden(lib):
    from zoomearth_b200 import geometry as G
    geo = json.load(open(os.path.join(GOLD, "geometry.json")))
    for c in geo["cut_image"]:
        assert list(G.cut_box(c["w"], c["h"], c["bbox"])) == c["box"], c
    for c in geo["resize_image"]:
        assert list(G.resize_dims(c["w"], c["h"], c["max_size"])[:2]) == c["size"], c
    for c in geo["smart_resize"]:
        if c["out"] == "ValueError":
            with pytest.raises(ValueError, match="absolute aspect ratio must be smaller than 200"):
                G.smart_resize(c["h"], c["w"], 28, 3136, c["max_pixels"])
        else:
            assert list(G.smart_resize(c["h"], c["w"], 28, 3136, c["max_pixels"])) == c["out"], c


@settings(max_examples=300, deadline=None)
@given(st.integers(64, 8000), st.integers(64, 8000), st.floats(-500, 8000), st.floats(-500, 8000),
       st.floats(0, 3000), st.floats(0, 3000))
def test_cut_box_property(w, h, x1, y1, dw, dh):
    from zoomearth_b200 import geometry as G
    b = (x1, y1, x1 + dw, y1 + dh)
    assert G.cut_box(w, h, b) == OG.cut_box(w, h, b)


@settings(max_examples=400, deadline=None)
@given(st.integers(1, 9000), st.integers(1, 9000), st.sampled_from([3136, 50176, 200704, 1003520, 12845056]))
def test_smart_resize_property(h, w, mx):
    from zoomearth_b200 import geometry as G
    try:
        ref = OG.smart_resize(h, w, 28, 3136, mx)
    except ValueError:
        with pytest.raises(ValueError):
            G.smart_resize(h, w, 28, 3136, mx)
        return
    assert G.smart_resize(h, w, 28, 3136, mx) == ref


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 3000), st.integers(1, 1200))
def test_resample_coeffs_property(in_size, out_size):
    from zoomearth_b200 import geometry as G
    ks, b, k = G.resample_coeffs(in_size, out_size)
    if in_size == out_size:                       # Pillow copies; the library states it as one exact tap
        assert ks == 1 and np.array_equal(b[:, 0], np.arange(out_size)) and (k == 1 << 22).all()
        return
    ks2, b2, k2 = OR.precompute_coeffs(in_size, 0, in_size, out_size)
    assert ks == ks2 and np.array_equal(b, b2) and np.array_equal(k, k2)


def test_normalize_lut_bitexact(lib):
    from zoomearth_b200 import _lib, geometry as G
    assert np.array_equal(G.normalize_lut(_lib.default_cfg()).view(np.uint32), OP.normalize_lut().view(np.uint32))


@pytest.mark.parametrize("grid", [[[1, 2, 2]], [[1, 8, 8]], [[1, 26, 36]], [[1, 70, 70], [1, 64, 92], [1, 4, 200]],
                                  [[2, 6, 10], [1, 254, 254]]])
def test_plan_vs_oracle(lib, grid):
    from zoomearth_b200 import Plan, _lib
    grid = np.array(grid)
    p = Plan(_lib.default_cfg(), grid)
    w, cu = OT.window_index(grid)
    assert np.array_equal(p.window_index, w)
    assert p.cu_window_seqlens_raw.tolist() == cu
    assert p.cu_window_seqlens.tolist() == OT.unique_consecutive(cu)
    assert p.cu_seqlens.tolist() == OT.cu_seqlens_full(grid)
    assert np.array_equal(p.pos_ids, OT.rot_pos_ids(grid))
    assert np.array_equal(p.reverse_index, np.argsort(w))
    assert p.num_patches == int((grid[:, 0] * grid[:, 1] * grid[:, 2]).sum()) and p.num_tokens * 4 == p.num_patches


def test_plan_vs_golden_hf(lib):
    from zoomearth_b200 import Plan, _lib
    z = np.load(os.path.join(GOLD, "tower.npz"))
    p = Plan(_lib.default_cfg(), z["grid"])
    assert np.array_equal(p.window_index, z["window_index"])
    assert p.cu_window_seqlens_raw.tolist() == z["cu_window_seqlens"].tolist()


def test_window_pos_closed_form_matches_argsort():
    """The closed form K1 uses to write patches in window order equals argsort(window_index)."""
    def window_pos(my, mx, lh, lw, ws=4):
        wy, wx = my // ws, mx // ws
        bh, bw = min(ws, lh - wy * ws), min(ws, lw - wx * ws)
        return wy * ws * lw + wx * ws * bh + (my - wy * ws) * bw + (mx - wx * ws)
    for gh, gw in [(2, 2), (8, 8), (26, 36), (70, 70), (10, 14), (4, 200), (18, 34)]:
        w, _ = OT.window_index(np.array([[1, gh, gw]]))
        rev = np.argsort(w)
        lh, lw = gh // 2, gw // 2
        got = np.array([window_pos(m // lw, m % lw, lh, lw) for m in range(lh * lw)])
        assert np.array_equal(got, rev), (gh, gw)


def test_plan_rejects_bad_grid(lib):
    from zoomearth_b200 import Plan, _lib
    with pytest.raises(_lib.ZoomVitError):
        Plan(_lib.default_cfg(), np.array([[1, 7, 8]]))


def test_geometry_errors(lib):
    from zoomearth_b200 import _lib, geometry as G
    cfg = _lib.default_cfg()
    with pytest.raises(ValueError, match="absolute aspect ratio must be smaller than 200"):
        G.geometry(cfg, [[10, 4000]], None)
    cfg.min_size = -1
    with pytest.raises(ValueError, match="Coordinate 'right' is less than 'left'"):
        G.geometry(cfg, [[100, 100]], [[50, 0, 40, 10]])
    with pytest.raises(ValueError, match="Coordinate 'lower' is less than 'upper'"):
        G.geometry(cfg, [[100, 100]], [[0, 50, 40, 10]])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_device_entry_points_fail_loudly_without_gpu(lib):
    """No CPU fallback: without a device every compute entry point returns ZV_ENODEV with a message."""
    from zoomearth_b200 import _lib
    cfg = _lib.default_cfg()
    crop = np.array([[0, 0, 56, 56]], np.int32)
    rhw = np.array([[56, 56]], np.int32)
    hw = np.array([[56, 56]], np.int32)
    pitch = np.array([168], np.int64)
    ptrs = (C.c_void_p * 1)(0x1000)
    rc = lib.zv_preprocess(C.byref(cfg), 1, ptrs, hw.ctypes.data, pitch.ctypes.data, crop.ctypes.data, rhw.ctypes.data,
                           None, 0x1000, 0, 0, 0x1000, 1 << 30, None)
    assert rc == _lib.ZV_ENODEV and b"no CUDA device" in lib.zv_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        from zoomearth_b200 import FusedVisual
        FusedVisual({})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        from zoomearth_b200 import FusedImageProcessor
        FusedImageProcessor()(images=[np.zeros((56, 56, 3), np.uint8)])


def test_missing_library_is_an_import_error(monkeypatch, lib):
    from zoomearth_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libzoomvit.so")
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_processor_host_surface(lib):
    from zoomearth_b200 import FusedImageProcessor
    fp = FusedImageProcessor(min_pixels=3136, max_pixels=128 * 128 * 28 * 28)
    assert fp.merge_size == 2 and fp.patch_size == 14 and fp.model_input_names == ["pixel_values", "image_grid_thw"]
    assert fp.get_number_of_image_patches(512, 512) == 36 * 36
    assert fp.get_number_of_image_patches(5000, 5000, {"max_pixels": 1003520}) == 4900
    fp.max_pixels = 1003520
    assert fp.size["longest_edge"] == 1003520 and fp.min_pixels == 3136


def test_micro_batches_split_a_ragged_crop_list_by_patch_budget():
    """ZoomEncoder.micro_batches: host-only planning (geometry through the C ABI), no GPU involved."""
    import types
    import numpy as np
    import torch
    from zoomearth_b200.processor import FusedImageProcessor
    from zoomearth_b200.zoom import ZoomEncoder
    enc = ZoomEncoder.__new__(ZoomEncoder)
    enc.processor = FusedImageProcessor(min_pixels=3136, max_pixels=12845056)
    enc.visual = types.SimpleNamespace()
    imgs = [torch.empty((5000, 5000, 3), dtype=torch.uint8, device="meta")]
    boxes = [(0, 0, 512, 512), (0, 0, 2048, 2048), (100, 100, 400, 300), (0, 0, 5000, 5000), (10, 10, 1034, 1034)]
    patches = [1296, 21316, 1296, 64516, 5476]                    # SURVEY 8 table; the 300x200 box becomes 512x512
    groups = enc.micro_batches(imgs, boxes, [0] * 5, max_patches=25000)
    assert groups == [[0, 1, 2], [3], [4]]                        # consecutive, in order; an oversize crop stands alone
    assert [sum(patches[i] for i in g) for g in groups] == [23908, 64516, 5476]
    assert enc.micro_batches(imgs, boxes, [0] * 5, max_patches=10 ** 9) == [[0, 1, 2, 3, 4]]
    assert enc.micro_batches(imgs, None) == [[0]]                  # global view of every image
