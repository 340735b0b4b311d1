"""CPU check of K1's tensor-core route (csrc/zv_k1_tc.cuh) without a GPU: ``zv_debug_k1_tc_host`` runs the real host
code (job descriptors, tap tables, work lists) over host memory and emulates the tcgen05 kernels lane by lane; the
result must equal the oracle (Pillow-exact resample + HF processor) bit for bit.  The GPU tests then only have to prove
the hardware layouts."""
import ctypes as C

import numpy as np
import pytest

from oracle import processor as OP, resample as OR
from zoomearth_b200 import _lib


def _img(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def _run(imgs, index, boxes, out_hw, patches, row_order=_lib.ORDER_HF):
    """-> (result, took) through the debug hook; patches: fp32 patch rows, else a list of uint8 images."""
    lib = _lib.lib()
    n = len(boxes)
    srcs = (C.c_void_p * n)(*[imgs[k].ctypes.data for k in index])
    src_hw = np.array([[imgs[k].shape[0], imgs[k].shape[1]] for k in index], np.int32)
    pitch = np.array([imgs[k].strides[0] for k in index], np.int64)
    box = np.array(boxes, np.int32)
    hw = np.array(out_hw, np.int32)
    took = np.zeros(n, np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    if patches:
        need = _lib.check(lib.zv_preprocess_workspace_bytes(n, P(box), P(hw)))
        rows = sum((h // 14) * (w // 14) for h, w in out_hw)
        out = np.zeros((rows, 1176), np.float32)
        ws = np.zeros(need + 512, np.uint8)
        base = (ws.ctypes.data + 255) & ~255
        cfg = _lib.default_cfg()
        _lib.check(lib.zv_debug_k1_tc_host(C.byref(cfg), n, srcs, P(src_hw), P(pitch), P(box), P(hw), P(out), row_order, None, None,
                                           C.c_void_p(base), need, P(took)))
        return out, took
    need = _lib.check(lib.zv_resize_u8_workspace_bytes(n, P(box), P(hw)))
    outs = [np.zeros((h, w, 3), np.uint8) for h, w in out_hw]
    dsts = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    dpitch = np.array([o.strides[0] for o in outs], np.int64)
    ws = np.zeros(need + 512, np.uint8)
    base = (ws.ctypes.data + 255) & ~255
    _lib.check(lib.zv_debug_k1_tc_host(None, n, srcs, P(src_hw), P(pitch), P(box), P(hw), None, 0, dsts, P(dpitch),
                                       C.c_void_p(base), need, P(took)))
    return outs, took


def _aligned(a):
    """copy into 16-byte aligned storage (the route needs a 16-byte aligned image base)"""
    buf = np.zeros(a.nbytes + 16, np.uint8)
    off = (-buf.ctypes.data) % 16
    v = buf[off:off + a.nbytes].view(np.uint8).reshape(a.shape)
    v[...] = a
    return v


CASES = [
    # (h, w, box, out_hw, note)
    (600, 800, (0, 0, 800, 600), (420, 560), "mild downscale"),
    (512, 512, (0, 0, 512, 512), (504, 504), "typical zoom crop"),
    (504, 504, (0, 0, 504, 504), (504, 504), "identity on both axes"),
    (64, 64, (8, 8, 40, 40), (56, 56), "upscale"),
    (900, 1300, (100, 200, 700, 648), (308, 420), "crop + 1.4x downscale, odd offsets"),
    (1200, 1600, (3, 5, 1599, 1197), (224, 308), "5.2x downscale, 23 taps"),
    (504, 700, (0, 0, 700, 504), (504, 728), "vertical identity"),
    (2000, 2000, (0, 0, 2000, 2000), (196, 196), "10x downscale, 4 K blocks"),
    (400, 1000, (1, 2, 999, 398), (140, 196), "pitch = 8 mod 16: two B variants (the 5000-pixel case)"),
    (800, 1300, (0, 0, 1300, 800), (168, 252), "pitch = 12 mod 16, two K blocks: four B variants, single-buffered B"),
]


@pytest.mark.parametrize("h,w,box,ohw,note", CASES)
def test_tc_route_patches_bitexact_vs_oracle(h, w, box, ohw, note):
    img = _aligned(_img(h * 3 + w, h, w))
    got, took = _run([img], [0], [box], [ohw], patches=True)
    assert took.tolist() == [1], note
    crop = OR.crop_u8(img, box)
    r = OR.resize_u8(crop, ohw[1], ohw[0])
    lut = OP.normalize_lut()
    ref, _ = OP.patchify(np.stack([lut[c][r[:, :, c]] for c in range(3)], 0))
    assert np.array_equal(got, ref), note


def test_tc_route_resize_u8_ragged_batch():
    imgs = [_aligned(_img(1, 700, 900)), _aligned(_img(2, 1000, 640))]
    boxes = [(0, 0, 900, 700), (10, 20, 410, 620), (100, 100, 640, 1000), (0, 0, 640, 996)]
    index = [0, 0, 1, 1]
    out_hw = [(350, 450), (301, 199), (512, 307), (249, 160)]
    outs, took = _run(imgs, index, boxes, out_hw, patches=False)
    assert took.tolist() == [1, 1, 1, 1]
    for o, b, k, (oh, ow) in zip(outs, boxes, index, out_hw):
        assert np.array_equal(o, OR.resize_u8(OR.crop_u8(imgs[k], b), ow, oh))


def test_tc_route_declines_what_it_cannot_do():
    img = _aligned(_img(3, 602, 803))[:, :801]                               # row pitch 2409: not a multiple of 4
    _, took = _run([img], [0], [(0, 0, 800, 600)], [(280, 392)], patches=True)
    assert took.tolist() == [0]
    img = _aligned(_img(3, 600, 4800))                                        # 12x horizontal downscale: more than four K blocks
    _, took = _run([img], [0], [(0, 0, 4800, 600)], [(280, 392)], patches=True)
    assert took.tolist() == [0]


@pytest.mark.parametrize("box", [(-20, -30, 480, 470), (300, 200, 900, 640), (-7, 100, 810, 400), (100, -50, 500, 700),
                                 (790, 590, 830, 630), (-60, -60, -10, -10), (0, 598, 800, 640)])
def test_tc_route_boxes_that_leave_the_image(box):
    """Image.crop's zero fill: rows above / below the image are zero-filled by the TMA unit, taps left / right of it are
    left out of B; a box wholly outside gives the resize of a black image"""
    img = _aligned(_img(5, 602, 800))
    w, h = box[2] - box[0], box[3] - box[1]
    ohw = (max(8, int(h / 1.7)), max(8, int(w / 1.3)))
    outs, took = _run([img], [0], [box], [ohw], patches=False)
    assert took.tolist() == [1]
    assert np.array_equal(outs[0], OR.resize_u8(OR.crop_u8(img, box), ohw[1], ohw[0])), box


@pytest.mark.parametrize("h", [601, 602, 603, 1203])
def test_tc_route_image_height_not_a_multiple_of_four(h):
    """the last h mod 4 rows come through a second job over an end-aligned tensor map (tc_split_rows): full-height crops,
    crops that end on the last row, and a crop inside the last few rows all stay on the route and match the oracle"""
    img = _aligned(_img(h, h, 640))
    boxes = [(0, 0, 640, h), (16, h - 300, 600, h), (0, h - 40, 640, h), (5, 3, 500, h - 5)]
    out_hw = [(h // 2, 320), (150, 292), (40, 640), (200, 180)]
    outs, took = _run([img], [0] * 4, boxes, out_hw, patches=False)
    assert took.tolist() == [1, 1, 1, 1]
    for o, b, (oh, ow) in zip(outs, boxes, out_hw):
        assert np.array_equal(o, OR.resize_u8(OR.crop_u8(img, b), ow, oh)), (h, b)


def test_tc_route_takes_unaligned_views():
    """a cropped view (base neither 16- nor 4-byte aligned, pitch of the parent image) stays on the route"""
    big = _aligned(_img(4, 1000, 1200))
    view = big[3:803, 5:905]
    assert view.ctypes.data % 16 == 15
    outs, took = _run([view], [0, 0], [(0, 0, 900, 800), (7, 9, 507, 409)], [(400, 452), (250, 313)], patches=False)
    assert took.tolist() == [1, 1]
    assert np.array_equal(outs[0], OR.resize_u8(np.ascontiguousarray(view), 452, 400))
    assert np.array_equal(outs[1], OR.resize_u8(OR.crop_u8(np.ascontiguousarray(view), (7, 9, 507, 409)), 313, 250))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_tc_route_random_geometries(seed):
    """random image sizes / pitches / boxes / scales (4x upscale .. 9x downscale), uint8 output of any size: whatever the
    route accepts must be bit-identical to the oracle; what it declines is a crop it cannot express (border, odd pitch)"""
    rng = np.random.default_rng(100 + seed)
    took_any = 0
    for _ in range(10):
        h, w = int(rng.integers(40, 420)), int(rng.integers(10, 90)) * 4
        img = _aligned(_img(int(rng.integers(1 << 30)), h, w))
        boxes, out_hw = [], []
        for _ in range(3):
            bw, bh = int(rng.integers(12, w + 1)), int(rng.integers(12, h + 1))
            x0, y0 = int(rng.integers(0, w - bw + 1)), int(rng.integers(0, h - bh + 1))
            boxes.append((x0, y0, x0 + bw, y0 + bh))
            s = float(rng.uniform(0.25, 9.0))
            out_hw.append((max(3, int(bh / s)), max(3, int(bw / s))))
        outs, took = _run([img], [0, 0, 0], boxes, out_hw, patches=False)
        for o, b, t, (oh, ow) in zip(outs, boxes, took, out_hw):
            if t:
                took_any += 1
                assert np.array_equal(o, OR.resize_u8(OR.crop_u8(img, b), ow, oh)), (h, w, b, oh, ow)
            else:
                assert max((b[2] - b[0]) / ow, (b[3] - b[1]) / oh) > 4.0, (h, w, b, oh, ow)   # only the K-block budget declines
    assert took_any >= 15


def test_tc_route_window_order_rows():
    """row_order = window: merge groups land at the tower's window position (closed form of argsort(window_index))"""
    from oracle import tower as OT
    img = _aligned(_img(9, 700, 1000))
    box, ohw = (20, 40, 980, 680), (392, 588)                  # grid 28 x 42: 3.5 x 5.25 windows of 4 x 4 merge groups
    got, took = _run([img], [0], [box], [ohw], patches=True, row_order=_lib.ORDER_WINDOW)
    assert took.tolist() == [1]
    r = OR.resize_u8(OR.crop_u8(img, box), ohw[1], ohw[0])
    lut = OP.normalize_lut()
    ref, grid = OP.patchify(np.stack([lut[c][r[:, :, c]] for c in range(3)], 0))
    widx, _ = OT.window_index(np.asarray(grid).reshape(1, 3))
    ref_w = ref.reshape(-1, 4, 1176)[np.asarray(widx)].reshape(-1, 1176)
    assert np.array_equal(got, ref_w)
