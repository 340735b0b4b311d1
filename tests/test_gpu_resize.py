"""zv_resize_u8 and the reference-faithful two-resample flow on the GPU (through the C ABI): bit-exact against live
Pillow, against tests/golden/flows.json (the reference's own cut_image / resize_image, all four call sites) and against
the demo.py two-step fixture of tests/golden/pixels.npz."""
import hashlib
import json
import os
import time

import numpy as np
import pytest
import torch

from oracle import flows as OF, geometry as OG, processor as OP, tower as OT

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _sha(t):
    a = t.cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def proc(cuda):
    from zoomearth_b200 import FusedImageProcessor
    return FusedImageProcessor(min_pixels=3136, max_pixels=12845056, device=cuda)


def test_resize_u8_bitexact_vs_live_pillow(cuda, proc):
    """Crop boxes inside / outside the image, down- and up-scales, same-size axes (Pillow copies), odd widths, every tap
    class of the dp4a path (5 .. 45 taps) and the per-tap path (> 45 taps)."""
    from PIL import Image
    rng = np.random.default_rng(21)
    img = rng.integers(0, 256, (1237, 1803, 3), dtype=np.uint8)
    dev = torch.from_numpy(img).to(cuda)
    pil = Image.fromarray(img)
    cases = [((0, 0, 1803, 1237), (512, 351)), ((10, 20, 1310, 920), (512, 354)), ((100, 100, 612, 612), (504, 504)),
             ((0, 0, 600, 600), (600, 300)), ((0, 0, 600, 600), (300, 600)), ((-30, -20, 200, 150), (140, 112)),
             ((5, 5, 33, 33), (56, 56)), ((1700, 1100, 1900, 1300), (333, 77)), ((0, 0, 1803, 1237), (1803, 1237)),
             ((3, 7, 1000, 1200), (997, 1193)), ((0, 0, 1803, 1237), (41, 29)), ((0, 0, 1803, 1237), (17, 1237)),
             ((200, 300, 1229, 811), (1024, 509)), ((0, 0, 1800, 1236), (2048, 1406))]
    outs = proc.resize_u8([dev], [b for b, _ in cases], [s for _, s in cases], image_index=[0] * len(cases))
    for (box, size), o in zip(cases, outs):
        ref = np.asarray(pil.crop(box).resize(size, Image.BICUBIC))
        assert o.shape == ref.shape and np.array_equal(o.cpu().numpy(), ref), (box, size)


def test_resize_u8_5000px_to_512_and_1024_beside_pillow(cuda, proc):
    """The dominant CPU cost of the reference loop (infer.py:215: resize_image(Image.open(fp)), 5000^2 -> 512^2, 41 taps;
    demo.py:133: -> 1024^2) on the device: bit-exact, timed beside Pillow on this box's host."""
    from PIL import Image
    img = np.random.default_rng(0).integers(0, 256, (5000, 5000, 3), dtype=np.uint8)
    dev = torch.from_numpy(img).to(cuda)
    pil = Image.fromarray(img)
    for ms in (512, 1024):
        t0 = time.perf_counter()
        ref = np.asarray(pil.resize((ms, ms), Image.BICUBIC))
        t_pil = time.perf_counter() - t0
        outs, inv = proc.cut_resize([dev], None, variant="infer", max_size=ms)
        assert np.array_equal(outs[0].cpu().numpy(), ref) and inv[0] == 5000 / ms
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            proc.cut_resize([dev], None, variant="infer", max_size=ms)
        e1.record()
        torch.cuda.synchronize()
        ms_gpu = e0.elapsed_time(e1) / 10
        print(f"PARITY resize_image 5000x5000 -> {ms}x{ms}: bit-exact vs Pillow; Pillow {t_pil * 1e3:.1f} ms (1 thread), "
              f"zv_resize_u8 {ms_gpu:.3f} ms ({75e6 / ms_gpu / 1e6:.0f} GB/s of source pixels)")


def test_cut_resize_variants_vs_reference_fixtures(cuda, proc):
    """Every pixel case of tests/golden/flows.json - the reference's own cut_image / resize_image from infer.py, SFT.py
    and customized_funcs.py run with real Pillow - reproduced on the device, compared by sha256 of the uint8 result."""
    flows = json.load(open(os.path.join(GOLD, "flows.json")))
    img = np.random.default_rng(303).integers(0, 256, (1500, 2100, 3), dtype=np.uint8)
    dev = torch.from_numpy(img).to(cuda)
    for c in flows["pixels"]:
        if c["flow"] == "resize_image":
            outs, _ = proc.cut_resize([dev], None, variant=c["variant"], max_size=c["max_size"])
        elif c["flow"] == "cut_image":
            outs, _ = proc.cut_resize([dev], [c["bbox"]], image_index=[0], variant=c["variant"], max_size=0)
        else:
            outs, _ = proc.cut_resize([dev], [c["bbox"]], image_index=[0], variant=c["variant"], max_size=c["max_size"])
        o = outs[0]
        assert [o.shape[1], o.shape[0]] == c["size"], c
        assert _sha(o.contiguous()) == c["sha256"], c
    outs, _ = proc.cut_resize([dev], [[1, 2, 3]], image_index=[0], variant="custom", max_size=0)   # len(bbox) != 4
    assert outs[0].data_ptr() == dev.data_ptr()


def test_two_resample_flow_matches_demo_fixture_and_oracle(cuda, proc):
    """cut_image -> resize_image(1024) -> processor (demo.py:140) through CUDA end to end: the fp32 pixel_values hash
    equals the fixture produced by the reference's own demo.py functions + the live HF processor (tests/golden/pixels.npz,
    last entry), which round 1 could only check in the oracle."""
    from zoomearth_b200 import FusedImageProcessor
    z = np.load(os.path.join(GOLD, "pixels.npz"))
    m = json.loads(bytes(z["meta"]).decode())[-1]
    big = np.random.default_rng(102).integers(0, 256, (1300, 1700, 3), dtype=np.uint8)
    dev = torch.from_numpy(big).to(cuda)
    p = FusedImageProcessor(min_pixels=3136, max_pixels=200704, device=cuda)
    two, inv = p.cut_resize([dev], [m["bbox"]], image_index=[0], variant="demo")
    assert [two[0].shape[1], two[0].shape[0]] == m["two_step_size"] and _sha(two[0].contiguous()) == m["two_step_sha256"]
    pv, grid, _ = p.preprocess_crops(two, None, torch.float32, False)
    assert grid[0].tolist() == m["grid"] and _sha(pv) == m["pv_sha256"]


@pytest.mark.parametrize("variant,max_size", [("infer", 512), ("demo", 1024), ("custom", 512)])
def test_zoom_encoder_pre_resize_vs_oracle(cuda, variant, max_size):
    """ZoomEncoder(pre_resize=...): global view and zoom crops through resize_image(cut_image(...)) -> processor -> tower,
    all on the device; patches bit-exact and embeddings within tolerance of the oracle's two-resample flow."""
    from zoomearth_b200 import FusedImageProcessor, FusedVisual, ZoomEncoder
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(31, cfg)
    fv = FusedVisual(sd, device=cuda, dtype=torch.float32, depth=cfg["depth"], fullatt=list(cfg["fullatt"]))
    enc = ZoomEncoder(fv, FusedImageProcessor(min_pixels=3136, max_pixels=12845056, device=cuda), pre_resize=variant)
    img = np.random.default_rng(12).integers(0, 256, (1400, 1900, 3), dtype=np.uint8)
    dev = enc.upload(img)
    boxes = [(100, 150, 1500, 1250), (300.5, 200.2, 700.9, 650.0), (-30, -20, 800, 900), (1500, 900, 1890, 1390)]
    for bx in (None, boxes):
        emb, grid, crop, pv = enc.encode([dev], bx, image_index=None if bx is None else [0] * len(bx), return_patches=True)
        rows, grids = [], []
        for b in ([None] if bx is None else bx):
            im = img if b is None else OF.cut_image(img, b, 512, variant)
            im, inv = OF.resize_image(im, max_size, variant)
            r, g, _ = OP.preprocess_u8([im], 3136, 12845056)
            rows.append(r)
            grids.append(g)
        rgrid = np.concatenate(grids, 0)
        assert grid.tolist() == rgrid.tolist()
        ref_pv = np.concatenate(rows, 0)
        widx, _ = OT.window_index(rgrid)
        ref_w = torch.from_numpy(ref_pv).view(-1, 4, 1176)[torch.from_numpy(widx)].reshape(-1, 1176).to(fv.operand_dtype)
        assert torch.equal(pv.cpu(), ref_w), "two-resample patches differ from the oracle"
        if bx is not None:
            assert [tuple(int(v) for v in c) for c in crop] == [OG.cut_box(1900, 1400, b) for b in bx]
        ref = OT.forward(sd, torch.from_numpy(ref_pv), rgrid, cfg)
        got = emb.double().cpu()
        cos = (torch.dot(got.flatten(), ref.double().flatten()) / (got.norm() * ref.double().norm())).item()
        maxrel = ((got - ref.double()).abs().max() / ref.abs().max()).item()
        assert cos >= 0.999 and maxrel <= 1e-2, f"{variant}: cos {cos} maxrel {maxrel}"


def test_zoom_session_global_max_size(cuda):
    """ZoomSession(global_max_size=512): the model sees resize_image(image) (infer.py:215), scale() is the reference's
    1/scale, and stage 2 feeds resize_image(cut_image(image, bbox * scale)) (infer.py:226-239) - grids and patches as in
    the unmodified loop."""
    from zoomearth_b200 import FusedImageProcessor, FusedVisual, ZoomEncoder, ZoomSession, extract_bbox
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(32, cfg)
    fv = FusedVisual(sd, device=cuda, dtype=torch.float32, depth=cfg["depth"], fullatt=list(cfg["fullatt"]))
    enc = ZoomEncoder(fv, FusedImageProcessor(min_pixels=3136, max_pixels=12845056, device=cuda))
    sess = ZoomSession(enc, global_max_size=512)
    img = np.random.default_rng(13).integers(0, 256, (2000, 3000, 3), dtype=np.uint8)
    sess.add_image("q", img)
    assert sess.scale("q") == OG.resize_dims(3000, 2000, 512)[2]
    embs, grid = sess.stage1(["q"])
    g_img, _ = OF.resize_image(img, 512, "infer")
    g_pv, g_grid, _ = OP.preprocess_u8([g_img], 3136, 12845056)
    assert grid.tolist() == g_grid.tolist() and embs[0].shape[0] * 4 == g_pv.shape[0]
    box = extract_bbox('{"bbox_2d": [100, 60, 300, 250]}', sess.scale("q"))[0]        # model box on the 512-px view -> source pixels
    out, crop = sess.stage2(["q"], [box])
    c_img, _ = OF.resize_image(OF.cut_image(img, box, 512, "infer"), 512, "infer")
    c_pv, c_grid, _ = OP.preprocess_u8([c_img], 3136, 12845056)
    (pair, pgrid), = out
    assert pgrid.tolist() == [g_grid[0].tolist(), c_grid[0].tolist()]
    assert tuple(int(v) for v in crop[0]) == OG.cut_box(3000, 2000, box)
    for e, pv, g in ((pair[0], g_pv, g_grid), (pair[1], c_pv, c_grid)):
        ref = OT.forward(sd, torch.from_numpy(pv), g, cfg)
        got = e.double().cpu()
        cos = (torch.dot(got.flatten(), ref.double().flatten()) / (got.norm() * ref.double().norm())).item()
        maxrel = ((got - ref.double()).abs().max() / ref.abs().max()).item()
        assert cos >= 0.999 and maxrel <= 1e-2, f"cos {cos} maxrel {maxrel}"
