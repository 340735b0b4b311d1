"""BASELINE.json configs as parity cases (GPU, through the C ABI), at sizes the CPU oracle finishes in seconds, plus
size-independent properties at the full sizes."""
import numpy as np
import pytest
import torch

from oracle import geometry as OG, processor as OP, tower as OT

pytestmark = pytest.mark.gpu


def _metrics(got, ref):
    """(cosine of the flattened tensors, max|a-b| / max|b|), accumulated in float64 (fp32 sums over 1e8 elements drift)."""
    got, ref = got.detach().double().cpu().flatten(), ref.detach().double().cpu().flatten()
    cos = (torch.dot(got, ref) / (got.norm() * ref.norm())).item()
    return cos, ((got - ref).abs().max() / ref.abs().max()).item()


def _encoder(cuda, cfg, sd, max_pixels, operand_dtype=torch.float16):
    from zoomearth_b200 import FusedImageProcessor, FusedVisual, ZoomEncoder
    fv = FusedVisual(sd, device=cuda, dtype=torch.float32, operand_dtype=operand_dtype, depth=cfg["depth"],
                     fullatt=list(cfg["fullatt"]))
    return ZoomEncoder(fv, FusedImageProcessor(min_pixels=3136, max_pixels=max_pixels, device=cuda))


def test_config1_one_zoom_step_on_5000px_image(cuda):
    """configs[0]: one synthetic 5000x5000 image, bbox (1000,1200,2300,2100), max_pixels 12845056 ->
    crop 1300x900 -> 896x1288 -> grid (1,64,92), 5888 patches; 4-block tower vs the fp32 oracle."""
    cfg = OT.small_cfg(depth=4, fullatt=(1, 3))
    sd = OT.make_weights(11, cfg)
    img = np.random.default_rng(0).integers(0, 256, (5000, 5000, 3), dtype=np.uint8)
    enc = _encoder(cuda, cfg, sd, 12845056)
    emb, grid, crop = enc.encode([enc.upload(img)], [(1000, 1200, 2300, 2100)])
    assert tuple(int(v) for v in crop[0]) == (1000, 1200, 2300, 2100) and grid.tolist() == [[1, 64, 92]]
    _, pv, rgrid = OP.zoom_step_u8(img, (1000, 1200, 2300, 2100), 512, 3136, 12845056)
    ref = OT.forward(sd, torch.from_numpy(pv), rgrid, cfg)
    cos, maxrel = _metrics(emb, ref)
    assert emb.shape == (1472, 2048) and cos >= 0.999 and maxrel <= 1e-2, f"cos {cos} maxrel {maxrel}"


def test_config3_nested_zoom_trajectory_ragged(cuda):
    """configs[2] in miniature: per question three nested boxes through cut_image (min 512), one ragged batch per zoom
    depth mixing crop sizes; boxes, grids and embeddings against the oracle."""
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(12, cfg)
    enc = _encoder(cuda, cfg, sd, 401408)
    imgs = [np.random.default_rng(100 + q).integers(0, 256, (1800, 2200, 3), dtype=np.uint8) for q in range(3)]
    dev = [enc.upload(i) for i in imgs]
    for depth in range(3):
        boxes = []
        for q in range(3):
            rng = np.random.default_rng(1000 + q)
            x0, y0, w, h = 0, 0, 2200, 1800
            for d in range(depth + 1):
                lo, hi = [(900, 1700), (450, 900), (120, 450)][d]
                nw, nh = int(rng.integers(lo, min(hi, w))), int(rng.integers(lo, min(hi, h)))
                x0, y0 = x0 + int(rng.integers(0, w - nw + 1)), y0 + int(rng.integers(0, h - nh + 1))
                w, h = nw, nh
            boxes.append((x0, y0, x0 + w, y0 + h))
        emb, grid, crop = enc.encode(dev, boxes, image_index=[0, 1, 2])
        rows, grids = [], []
        for q, b in enumerate(boxes):
            box, pv, g = OP.zoom_step_u8(imgs[q], b, 512, 3136, 401408)
            assert tuple(int(v) for v in crop[q]) == box
            rows.append(pv)
            grids.append(g)
        rgrid = np.concatenate(grids, 0)
        assert grid.tolist() == rgrid.tolist()
        ref = OT.forward(sd, torch.from_numpy(np.concatenate(rows, 0)), rgrid, cfg)
        cos, maxrel = _metrics(emb, ref)
        assert cos >= 0.999 and maxrel <= 1e-2, f"depth {depth}: cos {cos} maxrel {maxrel}"


def test_config4_mixed_crops_batch_equals_per_crop(cuda):
    """configs[3] property: a ragged batch of mixed-size crops gives, crop by crop, what each crop gives alone
    (crops are independent sequences - the basis of sharding by crop), and the LPT partition covers every crop once."""
    from zoomearth_b200 import sharding
    cfg = OT.small_cfg(depth=2, fullatt=(0,))
    sd = OT.make_weights(13, cfg)
    enc = _encoder(cuda, cfg, sd, 12845056, operand_dtype=torch.bfloat16)
    img = np.random.default_rng(4).integers(0, 256, (2400, 2400, 3), dtype=np.uint8)
    dev = enc.upload(img)
    rng = np.random.default_rng(4)
    boxes = []
    for _ in range(10):
        w, h = int(rng.integers(256, 1400)), int(rng.integers(256, 1400))
        x, y = int(rng.integers(0, 2400 - w)), int(rng.integers(0, 2400 - h))
        boxes.append((x, y, x + w, y + h))
    emb, grid, crop = enc.encode([dev], boxes, image_index=[0] * 10)
    tokens = enc.tokens_per_crop(grid.numpy())
    starts = np.concatenate([[0], np.cumsum(tokens)])
    for i in (0, 3, 9):
        e1, g1, _ = enc.encode([dev], [boxes[i]], image_index=[0])
        assert g1.tolist() == [grid[i].tolist()]
        cos, maxrel = _metrics(emb[starts[i]:starts[i + 1]], e1)
        assert cos >= 0.9995 and maxrel <= 1e-2, f"crop {i}: cos {cos} maxrel {maxrel}"
    parts = sharding.partition(sharding.crop_cost(grid.numpy()), 4)
    assert sorted(i for p in parts for i in p) == list(range(10))


def test_config5_max_resolution_crop(cuda):
    """configs[4]: a 3584x3584 crop at max_pixels = 16384*28*28 -> grid (1,256,256), 65 536 patches, one full-attention
    segment of 65 536.  Too large for the CPU oracle; checked through properties: geometry bit-exact, patches bit-exact
    against live Pillow on a sampled set of rows, embeddings finite, and the first-window embeddings equal those of a
    run where the tower sees the same patches in HF order (gather path) instead of the fused window order."""
    from PIL import Image
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(14, cfg)
    enc = _encoder(cuda, cfg, sd, 16384 * 28 * 28, operand_dtype=torch.bfloat16)
    img = np.random.default_rng(5).integers(0, 256, (3700, 3700, 3), dtype=np.uint8)
    box = (50, 60, 3634, 3644)
    dev = enc.upload(img)
    emb, grid, crop, pv = enc.encode([dev], [box], return_patches=True)
    assert grid.tolist() == [[1, 256, 256]] and emb.shape == (16384, 2048)
    assert OG.smart_resize(3584, 3584, 28, 3136, 16384 * 28 * 28) == (3584, 3584)
    assert torch.isfinite(emb).all()
    # same-size crop: Pillow copies the pixels; window-ordered bf16 patches must equal the LUT of the raw crop
    cropped = np.asarray(Image.fromarray(img).crop(box))
    lut = OP.normalize_lut()
    ref_pv, _ = OP.patchify(np.stack([lut[c][cropped[:, :, c]] for c in range(3)], 0))
    widx, _ = OT.window_index(np.array([[1, 256, 256]]))
    ref_w = torch.from_numpy(ref_pv).view(-1, 4, 1176)[torch.from_numpy(widx)].reshape(-1, 1176).to(torch.bfloat16)
    assert torch.equal(pv.cpu(), ref_w)
    emb_hf = enc.visual(torch.from_numpy(ref_pv).to(cuda), grid)            # fp32 patches, HF order -> gather path
    cos, maxrel = _metrics(emb, emb_hf)
    assert cos >= 0.9999 and maxrel <= 5e-3, f"cos {cos} maxrel {maxrel}"


def _k1_both(enc, images, boxes=None, image_index=None, apply_cut_image=True):
    """K1 twice over the same crops: fp32 patches in HF order (= the HF processor's pixel_values, the oracle tower's
    input) and the 16-bit window-ordered patches the fused path feeds the tower."""
    hf, grid, crop = enc.processor.preprocess_crops(images, boxes, torch.float32, False, image_index=image_index,
                                                    apply_cut_image=apply_cut_image and boxes is not None)
    return hf, grid, crop


def test_config2_bench_workload_full_depth_vs_gpu_oracle(cuda, full_sd_cuda, full_visual):
    """configs[1], the workload bench.py times, at its full size and depth: 64 synthetic 5000x5000 images -> global view
    980x980 (grid 1x70x70) -> all 32 blocks, fp16 operands, against the fp32 oracle tower run on the GPU (TF32 off).
    Image 0's pixel_values are pinned bit-exact to the live HF PIL processor; the oracle consumes K1's fp32 output."""
    from oracle import hf_live
    from zoomearth_b200 import FusedImageProcessor, ZoomEncoder
    n_img = 64
    enc = ZoomEncoder(full_visual, FusedImageProcessor(min_pixels=3136, max_pixels=1280 * 28 * 28, device=cuda))
    g = torch.Generator(device=cuda).manual_seed(99)
    images = [torch.randint(0, 256, (5000, 5000, 3), generator=g, dtype=torch.uint8, device=cuda) for _ in range(n_img)]
    pv_hf, grid, _ = _k1_both(enc, images)
    assert grid.tolist() == [[1, 70, 70]] * n_img
    ref_pv0, ref_grid0 = hf_live.hf_preprocess([images[0].cpu().numpy()], 3136, 1280 * 28 * 28)
    assert torch.equal(pv_hf[:4900].cpu(), ref_pv0) and ref_grid0.tolist() == [[1, 70, 70]]
    emb, grid2, _ = enc.encode(images, None)                       # the timed path: fp16 window-ordered patches -> tower
    del images
    assert emb.shape == (n_img * 1225, 2048) and grid2.tolist() == grid.tolist()
    ref = OT.forward(full_sd_cuda, pv_hf, grid.numpy(), device=cuda)
    cos, maxrel = _metrics(emb, ref)
    worst = max(_metrics(emb[i * 1225:(i + 1) * 1225], ref[i * 1225:(i + 1) * 1225])[1] for i in range(n_img))
    print(f"PARITY configs[1] 64 x 70x70, depth 32, fp16 operands vs fp32 oracle (GPU): cosine {cos:.6f}, "
          f"max rel err {maxrel:.3e} (worst single image {worst:.3e})")
    assert cos >= 0.999 and maxrel <= 1e-2 and worst <= 1e-2, f"cos {cos} maxrel {maxrel} worst image {worst}"


def test_config4_mixed_crops_full_depth_crop_by_crop_vs_oracle(cuda, full_sd_cuda, full_visual):
    """configs[3] in miniature at full depth: 12 mixed-size crops (256-1400 px per side, cut_image rule) as ONE ragged
    batch; every crop's embeddings against the fp32 oracle of that crop (not against another run of the CUDA path)."""
    from zoomearth_b200 import FusedImageProcessor, ZoomEncoder
    enc = ZoomEncoder(full_visual, FusedImageProcessor(min_pixels=3136, max_pixels=12845056, device=cuda))
    img = np.random.default_rng(4).integers(0, 256, (2400, 2400, 3), dtype=np.uint8)
    dev = enc.upload(img)
    rng = np.random.default_rng(44)
    boxes = []
    for _ in range(12):
        w, h = int(rng.integers(256, 1400)), int(rng.integers(256, 1400))
        x, y = int(rng.integers(0, 2400 - w)), int(rng.integers(0, 2400 - h))
        boxes.append((x, y, x + w, y + h))
    emb, grid, crop = enc.encode([dev], boxes, image_index=[0] * 12)
    tokens = enc.tokens_per_crop(grid.numpy())
    starts = np.concatenate([[0], np.cumsum(tokens)])
    for i, b in enumerate(boxes):
        box, pv, g = OP.zoom_step_u8(img, b, 512, 3136, 12845056)
        assert tuple(int(v) for v in crop[i]) == box and grid[i].tolist() == g[0].tolist()
        ref = OT.forward(full_sd_cuda, torch.from_numpy(pv), g, device=cuda)
        cos, maxrel = _metrics(emb[starts[i]:starts[i + 1]], ref)
        assert cos >= 0.999 and maxrel <= 1e-2, f"crop {i} {box}: cos {cos} maxrel {maxrel}"


def test_config5_max_resolution_full_depth_vs_gpu_oracle(cuda, full_sd_cuda, full_visual):
    """configs[4]: one 3584x3584 crop at max_pixels = 16384*28*28 -> grid (1,256,256): a single full-attention segment
    of 65 536 patches through all 32 blocks, against the fp32 oracle on the GPU (q rows in chunks of 2048)."""
    from zoomearth_b200 import FusedImageProcessor, ZoomEncoder
    enc = ZoomEncoder(full_visual, FusedImageProcessor(min_pixels=3136, max_pixels=16384 * 28 * 28, device=cuda))
    g = torch.Generator(device=cuda).manual_seed(5)
    img = torch.randint(0, 256, (3700, 3700, 3), generator=g, dtype=torch.uint8, device=cuda)
    box = (50, 60, 3634, 3644)
    pv_hf, grid, crop = _k1_both(enc, [img], [box])
    assert grid.tolist() == [[1, 256, 256]] and tuple(int(v) for v in crop[0]) == box
    emb, _, _ = enc.encode([img], [box])
    ref = OT.forward(full_sd_cuda, pv_hf, grid.numpy(), device=cuda, q_chunk=2048)
    cos, maxrel = _metrics(emb, ref)
    print(f"PARITY configs[4] one 65 536-patch segment, depth 32, fp16 operands vs fp32 oracle (GPU): cosine {cos:.6f}, "
          f"max rel err {maxrel:.3e}")
    assert emb.shape == (16384, 2048) and cos >= 0.999 and maxrel <= 1e-2, f"cos {cos} maxrel {maxrel}"


def test_install_swaps_hf_visual_and_matches_it(cuda):
    """Drop-in through HF's own call path: install() replaces model.visual of a (tiny-LM) Qwen2.5-VL model; the HF
    get_image_features() then returns what the original HF tower returned, within the stated tolerance."""
    import transformers
    from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration
    from zoomearth_b200 import install
    vis = dict(depth=2, hidden_size=1280, intermediate_size=3420, num_heads=16, out_hidden_size=2048, patch_size=14,
               spatial_merge_size=2, temporal_patch_size=2, window_size=112, fullatt_block_indexes=[1], hidden_act="silu")
    text = dict(hidden_size=2048, intermediate_size=256, num_hidden_layers=1, num_attention_heads=16, num_key_value_heads=2,
                vocab_size=152064, max_position_embeddings=4096,
                rope_parameters={"rope_type": "default", "mrope_section": [16, 24, 24], "rope_theta": 1000000.0})
    try:
        config = Qwen2_5_VLConfig(vision_config=vis, text_config=text)
    except Exception as e:                       # config schema differs across transformers majors
        pytest.skip(f"cannot build a tiny Qwen2_5_VLConfig on transformers {transformers.__version__}: {e}")
    torch.manual_seed(0)
    model = Qwen2_5_VLForConditionalGeneration(config).eval()
    owner = model if hasattr(model, "visual") else model.model
    hf_visual = owner.visual
    grid = torch.tensor([[1, 16, 20], [1, 8, 8]])
    pv = torch.randn(int((grid[:, 1] * grid[:, 2]).sum()), 1176, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = hf_visual(pv, grid_thw=grid)
    ref = ref.pooler_output if hasattr(ref, "pooler_output") else ref
    with torch.no_grad():
        ref_feat = owner.get_image_features(pv, grid) if hasattr(owner, "get_image_features") else None
    fv, _ = install(model, device=cuda, dtype=torch.float32)
    assert owner.visual is fv and fv.dtype == torch.float32 and fv.spatial_merge_size == 2
    out = owner.visual(pv.to(cuda), grid_thw=grid)
    out = out.pooler_output if hasattr(out, "pooler_output") else out
    cos, maxrel = _metrics(out, ref)
    assert cos >= 0.999 and maxrel <= 1e-2, f"cos {cos} maxrel {maxrel}"
    # ... and through HF's own call path: get_image_features() casts pixel_values to visual.dtype, calls
    # self.visual(pixel_values, grid_thw=...) and splits the result per image (HF modeling_qwen2_5_vl.py:1159-1177)
    if ref_feat is not None:
        with torch.no_grad():
            feat = owner.get_image_features(pv.to(cuda), grid.to(cuda))

        def _per_image(f):
            f = f.pooler_output if hasattr(f, "pooler_output") else f
            return list(f) if isinstance(f, (list, tuple)) else [f]
        got, want = _per_image(feat), _per_image(ref_feat)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert a.shape == b.shape
            cos, maxrel = _metrics(a, b)
            assert cos >= 0.999 and maxrel <= 1e-2, f"get_image_features: cos {cos} maxrel {maxrel}"


def test_zoom_session_caches_global_view_and_batches_crops(cuda):
    """ZoomSession: stage 1 is computed once per image; stage 2 returns, per question, the [global, crop] pair the
    reference hands the model, equal to encoding them directly."""
    from zoomearth_b200 import ZoomSession
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(15, cfg)
    enc = _encoder(cuda, cfg, sd, 200704, operand_dtype=torch.bfloat16)
    sess = ZoomSession(enc)
    imgs = {k: np.random.default_rng(50 + i).integers(0, 256, (900 + 100 * i, 1200, 3), dtype=np.uint8)
            for i, k in enumerate("ab")}
    for k, im in imgs.items():
        sess.add_image(k, im)
    e1, g1 = sess.stage1(["a", "b"])
    launches = enc.last_launches
    e1b, _ = sess.stage1(["b", "a"])
    assert e1b[0].data_ptr() == e1[1].data_ptr() and enc.last_launches == launches      # served from the cache
    keys, boxes = ["a", "b", "a"], [(100, 100, 700, 600), (50.5, 60.2, 300.0, 200.9), (400, 300, 1100, 880)]
    out, crop = sess.stage2(keys, boxes)
    assert len(out) == 3
    for (embs, grid), k, b, c in zip(out, keys, boxes, crop):
        assert tuple(int(v) for v in c) == sess.crop_box(k, b) == OG.cut_box(1200, imgs[k].shape[0], b)
        ref, rgrid, _ = enc.encode([sess.add_image(k, None)], [b], image_index=[0])
        assert grid[1].tolist() == rgrid[0].tolist() and embs[0].shape[0] * 4 == int(grid[0, 1] * grid[0, 2])
        cos, maxrel = _metrics(embs[1], ref)
        assert cos >= 0.9995 and maxrel <= 1e-2


def test_cuda_graph_replay_matches_eager(cuda):
    """forward(use_graph=True) replays a captured graph of the same launch sequence: identical results, also on reuse."""
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(16, cfg)
    enc = _encoder(cuda, cfg, sd, 401408, operand_dtype=torch.bfloat16)
    img = np.random.default_rng(8).integers(0, 256, (1500, 1500, 3), dtype=np.uint8)
    dev = enc.upload(img)
    for boxes in ([(100, 100, 612, 612)], [(0, 0, 900, 700), (300, 200, 1400, 1300)], [(100, 100, 612, 612)]):
        eager, g1, _ = enc.encode([dev], boxes, image_index=[0] * len(boxes))
        graph, g2, _ = enc.encode([dev], boxes, image_index=[0] * len(boxes), use_graph=True)
        assert g1.tolist() == g2.tolist() and torch.equal(eager, graph)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on the box")
def test_fused_peer_gather_equals_nccl_two_gpus():
    """tools/multi_gpu_check.py under torchrun: the merger GEMM's peer stores reproduce NCCL's all-gather bitwise."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29561", os.path.join(root, "tools", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.count("fused gather == nccl all_gather: True") == 2, r.stdout + r.stderr


def test_encode_batched_equals_one_batch(cuda):
    """Micro-batching a ragged crop list (bounded workspace) changes nothing but rounding: every crop is an independent
    sequence (only the position of a segment inside its batch moves the KV tile boundaries of the full-attention
    layers, i.e. the order of the fp32 sums)."""
    from zoomearth_b200 import synthetic
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(17, cfg)
    enc = _encoder(cuda, cfg, sd, 401408, operand_dtype=torch.bfloat16)
    imgs = [np.random.default_rng(70 + i).integers(0, 256, (1600, 1600, 3), dtype=np.uint8) for i in range(2)]
    dev = [enc.upload(i) for i in imgs]
    boxes, index = synthetic.mixed_crop_boxes(9, n_images=2, img=1600, lo=200, hi=1200, seed=3)
    whole, grid, crop = enc.encode(dev, boxes, image_index=index)
    parts, grid2, crop2 = enc.encode_batched(dev, boxes, index, max_patches=1500)
    assert len(enc.micro_batches(dev, boxes, index, max_patches=1500)) > 2
    assert grid.tolist() == grid2.tolist() and np.array_equal(crop, crop2) and parts.shape == whole.shape
    cos, maxrel = _metrics(parts, whole)
    assert cos >= 0.9999 and maxrel <= 5e-3, f"cos {cos} maxrel {maxrel}"
