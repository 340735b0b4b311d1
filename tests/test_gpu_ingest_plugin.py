"""GPU: decode-once ingest feeding the zoom session from file paths, and the vLLM vision-transformer adapter against the
oracle tower."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import flows as OF, processor as OP, tower as OT

pytestmark = pytest.mark.gpu


def _metrics(got, ref):
    got, ref = got.detach().double().cpu().flatten(), ref.detach().double().cpu().flatten()
    return (torch.dot(got, ref) / (got.norm() * ref.norm())).item(), ((got - ref).abs().max() / ref.abs().max()).item()


def test_session_from_paths_decodes_each_file_once(cuda, tmp_path):
    """The reference loop opens the file for stage 1 and again for stage 2 (infer.py:215,237); ZoomSession.add_image(path)
    decodes it once, keeps it resident, and both stages read the resident pixels.  Embeddings against the oracle of the
    reference-faithful flow on the pixels Pillow decodes."""
    from zoomearth_b200 import FusedImageProcessor, FusedVisual, ZoomEncoder, ZoomSession
    from zoomearth_b200.ingest import ImageStore
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(41, cfg)
    fv = FusedVisual(sd, device=cuda, dtype=torch.float32, depth=cfg["depth"], fullatt=list(cfg["fullatt"]))
    enc = ZoomEncoder(fv, FusedImageProcessor(min_pixels=3136, max_pixels=12845056, device=cuda))
    store = ImageStore(device=cuda)
    sess = ZoomSession(enc, global_max_size=512, store=store)
    rng = np.random.default_rng(17)
    paths = []
    for i, (fmt, ext) in enumerate((("TIFF", "tif"), ("PNG", "png"))):
        p = str(tmp_path / f"img{i}.{ext}")
        Image.fromarray(rng.integers(0, 256, (1100 + 200 * i, 1600, 3), dtype=np.uint8)).save(p, format=fmt)
        paths.append(p)
    store.prefetch(paths)
    for p in paths:
        sess.add_image(p, p)
    embs, grid = sess.stage1(paths)
    out, crop = sess.stage2(paths, [(300, 200, 1200, 900), (50.5, 60.2, 400.0, 300.9)])
    assert store.decodes == 2
    for i, p in enumerate(paths):
        img = np.asarray(Image.open(p).convert("RGB"))
        assert torch.equal(sess.add_image(p, p).cpu(), torch.from_numpy(img))
        g_img, _ = OF.resize_image(img, 512, "infer")
        pv, g, _ = OP.preprocess_u8([g_img], 3136, 12845056)
        cos, maxrel = _metrics(embs[i], OT.forward(sd, torch.from_numpy(pv), g, cfg))
        assert grid[i].tolist() == g[0].tolist() and cos >= 0.999 and maxrel <= 1e-2
    store.close()


def test_vllm_adapter_forward_vs_oracle(cuda):
    """FusedVllmVisionTransformer with vLLM's call signature (x, grid_thw as list of lists) on checkpoint-named weights."""
    pytest.importorskip("vllm")
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLVisionConfig
    from zoomearth_b200.vllm_plugin import FusedVllmVisionTransformer
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(42, cfg)
    vc = Qwen2_5_VLVisionConfig(depth=2, hidden_size=1280, intermediate_size=3420, num_heads=16, out_hidden_size=2048,
                                patch_size=14, spatial_merge_size=2, temporal_patch_size=2, window_size=112,
                                fullatt_block_indexes=[1], hidden_act="silu")
    m = FusedVllmVisionTransformer(vc, norm_eps=1e-6, dtype=torch.float16, device=cuda)
    assert m.load_weights(list(sd.items())) == set(sd)
    grid = [[1, 16, 20], [1, 8, 8]]
    pv = torch.randn(16 * 20 + 64, 1176, generator=torch.Generator().manual_seed(3))
    out = m(pv.to(cuda), grid)
    assert out.dtype == torch.float16 and out.shape == (96, 2048) and m.dtype == torch.float16
    cos, maxrel = _metrics(out, OT.forward(sd, pv, np.array(grid), cfg))
    assert cos >= 0.999 and maxrel <= 1e-2, f"cos {cos} maxrel {maxrel}"
