"""LM hand-off on the GPU: the tower's last GEMM writes embeddings straight into inputs_embeds
(zv_visual_forward_into).  Data movement only, so the bar is bit-exact: identical to
``inputs_embeds.masked_scatter(image_mask, visual(pixel_values, grid_thw))`` (HF modeling_qwen2_5_vl.py:1301-1307)."""
import numpy as np
import pytest
import torch

from oracle import handoff as OH, tower as OT

pytestmark = pytest.mark.gpu
IMG, VSTART, VEND, PAD = 151655, 151652, 151653, 151643


def _prompt(grid, rng, L=None):
    rows = []
    k = 0
    for n in (2, 1):
        ids = rng.integers(0, 1000, 3).tolist()
        for _ in range(n):
            t = int(grid[k, 1] * grid[k, 2]) // 4
            ids += [VSTART] + [IMG] * t + [VEND] + rng.integers(0, 1000, 2).tolist()
            k += 1
        rows.append(ids)
    L = max(len(r) for r in rows)
    out = np.full((len(rows), L), PAD, np.int64)
    for i, r in enumerate(rows):
        out[i, L - len(r):] = r
    return out


@pytest.mark.parametrize("embed_dtype", [torch.bfloat16, torch.float32])
def test_embed_images_equals_masked_scatter(cuda, embed_dtype):
    from zoomearth_b200 import FusedVisual, embed_images
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(21, cfg)
    fv = FusedVisual(sd, device=cuda, dtype=embed_dtype, operand_dtype=torch.bfloat16, depth=cfg["depth"],
                     fullatt=list(cfg["fullatt"]))
    grid = torch.tensor([[1, 16, 20], [1, 8, 8], [1, 26, 36]])
    rng = np.random.default_rng(2)
    ids = torch.from_numpy(_prompt(grid.numpy(), rng)).to(cuda)
    S = int((grid[:, 1] * grid[:, 2]).sum())
    pv = torch.randn(S, 1176, generator=torch.Generator().manual_seed(5)).to(cuda)
    emb0 = torch.randn(ids.shape[0], ids.shape[1], 2048, generator=torch.Generator().manual_seed(6)).to(embed_dtype).to(cuda)
    feats = fv(pv, grid)
    ref = emb0.masked_scatter((ids == IMG).unsqueeze(-1).expand_as(emb0), feats)
    got = embed_images(fv, emb0.clone(), ids, pv, grid)
    assert torch.equal(got, ref)
    # the oracle's statement of the same scatter
    assert np.array_equal(OH.masked_scatter(emb0.float().cpu().numpy(), ids.cpu().numpy(), feats.float().cpu().numpy(), IMG),
                          ref.float().cpu().numpy())
    short = ids.clone()
    short[0, int((ids[0] == IMG).nonzero()[0])] = PAD          # one placeholder fewer than embeddings
    with pytest.raises(ValueError, match="do not match"):
        embed_images(fv, emb0.clone(), short, pv, grid)


def test_embed_images_from_fused_preprocess(cuda):
    """Zoom fast path end to end: resident image -> K1 (window-ordered patches) -> tower -> inputs_embeds rows."""
    from zoomearth_b200 import FusedImageProcessor, FusedVisual, embed_images
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    fv = FusedVisual(OT.make_weights(22, cfg), device=cuda, dtype=torch.bfloat16, operand_dtype=torch.bfloat16, depth=cfg["depth"], fullatt=list(cfg["fullatt"]))
    proc = FusedImageProcessor(min_pixels=3136, max_pixels=200704, device=cuda)
    img = torch.from_numpy(np.random.default_rng(9).integers(0, 256, (700, 900, 3), dtype=np.uint8)).to(cuda)
    pv, grid, _ = proc.preprocess_crops([img], [(10, 20, 600, 500)], out_dtype=torch.bfloat16, window_order=True)
    T = int(grid[0, 1] * grid[0, 2]) // 4
    ids = torch.tensor([[1, 2, VSTART] + [IMG] * T + [VEND, 3]], device=cuda)
    emb0 = torch.zeros(1, ids.shape[1], 2048, dtype=torch.bfloat16, device=cuda)
    ref = emb0.clone()
    ref[0, 3:3 + T] = fv(pv, grid, window_order=True)
    got = embed_images(fv, emb0, ids, pv, grid, window_order=True)
    assert got.data_ptr() == emb0.data_ptr() and torch.equal(got, ref)
