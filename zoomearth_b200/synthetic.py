"""Seeded random-init weights of the Qwen2.5-VL-3B vision tower (HF state-dict names), for benchmarks and
smoke runs where no checkpoint is available (there is no network).  Benchmark input, not compute."""
import torch

VISION_3B = dict(depth=32, hidden=1280, heads=16, inter=3420, out_hidden=2048, patch=14, temporal=2, in_ch=3)


def random_vision_state_dict(seed=0, device="cpu", std=0.02, **overrides):
    c = dict(VISION_3B)
    c.update(overrides)
    g = torch.Generator(device=device).manual_seed(seed)
    H, I, O = c["hidden"], c["inter"], c["out_hidden"]
    K = c["in_ch"] * c["temporal"] * c["patch"] ** 2

    def lin(o, i):
        return torch.randn(o, i, generator=g, device=device) * std

    def vec(n, s=0.02, m=0.0):
        return torch.randn(n, generator=g, device=device) * s + m

    sd = {"patch_embed.proj.weight": lin(H, K).view(H, c["in_ch"], c["temporal"], c["patch"], c["patch"])}
    for l in range(c["depth"]):
        p = f"blocks.{l}."
        sd[p + "norm1.weight"] = vec(H, 0.1, 1.0)
        sd[p + "norm2.weight"] = vec(H, 0.1, 1.0)
        sd[p + "attn.qkv.weight"] = lin(3 * H, H)
        sd[p + "attn.qkv.bias"] = vec(3 * H)
        sd[p + "attn.proj.weight"] = lin(H, H)
        sd[p + "attn.proj.bias"] = vec(H)
        sd[p + "mlp.gate_proj.weight"] = lin(I, H)
        sd[p + "mlp.gate_proj.bias"] = vec(I)
        sd[p + "mlp.up_proj.weight"] = lin(I, H)
        sd[p + "mlp.up_proj.bias"] = vec(I)
        sd[p + "mlp.down_proj.weight"] = lin(H, I)
        sd[p + "mlp.down_proj.bias"] = vec(H)
    sd["merger.ln_q.weight"] = vec(H, 0.1, 1.0)
    sd["merger.mlp.0.weight"] = lin(4 * H, 4 * H)
    sd["merger.mlp.0.bias"] = vec(4 * H)
    sd["merger.mlp.2.weight"] = lin(O, 4 * H)
    sd["merger.mlp.2.bias"] = vec(O)
    return sd


# ------------------------------------------------------------------------------------------------------------------
# Crop-box workloads of BASELINE.json configs[2..4].  Boxes are in source-image pixels, (x0, y0, x1, y1),
# BEFORE the reference's cut_image() rule (src/eval/infer.py:41-76), which the encoder applies.
import numpy as np


def trajectory_boxes(n_questions=256, img=5000):
    """configs[2]: per question three nested boxes, (w1,h1)~U[1536,3072]^2 inside the image, (w2,h2)~U[768,1536]^2
    inside box 1, (w3,h3)~U[256,768]^2 inside box 2, rng = default_rng(1000 + q).  Returns (n_questions, 3, 4) int64."""
    out = np.zeros((n_questions, 3, 4), np.int64)
    for q in range(n_questions):
        rng = np.random.default_rng(1000 + q)
        x0, y0, w, h = 0, 0, img, img
        for d, (lo, hi) in enumerate(((1536, 3072), (768, 1536), (256, 768))):
            nw = int(rng.integers(lo, min(hi, w) + 1))
            nh = int(rng.integers(lo, min(hi, h) + 1))
            x0 += int(rng.integers(0, w - nw + 1))
            y0 += int(rng.integers(0, h - nh + 1))
            w, h = nw, nh
            out[q, d] = (x0, y0, x0 + w, y0 + h)
    return out


def mixed_crop_boxes(n=1024, n_images=16, img=5000, lo=256, hi=2048, seed=4):
    """configs[3]: n boxes with independent sides ~U[lo,hi] placed uniformly in one of n_images source images.
    Returns (boxes (n, 4) int64, image_index (n,) int64)."""
    rng = np.random.default_rng(seed)
    w = rng.integers(lo, hi + 1, n)
    h = rng.integers(lo, hi + 1, n)
    x = (rng.random(n) * (img - w + 1)).astype(np.int64)
    y = (rng.random(n) * (img - h + 1)).astype(np.int64)
    return np.stack([x, y, x + w, y + h], 1).astype(np.int64), rng.integers(0, n_images, n)


def maxres_boxes(n_3584=64, n_full=64, img=5000):
    """configs[4]: 3584x3584 boxes (grid 256x256 at max_pixels = 16384*28*28) and full 5000x5000 images (254x254)."""
    rng = np.random.default_rng(5)
    xy = rng.integers(0, img - 3584 + 1, (n_3584, 2))
    a = np.concatenate([xy, xy + 3584], 1)
    b = np.tile(np.array([[0, 0, img, img]]), (n_full, 1))
    return np.concatenate([a, b], 0).astype(np.int64)
