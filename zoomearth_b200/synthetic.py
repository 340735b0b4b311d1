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
