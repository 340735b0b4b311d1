"""Oracle: Qwen2.5-VL vision tower forward, restated in plain torch fp32.  Test infrastructure only.

Follows HF ``models/qwen2_5_vl/modeling_qwen2_5_vl.py``:
  rot_pos_ids        :382-409      window_index      :411-451
  rmsnorm            :57-71        patch embed       :91-114
  rotary             :117-130,149-167
  attention          :207-287      mlp               :77-88
  block              :290-321      merger            :133-146
  forward            :455-518
State-dict names are HF's (relative to the tower).  ``emulate_bf16=True`` rounds
GEMM operands to bf16 (fp32 accumulate, fp32 residual) - the precision policy of
the CUDA path - so index bugs can be told apart from rounding.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

CFG = dict(depth=32, hidden=1280, heads=16, inter=3420, out_hidden=2048, patch=14, merge=2,
           temporal=2, window=112, fullatt=(7, 15, 23, 31), in_ch=3, eps=1e-6)


def small_cfg(depth=2, fullatt=(1,)):
    c = dict(CFG)
    c.update(depth=depth, fullatt=tuple(fullatt))
    return c


# ---------------------------------------------------------------- integer bookkeeping (bit-exact artefacts)
def rot_pos_ids(grid_thw, merge=2):
    """(S, 2) int64 (h, w) ids in merge-group raster order.  HF :382-401."""
    out = []
    for t, h, w in np.asarray(grid_thw).tolist():
        hp = np.arange(h)[:, None].repeat(w, 1).reshape(h // merge, merge, w // merge, merge)
        wp = np.arange(w)[None, :].repeat(h, 0).reshape(h // merge, merge, w // merge, merge)
        hp = hp.transpose(0, 2, 1, 3).reshape(-1)
        wp = wp.transpose(0, 2, 1, 3).reshape(-1)
        out.append(np.tile(np.stack([hp, wp], -1), (t, 1)))
    return np.concatenate(out, 0).astype(np.int64)


def window_index(grid_thw, window=112, merge=2, patch=14):
    """(window_index (T,) int64, cu_window_seqlens list with HF's zero-length entries).  HF :411-451."""
    ws = window // merge // patch
    unit = merge * merge
    widx, cu, base = [], [0], 0
    for t, gh, gw in np.asarray(grid_thw).tolist():
        lh, lw = gh // merge, gw // merge
        index = np.arange(t * lh * lw).reshape(t, lh, lw)
        ph, pw = ws - lh % ws, ws - lw % ws                  # a FULL extra window when already divisible
        nh, nw = (lh + ph) // ws, (lw + pw) // ws
        pad = np.full((t, lh + ph, lw + pw), -100, np.int64)
        pad[:, :lh, :lw] = index
        pad = pad.reshape(t, nh, ws, nw, ws).transpose(0, 1, 3, 2, 4).reshape(t, nh * nw, ws, ws)
        seqlens = (pad != -100).sum((2, 3)).reshape(-1)
        flat = pad.reshape(-1)
        widx.append(flat[flat != -100] + base)
        cu.extend((np.cumsum(seqlens) * unit + cu[-1]).tolist())
        base += t * lh * lw
    return np.concatenate(widx).astype(np.int64), cu


def unique_consecutive(xs):
    out = [xs[0]]
    for v in xs[1:]:
        if v != out[-1]:
            out.append(v)
    return out


def cu_seqlens_full(grid_thw):
    g = np.asarray(grid_thw)
    lens = np.repeat(g[:, 1] * g[:, 2], g[:, 0])
    return [0] + np.cumsum(lens).tolist()


# ---------------------------------------------------------------- weights
def make_weights(seed=0, cfg=CFG, std=0.02):
    """Seeded random state dict with HF names; biases and norm gains are non-trivial on purpose."""
    g = torch.Generator().manual_seed(seed)
    H, I, O = cfg["hidden"], cfg["inter"], cfg["out_hidden"]
    K = cfg["in_ch"] * cfg["temporal"] * cfg["patch"] ** 2

    def lin(o, i):
        return torch.randn(o, i, generator=g) * std

    def vec(n, s=0.02, m=0.0):
        return torch.randn(n, generator=g) * s + m

    sd = {"patch_embed.proj.weight": lin(H, K).view(H, cfg["in_ch"], cfg["temporal"], cfg["patch"], cfg["patch"])}
    for l in range(cfg["depth"]):
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


# ---------------------------------------------------------------- forward
def _rms(x, w, eps):
    v = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x.float() * torch.rsqrt(v + eps))


def _mm(x, w, emulate_bf16):
    if emulate_bf16:
        return x.to(torch.bfloat16).float() @ w.to(torch.bfloat16).float().t()
    return x @ w.t()


def _attend(qs, ks, vs, hd, q_chunk):
    """softmax(q k^T / sqrt(hd)) v for one segment, (heads, n, hd) each; q rows in chunks so that a 65 536-patch
    segment (configs[4]) does not materialise a heads x n x n score tensor.  Same arithmetic chunked or not."""
    n = qs.shape[1]
    if n <= q_chunk:
        return torch.softmax(qs @ ks.transpose(1, 2) / math.sqrt(hd), -1) @ vs
    outs = []
    for a in range(0, n, q_chunk):
        outs.append(torch.softmax(qs[:, a:a + q_chunk] @ ks.transpose(1, 2) / math.sqrt(hd), -1) @ vs)
    return torch.cat(outs, 1)


def _attend_grouped(q, k, v, seg, hd, q_chunk):
    """Same per-segment attention with the segments of equal length stacked into one batched matmul (GPU checker only:
    thousands of 64-row windows would otherwise be a Python loop of tiny launches)."""
    S, heads, _ = q.shape
    starts = torch.tensor(seg[:-1], device=q.device)
    lens = torch.tensor(seg[1:], device=q.device) - starts
    out = torch.empty(S, heads * hd, dtype=q.dtype, device=q.device)
    for n in torch.unique(lens).tolist():
        st = starts[lens == n]
        for c0 in range(0, st.numel(), 4096):
            idx = (st[c0:c0 + 4096, None] + torch.arange(n, device=q.device)[None, :]).reshape(-1)
            qs, ks, vs = (t[idx].view(-1, n, heads, hd).permute(0, 2, 1, 3).reshape(-1, n, hd) for t in (q, k, v))
            o = _attend(qs, ks, vs, hd, q_chunk)                                                 # (G * heads, n, hd)
            out[idx] = o.view(-1, heads, n, hd).permute(0, 2, 1, 3).reshape(-1, heads * hd)
    return out


def forward(sd, pixel_values, grid_thw, cfg=CFG, emulate_bf16=False, return_hidden=False, device=None, q_chunk=4096):
    """pixel_values (S,1176) f32 in HF row order, grid_thw (N,3) -> (T, out_hidden) f32 in HF output order.

    ``device``: run the same fp32 arithmetic on that device (the GPU tests use ``cuda`` as the checker for shapes the
    CPU cannot finish: TF32 is switched off for the call, so every matmul is IEEE fp32); the result comes back on
    the CPU.  Weights are moved per call - the caller may pass a state dict that already lives on the device."""
    if device is not None and torch.device(device).type == "cuda":
        old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.get_float32_matmul_precision())
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        torch.set_float32_matmul_precision("highest")
        try:
            sd_d = {k: v.to(device) for k, v in sd.items()}
            out = _forward(sd_d, pixel_values.to(device), grid_thw, cfg, emulate_bf16, return_hidden, q_chunk)
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old[0], old[1]
            torch.set_float32_matmul_precision(old[2])
        return tuple(t.cpu() for t in out) if return_hidden else out.cpu()
    return _forward(sd, pixel_values, grid_thw, cfg, emulate_bf16, return_hidden, q_chunk)


def _forward(sd, pixel_values, grid_thw, cfg, emulate_bf16, return_hidden, q_chunk):
    dev = pixel_values.device
    H, heads = cfg["hidden"], cfg["heads"]
    hd = H // heads
    eps = cfg["eps"]
    unit = cfg["merge"] ** 2
    bf = emulate_bf16
    x = _mm(pixel_values.float(), sd["patch_embed.proj.weight"].reshape(H, -1), bf)
    S = x.shape[0]
    pos = torch.from_numpy(rot_pos_ids(grid_thw, cfg["merge"])).to(dev)
    inv_freq = (1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float) / (hd // 2)))).to(dev)
    rot = torch.cat([pos[:, 0:1].float() * inv_freq, pos[:, 1:2].float() * inv_freq], -1)       # (S, hd/2)
    widx_np, cu_win = window_index(grid_thw, cfg["window"], cfg["merge"], cfg["patch"])
    widx = torch.from_numpy(widx_np).to(dev)
    cu_win = unique_consecutive(cu_win)
    cu_full = cu_seqlens_full(grid_thw)
    x = x.view(S // unit, unit, H)[widx].reshape(S, H)
    rot = rot.view(S // unit, unit, -1)[widx].reshape(S, -1)
    emb = torch.cat([rot, rot], -1)
    cos, sin = emb.cos()[:, None, :], emb.sin()[:, None, :]

    def rope(t):                                                                                  # (S, heads, hd)
        half = hd // 2
        r = torch.cat([-t[..., half:], t[..., :half]], -1)
        return t * cos + r * sin

    for l in range(cfg["depth"]):
        p = f"blocks.{l}."
        seg = cu_full if l in cfg["fullatt"] else cu_win
        y = _rms(x, sd[p + "norm1.weight"], eps)
        qkv = _mm(y, sd[p + "attn.qkv.weight"], bf) + sd[p + "attn.qkv.bias"]
        q, k, v = qkv.view(S, 3, heads, hd).unbind(1)
        q, k = rope(q), rope(k)
        if bf:
            q, k, v = (t.to(torch.bfloat16).float() for t in (q, k, v))
        if dev.type == "cuda" and len(seg) > 65:
            att = _attend_grouped(q, k, v, seg, hd, q_chunk)
        else:
            outs = []
            for a, b in zip(seg[:-1], seg[1:]):
                qs, ks, vs = (t[a:b].transpose(0, 1) for t in (q, k, v))                         # (heads, n, hd)
                outs.append(_attend(qs, ks, vs, hd, q_chunk).transpose(0, 1).reshape(b - a, H))
            att = torch.cat(outs, 0)
        x = x + _mm(att, sd[p + "attn.proj.weight"], bf) + sd[p + "attn.proj.bias"]
        y = _rms(x, sd[p + "norm2.weight"], eps)
        gate = _mm(y, sd[p + "mlp.gate_proj.weight"], bf) + sd[p + "mlp.gate_proj.bias"]
        up = _mm(y, sd[p + "mlp.up_proj.weight"], bf) + sd[p + "mlp.up_proj.bias"]
        x = x + _mm(F.silu(gate) * up, sd[p + "mlp.down_proj.weight"], bf) + sd[p + "mlp.down_proj.bias"]
    z = _rms(x, sd["merger.ln_q.weight"], eps).view(S // unit, unit * H)
    z = F.gelu(_mm(z, sd["merger.mlp.0.weight"], bf) + sd["merger.mlp.0.bias"])
    z = _mm(z, sd["merger.mlp.2.weight"], bf) + sd["merger.mlp.2.bias"]
    out = z[torch.argsort(widx)]
    if return_hidden:
        return out, x
    return out
