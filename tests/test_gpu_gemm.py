"""tcgen05 GEMM parity (GPU, through the C ABI): every fused epilogue vs a plain torch fp32 reference on
bf16-rounded operands.  Tolerances: fp32 outputs 2e-3 relative to max|ref| (accumulation order), bf16
outputs one bf16 ulp of max|ref| on top."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(m, n, k, dev, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = (torch.randn(m, k, generator=g) * scale).to(torch.bfloat16).to(dev)
    b = (torch.randn(n, k, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    bias = torch.randn(n, generator=g).to(dev)
    return a, b, bias


_ZT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _run(lib, epi, a, b, bias, out, m, n, k, pos=None, rope=None, scatter=None, heads=16, lda=None, ldb=None):
    from zoomearth_b200 import _lib
    stream = torch.cuda.current_stream().cuda_stream
    dt = _ZT[out.dtype]
    _lib.check(lib.zv_gemm_ex(epi, a.data_ptr(), lda or a.stride(0), b.data_ptr(), ldb or b.stride(0),
                              None if bias is None else bias.data_ptr(), out.data_ptr(), out.stride(0), dt, m, n, k,
                              None if pos is None else pos.data_ptr(), None if rope is None else rope.data_ptr(),
                              None if scatter is None else scatter.data_ptr(), heads, _ZT[a.dtype], stream))
    torch.cuda.synchronize()


def _close(got, ref, tol):
    err = (got.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * scale, f"max err {err:.4g} vs scale {scale:.4g} (tol {tol})"


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (128, 128, 128), (300, 1280, 1280), (1000, 1280, 1176),
                                   (77, 384, 200), (4096, 2048, 5120), (129, 256, 3456)])
def test_gemm_store_fp32(cuda, lib, m, n, k):
    a, b, bias = _mk(m, n, k, cuda, seed=m + n + k)
    out = torch.full((m, n), float("nan"), device=cuda)
    _run(lib, 0, a, b, bias, out, m, n, k)
    ref = a.float() @ b.float().t() + bias
    _close(out, ref, 2e-3)


def test_gemm_store_bf16_no_bias(cuda, lib):
    m, n, k = 513, 512, 320
    a, b, _ = _mk(m, n, k, cuda, seed=3)
    out = torch.zeros((m, n), dtype=torch.bfloat16, device=cuda)
    _run(lib, 0, a, b, None, out, m, n, k)
    _close(out, a.float() @ b.float().t(), 6e-3)


@pytest.mark.parametrize("m,k", [(700, 3456), (1296, 1280), (20000, 1280)])
def test_gemm_residual(cuda, lib, m, k):
    """small M takes the 128-wide residual tile (twice the CTA pairs in flight), large M the 256-wide one"""
    n = 1280
    a, b, bias = _mk(m, n, k, cuda, seed=4)
    x0 = torch.randn(m, n, device=cuda)
    x = x0.clone()
    _run(lib, 2, a, b, bias, x, m, n, k)
    _close(x, x0 + a.float() @ b.float().t() + bias, 2e-3)


def test_gemm_swiglu(cuda, lib):
    m, inter, k = 333, 384, 1280            # packed: per 128 outputs, 128 gate rows then 128 up rows
    g = torch.Generator().manual_seed(5)
    a = torch.randn(m, k, generator=g).to(torch.bfloat16).to(cuda)
    wg = (torch.randn(inter, k, generator=g) * 0.05).to(torch.bfloat16).to(cuda)
    wu = (torch.randn(inter, k, generator=g) * 0.05).to(torch.bfloat16).to(cuda)
    bg, bu = torch.randn(inter, generator=g).to(cuda), torch.randn(inter, generator=g).to(cuda)
    w = torch.stack([wg.view(-1, 128, k), wu.view(-1, 128, k)], 1).reshape(2 * inter, k).contiguous()
    bias = torch.stack([bg.view(-1, 128), bu.view(-1, 128)], 1).reshape(-1).contiguous()
    out = torch.zeros((m, inter), dtype=torch.bfloat16, device=cuda)
    _run(lib, 3, a, w, bias, out, m, 2 * inter, k)
    ref = torch.nn.functional.silu(a.float() @ wg.float().t() + bg) * (a.float() @ wu.float().t() + bu)
    _close(out, ref, 6e-3)


def test_gemm_gelu(cuda, lib):
    m, n, k = 260, 512, 5120
    a, b, bias = _mk(m, n, k, cuda, seed=6, scale=0.3)
    out = torch.zeros((m, n), dtype=torch.bfloat16, device=cuda)
    _run(lib, 4, a, b, bias, out, m, n, k)
    _close(out, torch.nn.functional.gelu(a.float() @ b.float().t() + bias), 6e-3)


def test_gemm_scatter(cuda, lib):
    m, n, k = 200, 2048, 640
    a, b, bias = _mk(m, n, k, cuda, seed=7)
    perm = torch.randperm(m, generator=torch.Generator().manual_seed(1)).to(torch.int32).to(cuda)
    out = torch.zeros((m, n), dtype=torch.bfloat16, device=cuda)
    _run(lib, 5, a, b, bias, out, m, n, k, scatter=perm)
    ref = torch.empty((m, n), device=cuda)
    ref[perm.long()] = a.float() @ b.float().t() + bias
    _close(out, ref, 6e-3)


def test_gemm_qkv_rope(cuda, lib):
    """bias + 2D rotary on the q,k heads (HF apply_rotary_pos_emb_vision), v heads untouched."""
    m, heads, hd, k = 450, 16, 80, 1280
    n = 3 * heads * hd
    a, b, bias = _mk(m, n, k, cuda, seed=8)
    g = torch.Generator().manual_seed(9)
    pos = torch.randint(0, 50, (m, 2), generator=g, dtype=torch.int32).to(cuda)
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 40, 2, dtype=torch.float) / 40))
    ang = torch.arange(50, dtype=torch.float)[:, None] * inv_freq[None, :]
    rope = torch.stack([ang.cos(), ang.sin()], -1).contiguous().to(cuda)            # (50, 20, 2)
    out = torch.zeros((m, n), dtype=torch.bfloat16, device=cuda)
    _run(lib, 1, a, b, bias, out, m, n, k, pos=pos, rope=rope, heads=heads)
    y = (a.float() @ b.float().t() + bias).view(m, 3, heads, hd)
    rot = torch.cat([ang.to(cuda)[pos[:, 0].long()], ang.to(cuda)[pos[:, 1].long()]], -1)   # (m, 40)
    emb = torch.cat([rot, rot], -1)[:, None, :]
    def rope_fn(t):
        r = torch.cat([-t[..., 40:], t[..., :40]], -1)
        return t * emb.cos() + r * emb.sin()
    ref = torch.stack([rope_fn(y[:, 0]), rope_fn(y[:, 1]), y[:, 2]], 1).reshape(m, n)
    _close(out, ref, 6e-3)


def test_gemm_fp16_operands(cuda, lib):
    m, n, k = 300, 512, 1280
    g = torch.Generator().manual_seed(12)
    a = torch.randn(m, k, generator=g).to(torch.float16).to(cuda)
    b = (torch.randn(n, k, generator=g) * 0.05).to(torch.float16).to(cuda)
    bias = torch.randn(n, generator=g).to(cuda)
    out = torch.zeros((m, n), dtype=torch.float16, device=cuda)
    _run(lib, 0, a, b, bias, out, m, n, k)
    _close(out, a.float() @ b.float().t() + bias, 1.5e-3)
