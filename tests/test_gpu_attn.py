"""Varlen attention parity (GPU, through the C ABI) vs torch SDPA per segment, fp32 reference on bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("segs", [[64], [64, 64, 32, 16, 4], [100], [4900 // 4 * 4], [1, 7, 129, 64, 200], [2048, 36],
                                  # window layers (every segment <= 64 rows: the tcgen05 window kernel): the bench shape's window
                                  # sizes, every multiple of 4, blocks that end with tiny windows, odd sizes, hundreds of windows
                                  [64] * 8 + [48] + [64] * 8 + [48] + [48] * 8 + [36], list(range(4, 68, 4)), [64, 60, 4, 4, 4, 64, 12],
                                  [4], [1, 7, 33, 64, 5, 63, 2], [64] * 300 + [16] * 7 + [48] * 41,
                                  [4] * 301 + [8] * 50 + [60, 64, 4] * 9])
def test_attention_vs_sdpa(cuda, lib, segs, dtype):
    import numpy as np
    from zoomearth_b200 import _lib
    heads, hd = 16, 80
    S = sum(segs)
    g = torch.Generator().manual_seed(S)
    qkv = torch.randn(S, 3, heads, hd, generator=g).to(dtype).to(cuda)
    out = torch.full((S, heads * hd), float("nan"), dtype=dtype, device=cuda)
    cu = np.concatenate([[0], np.cumsum(segs)]).astype(np.int32)
    work = torch.empty(16 * (S // 64 + len(segs) + 1) * 4 + 4096 + (S + 8) * heads * hd * 2, dtype=torch.uint8, device=cuda)
    _lib.check(lib.zv_attention(qkv.data_ptr(), out.data_ptr(), heads, hd, cu.ctypes.data, len(segs), work.data_ptr(),
                                work.numel(), 2 if dtype == torch.float16 else 1, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    q, k, v = (t.float().cpu().transpose(0, 1) for t in qkv.unbind(1))      # (heads, S, hd)
    refs = []
    for a, b in zip(cu[:-1], cu[1:]):
        refs.append(torch.nn.functional.scaled_dot_product_attention(q[:, a:b], k[:, a:b], v[:, a:b]))
    ref = torch.cat(refs, 1).transpose(0, 1).reshape(S, heads * hd)
    # relative bounds: outputs of a long segment are averages of thousands of values (magnitude ~ 0.03), so an absolute
    # bound has no teeth there.  The error budget is the 16-bit rounding of P and of the output: 2^-9 (bf16) / 2^-12 (fp16)
    # relative to the row's largest value.
    out = out.float().cpu()
    scale = ref.abs().max().item()
    err = (out - ref).abs().max().item() / scale
    tol = 6e-3 if dtype == torch.bfloat16 else 1e-3
    assert err < tol, f"max err / max|ref| = {err} (tolerance {tol})"
    rel_fro = ((out - ref).norm() / ref.norm()).item()
    assert rel_fro < (3.5e-3 if dtype == torch.bfloat16 else 5e-4), f"relative Frobenius error {rel_fro}"
    assert torch.nn.functional.cosine_similarity(out.flatten(), ref.flatten(), dim=0) > 0.9999
