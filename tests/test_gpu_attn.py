"""Varlen attention parity (GPU, through the C ABI) vs torch SDPA per segment, fp32 reference on bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("segs", [[64], [64, 64, 32, 16, 4], [100], [4900 // 4 * 4], [1, 7, 129, 64, 200], [2048, 36]])
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
    q, k, v = (t.float().transpose(0, 1) for t in qkv.unbind(1))            # (heads, S, hd)
    refs = []
    for a, b in zip(cu[:-1], cu[1:]):
        refs.append(torch.nn.functional.scaled_dot_product_attention(q[:, a:b], k[:, a:b], v[:, a:b]))
    ref = torch.cat(refs, 1).transpose(0, 1).reshape(S, heads * hd)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, f"max err {err}"
    assert torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0) > 0.9999
