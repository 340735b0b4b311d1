import sys, numpy as np, torch
sys.path.insert(0, ".")
from zoomearth_b200 import _lib
lib = _lib.lib()
dev = torch.device("cuda", 0)
heads, hd = 16, 80
def run(segs, dtype=torch.bfloat16, seed=0):
    S = sum(segs)
    g = torch.Generator().manual_seed(seed)
    qkv = torch.randn(S, 3, heads, hd, generator=g).to(dtype).to(dev)
    out = torch.full((S, heads * hd), float("nan"), dtype=dtype, device=dev)
    cu = np.concatenate([[0], np.cumsum(segs)]).astype(np.int32)
    work = torch.empty(16 * (S // 64 + len(segs) + 1) * 4 + 4096 + (S + 8) * heads * hd * 2, dtype=torch.uint8, device=dev)
    _lib.check(lib.zv_attention(qkv.data_ptr(), out.data_ptr(), heads, hd, cu.ctypes.data, len(segs), work.data_ptr(), work.numel(), 1, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    q, k, v = (t.float().transpose(0, 1) for t in qkv.unbind(1))
    refs = [torch.nn.functional.scaled_dot_product_attention(q[:, a:b], k[:, a:b], v[:, a:b]) for a, b in zip(cu[:-1], cu[1:])]
    ref = torch.cat(refs, 1).transpose(0, 1).reshape(S, heads * hd)
    for i, (a, b) in enumerate(zip(cu[:-1], cu[1:])):
        e = (out[a:b].float() - ref[a:b]).abs().max().item()
        print(f"  seg {i} [{a},{b}) max abs err {e:.4f} ref max {ref[a:b].abs().max().item():.3f}")
for segs in ([200, 200], [4, 200, 300], [12, 500], [100, 4900], [36, 1224, 100]):
    print(segs); run(segs)
