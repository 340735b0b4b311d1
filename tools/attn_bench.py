"""Micro-benchmark of the full-attention kernel (zv_attention) for build variants of libzoomvit.so.

    python tools/attn_bench.py [--lib path/to/variant.so]

Segments of 4900 patches (the bench shape, 8 images) and one 65 536-patch segment, fp16; CUDA events, TFLOP/s on the
algorithmic 4 S^2 hidden FLOPs; accuracy against fp32 SDPA on a 2 048-patch segment."""
import argparse, json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser(); ap.add_argument("--lib", default=None); args = ap.parse_args()
from zoomearth_b200 import _lib
if args.lib:
    _lib.LIB_PATH = os.path.abspath(args.lib)
lib = _lib.lib()
dev = torch.device("cuda", 0)
heads, hd = 16, 80
def run(segs, reps):
    S = sum(segs)
    qkv = torch.randn(S, 3, heads, hd, device=dev).half()
    out = torch.empty(S, heads * hd, dtype=torch.float16, device=dev)
    cu = np.concatenate([[0], np.cumsum(segs)]).astype(np.int32)
    work = torch.empty(16 * (S // 64 + len(segs) + 1) * 4 + 8192 + (S + 256) * 8, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: _lib.check(lib.zv_attention(qkv.data_ptr(), out.data_ptr(), heads, hd, cu.ctypes.data, len(segs), work.data_ptr(), work.numel(), 2, st))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = sum(4.0 * s * s * heads * hd for s in segs)
    return ms, fl / ms / 1e9, qkv, out, cu
res = {}
# window layers of the bench shape: per image 8 rows of (8 x 64 + 48) and one row of (8 x 48 + 36); HBM-bound: GB/s on 10 240 B per patch
win = ([64] * 8 + [48]) * 8 + [48] * 8 + [36]
for n_img in (8, 64):
    ms, tf, *_ = run(win * n_img, 20)
    res[f"windows, {n_img} x 70x70"] = {"ms": ms, "gbps": sum(win) * n_img * 10240 / ms / 1e6}
ms, tf, *_ = run([4900] * 8, 20); res["8 x 4900"] = {"ms": ms, "tflops": tf}
ms, tf, *_ = run([4900] * 64, 5); res["64 x 4900"] = {"ms": ms, "tflops": tf}
ms, tf, *_ = run([65536], 5); res["1 x 65536"] = {"ms": ms, "tflops": tf}
ms, tf, qkv, out, cu = run([64, 48, 36, 64, 4, 60], 2)
q, k, v = (t.float().transpose(0, 1) for t in qkv.unbind(1))
ref = torch.cat([torch.nn.functional.scaled_dot_product_attention(q[:, a:b], k[:, a:b], v[:, a:b]) for a, b in zip(cu[:-1], cu[1:])], 1).transpose(0, 1).reshape(-1, heads * hd)
o = out.float()
res["window accuracy vs fp32 sdpa"] = {"max_err_over_max": ((o - ref).abs().max() / ref.abs().max()).item(), "rel_fro": ((o - ref).norm() / ref.norm()).item()}
ms, tf, qkv, out, cu = run([2048, 777], 2)
q, k, v = (t.float().transpose(0, 1) for t in qkv.unbind(1))
ref = torch.cat([torch.nn.functional.scaled_dot_product_attention(q[:, a:b], k[:, a:b], v[:, a:b]) for a, b in zip(cu[:-1], cu[1:])], 1).transpose(0, 1).reshape(-1, heads * hd)
o = out.float()
res["accuracy vs fp32 sdpa"] = {"max_err_over_max": ((o - ref).abs().max() / ref.abs().max()).item(), "rel_fro": ((o - ref).norm() / ref.norm()).item()}
print(json.dumps({"lib": args.lib or "default", **res}))
