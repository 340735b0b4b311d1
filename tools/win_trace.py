"""Debug: per-item clock stamps of CTA 0 of the window-attention kernel (variant built with -DZV_WIN_TRACE)."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zoomearth_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
lib = _lib.lib()
dev = torch.device("cuda", 0)
heads, hd = 16, 80
win = (([64] * 8 + [48]) * 8 + [48] * 8 + [36]) * 64
S = sum(win)
qkv = torch.randn(S, 3, heads, hd, device=dev).half()
out = torch.empty(S, heads * hd, dtype=torch.float16, device=dev)
cu = np.concatenate([[0], np.cumsum(win)]).astype(np.int32)
work = torch.empty(16 * (S // 64 + len(win) + 1) * 4 + 8192 + (S + 256) * 8, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.zv_attention(qkv.data_ptr(), out.data_ptr(), heads, hd, cu.ctypes.data, len(win), work.data_ptr(), work.numel(), 2, st))
torch.cuda.synchronize()
h = C.CDLL(_lib.LIB_PATH)
buf = np.zeros(16 * 64, np.int64)
h.zv_debug_win_trace(buf.ctypes.data_as(C.c_void_p))
t = buf.reshape(16, 64)
t0 = t[0, 0]
names = ["load issue", "O loaded", "ep pre-wait", "ep post-wait", "ep fenced", "-", "sm reach", "sm S ready", "sm S loaded", "sm P out", "ep reach", "ep O ready", "ep stored", "sm max", "sm exps", "sm P issued"]
print("item " + " ".join(f"{n:>15s}" for n in names))
for k in range(3, 16):
    print(f"{k:4d} " + " ".join(f"{int(t[i, k] - t0):15d}" for i in range(16)))
print("period (load issue):", np.diff(t[0, 3:30]).tolist())
