"""One 512-px zoom step (eager launches): per-kernel-class device time from the library's event timing."""
import sys, json, ctypes as C
import torch
sys.path.insert(0, ".")
from zoomearth_b200 import FusedImageProcessor, FusedVisual, ZoomEncoder, _lib
from zoomearth_b200.synthetic import random_vision_state_dict
dev = torch.device("cuda", 0)
fv = FusedVisual(random_vision_state_dict(0, device=dev), device=dev, dtype=torch.float16)
enc = ZoomEncoder(fv, FusedImageProcessor(min_pixels=3136, max_pixels=128 * 128 * 28 * 28, device=dev))
img = torch.randint(0, 256, (5000, 5000, 3), dtype=torch.uint8, device=dev)
lib = _lib.lib()
K = {"k1_hpass": 0, "k1_vpass": 1, "gemm_store": 2, "gemm_qkv": 3, "gemm_resid": 4, "gemm_swiglu": 5, "gemm_gelu": 6, "gemm_scatter": 7, "attn_window": 8, "attn_full": 9, "rmsnorm": 10, "gather": 11}
for _ in range(5): enc.encode([img], [(2000, 2000, 2512, 2512)], image_index=[0])
torch.cuda.synchronize()
lib.zv_timing_reset(); lib.zv_timing_enable(1)
n = 20
for _ in range(n): enc.encode([img], [(2000, 2000, 2512, 2512)], image_index=[0])
torch.cuda.synchronize(); lib.zv_timing_enable(0)
out = {}
for k, c in K.items():
    t, m = C.c_double(), C.c_int64()
    lib.zv_timing_read(c, C.byref(t), C.byref(m))
    out[k] = {"us_per_launch": round(t.value / max(1, m.value) * 1e3, 2), "launches_per_step": m.value // n, "ms_per_step": round(t.value / n, 4)}
out["sum_ms"] = round(sum(v["ms_per_step"] for v in out.values()), 3)
print(json.dumps(out, indent=1))
