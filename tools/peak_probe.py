"""cuBLAS dense GEMM, sustained (seconds-long loop under the power cap): bf16 vs fp16 operands, N(0,1) vs small-magnitude
data.  Explains how much of a bench difference between operand dtypes is the power cap (bit activity), not the kernel."""
import json, subprocess, sys, time
import torch
dev = torch.device("cuda", 0)
N = 8192
def clocks():
    r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
    return r.stdout.strip()
out = {}
for name, dt, scale in (("bf16 randn", torch.bfloat16, 1.0), ("fp16 randn", torch.float16, 1.0), ("bf16 randn*0.02", torch.bfloat16, 0.02),
                        ("fp16 randn*0.02", torch.float16, 0.02), ("bf16 randn", torch.bfloat16, 1.0)):
    a = (torch.randn(N, N, device=dev) * scale).to(dt); b = (torch.randn(N, N, device=dev) * scale).to(dt)
    for _ in range(5): a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); n = 0; e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(20): a @ b
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    c = clocks()
    out[name + (" (2nd)" if name in out else "")] = {"tflops": 2 * N**3 * n / (e0.elapsed_time(e1) / 1e3) / 1e12, "sm_mhz,power_w": c}
print(json.dumps(out, indent=1))
