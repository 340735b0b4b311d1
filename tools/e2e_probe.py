#!/usr/bin/env python
"""Where the end-to-end (host pixels in, embeddings out) leg of bench.py loses against the resident-image leg:
times encode_host under several chunk schedules, the same chunking without the H2D copy, and the copy alone."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zoomearth_b200 import FusedImageProcessor, FusedVisual, ZoomEncoder
from zoomearth_b200.synthetic import random_vision_state_dict

dev = torch.device("cuda", 0)
visual = FusedVisual(random_vision_state_dict(0, device=dev), device=dev, dtype=torch.float16)
enc = ZoomEncoder(visual, FusedImageProcessor(min_pixels=3136, max_pixels=1280 * 28 * 28, device=dev))
g = torch.Generator(device=dev).manual_seed(1)
images = [torch.randint(0, 256, (5000, 5000, 3), generator=g, dtype=torch.uint8, device=dev) for _ in range(64)]
host = [im.cpu().pin_memory() for im in images]
out_host = torch.empty((64 * 1225, 2048), dtype=torch.float16).pin_memory()


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


res = {}
res["resident_one_batch_ms"] = timed(lambda: enc.encode(images, None))
for c in (4, 8, 16):
    def chunks_resident():
        for i in range(0, 64, c):
            enc.encode(images[i:i + c], None)
    res[f"resident_chunks_of_{c}_ms"] = timed(chunks_resident)
def h2d():
    t = [h.to(dev, non_blocking=True) for h in host]
res["h2d_only_ms"] = timed(h2d)
for sched in ("8", "16", "ramp"):
    if sched == "ramp":
        fn = lambda: enc.encode_host(host, out_host=out_host, schedule=[2, 6, 8, 16, 32])
    else:
        fn = lambda: enc.encode_host(host, chunk=int(sched), out_host=out_host)
    res[f"encode_host_{sched}_ms"] = timed(fn)
print(json.dumps(res))
