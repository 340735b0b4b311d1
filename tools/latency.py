"""Single-zoom-step latency (one 512x512 crop -> 504x504, 1296 patches, 324 tokens), eager launches vs CUDA-graph replay."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from zoomearth_b200 import _lib
if len(sys.argv) > 2 and sys.argv[1] == "--lib":            # build variants (experiments only)
    import os
    _lib.LIB_PATH = os.path.abspath(sys.argv[2])
from zoomearth_b200 import FusedImageProcessor, FusedVisual, ZoomEncoder
from zoomearth_b200.synthetic import random_vision_state_dict
dev = torch.device("cuda", 0)
fv = FusedVisual(random_vision_state_dict(0, device=dev), device=dev, dtype=torch.float16)
enc = ZoomEncoder(fv, FusedImageProcessor(min_pixels=3136, max_pixels=128 * 128 * 28 * 28, device=dev))
img = torch.randint(0, 256, (5000, 5000, 3), dtype=torch.uint8, device=dev)
res = {}
for name, boxes in {"one 512px crop": [(2000, 2000, 2512, 2512)], "four crops 512..1024px": [(100, 100, 612, 612), (900, 900, 1700, 1500), (2000, 100, 3024, 1124), (3000, 3000, 3700, 3600)]}.items():
    for graph in (False, True):
        for _ in range(5):
            emb, grid, _ = enc.encode([img], boxes, image_index=[0] * len(boxes), use_graph=graph)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 30
        for _ in range(n):
            emb, grid, _ = enc.encode([img], boxes, image_index=[0] * len(boxes), use_graph=graph)
        torch.cuda.synchronize()
        res[f"{name} | {'graph' if graph else 'eager'}"] = {"ms": (time.perf_counter() - t0) / n * 1e3, "tokens": int(emb.shape[0])}
print(json.dumps(res, indent=1))
