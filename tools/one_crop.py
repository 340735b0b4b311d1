"""One 512-px zoom step, eager, a few times (for ncu launch lists of the batch-1 regime)."""
import sys
import torch
sys.path.insert(0, ".")
from zoomearth_b200 import FusedImageProcessor, FusedVisual, ZoomEncoder
from zoomearth_b200.synthetic import random_vision_state_dict
dev = torch.device("cuda", 0)
fv = FusedVisual(random_vision_state_dict(0, device=dev), device=dev, dtype=torch.float16)
enc = ZoomEncoder(fv, FusedImageProcessor(min_pixels=3136, max_pixels=128 * 128 * 28 * 28, device=dev))
img = torch.randint(0, 256, (5000, 5000, 3), dtype=torch.uint8, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    enc.encode([img], [(2000, 2000, 2512, 2512)], image_index=[0])
torch.cuda.synchronize()
