"""Generates tests/golden/*.npz / *.json from the reference itself, in the build container.

    python tests/golden/make_golden.py

* geometry.json   - the reference's own ``cut_image`` / ``resize_image`` / ``extract_bbox`` (imported from
                    /root/reference/src/demo.py and executed as-is on stub images that record the box / size
                    they are asked for) plus HF ``smart_resize``, over seeded random cases and the edge cases
                    recorded in SURVEY.md 8c.
* pixels.npz      - real Pillow crops/resizes and the live HF PIL-backend processor on small seeded images:
                    resized uint8 images (compact) and sha256 of the fp32 ``pixel_values``.
* tower.npz       - live HF ``Qwen2_5_VisionTransformerPretrainedModel`` (2 blocks, one full-attention) outputs,
                    window_index and cu_window_seqlens for a ragged grid, with oracle.tower.make_weights(5).
The reference tree and these imports do not exist on the GPU box; the fixtures travel instead.
"""
import hashlib
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_reference_demo():
    spec = importlib.util.spec_from_file_location("zoomearth_demo", "/root/reference/src/demo.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class StubImage:
    """Records what the reference asks Pillow to do instead of doing it."""
    def __init__(self, w, h):
        self.width, self.height, self.size = w, h, (w, h)

    def crop(self, box):
        return tuple(int(v) for v in box)

    def resize(self, size, resample=None):
        return StubImage(*size)


def main():
    demo = load_reference_demo()
    from PIL import Image
    from transformers.models.qwen2_vl.image_processing_pil_qwen2_vl import smart_resize
    from oracle import hf_live, tower as OT
    rng = np.random.default_rng(2024)

    # ---------------------------------------------------------------- geometry
    geo = {"cut_image": [], "resize_image": [], "smart_resize": [], "extract_bbox": []}
    cases = [((5000, 5000), b) for b in [(1000, 1200, 2300, 2100), (100.7, 50.2, 300.9, 260.1), (4900, 4950, 4990, 4999),
                                         (-20, -30, 100, 90), (2000, 2000, 2600, 2300), (0, 0, 5000, 5000),
                                         (10, 10, 522, 522), (10, 10, 521, 900), (4800, 100, 5200, 900)]]
    for _ in range(200):
        w, h = int(rng.integers(300, 6000)), int(rng.integers(300, 6000))
        x1, y1 = rng.uniform(-50, w), rng.uniform(-50, h)
        x2, y2 = x1 + rng.uniform(1, 1500), y1 + rng.uniform(1, 1500)
        cases.append(((w, h), (float(x1), float(y1), float(x2), float(y2))))
    for (w, h), b in cases:
        geo["cut_image"].append({"w": w, "h": h, "bbox": list(b), "box": list(demo.cut_image(StubImage(w, h), b))})
    for _ in range(100):
        w, h = int(rng.integers(20, 7000)), int(rng.integers(20, 7000))
        for ms in (512, 1024):
            demo_out = demo.resize_image(StubImage(w, h), ms)
            geo["resize_image"].append({"w": w, "h": h, "max_size": ms, "size": list(demo_out.size)})
    sr = [(5000, 5000, 1003520), (5000, 5000, 12845056), (512, 512, 12845056), (900, 1300, 1003520), (518, 70, 12845056),
          (42, 42, 12845056), (30, 30, 12845056), (2048, 2048, 12845056), (3584, 3584, 12845056), (3585, 3584, 12845056)]
    for _ in range(300):
        sr.append((int(rng.integers(10, 6000)), int(rng.integers(10, 6000)), int(rng.choice([3136, 200704, 1003520, 12845056]))))
    for h, w, mx in sr:
        try:
            out = list(smart_resize(h, w, 28, 3136, mx))
        except ValueError as e:
            out = "ValueError"
        geo["smart_resize"].append({"h": h, "w": w, "max_pixels": mx, "out": out})
    texts = ['[{"bbox_2d": [10, 20, 300, 400], "label": "x"}]', 'a "bbox_2d" : [1.5,2.5, 3.5 ,4.5] b "bbox_2d":[7,8,9,10]',
             '"bbox_2d": [1, two, 3, 4]', 'no box here', '"bbox_2d": [\n 5,\n 6,\n 7,\n 8\n]']
    for t in texts:
        geo["extract_bbox"].append({"text": t, "scale": 4.8828125, "out_int": demo.extract_bbox(t, 4.8828125)})
    json.dump(geo, open(os.path.join(HERE, "geometry.json"), "w"))

    # ---------------------------------------------------------------- pixels (real Pillow + live HF processor)
    px = {}
    meta = []
    img = np.random.default_rng(101).integers(0, 256, (420, 560, 3), dtype=np.uint8)    # tests regenerate it from the seed
    pil = Image.fromarray(img)
    flows = [((30, 40, 330, 250), 3136, 50176), ((-15, -10, 200, 180), 3136, 12845056), ((0, 0, 560, 420), 3136, 28224),
             ((100, 100, 128, 128), 3136, 12845056), ((5, 300, 555, 360), 3136, 100352)]
    for i, (box, mn, mx) in enumerate(flows):
        crop = pil.crop(box)
        pv, grid = hf_live.hf_preprocess([crop], mn, mx)
        gh, gw = int(grid[0, 1]), int(grid[0, 2])
        resized = np.asarray(crop.resize((gw * 14, gh * 14), Image.BICUBIC)) if crop.size != (gw * 14, gh * 14) else np.asarray(crop)
        if resized.size <= 200000:
            px[f"resized_{i}"] = resized
        meta.append({"box": list(box), "min_pixels": mn, "max_pixels": mx, "grid": grid[0].tolist(),
                     "pv_sha256": hashlib.sha256(pv.numpy().tobytes()).hexdigest(), "pv_sum": float(pv.double().sum()),
                     "resized_sha256": hashlib.sha256(resized.tobytes()).hexdigest(),
                     "pv_first3": pv[0, :3].tolist()})
    # the reference-faithful two-resample flow of demo.py: cut_image -> resize_image(1024) -> processor
    big = np.random.default_rng(102).integers(0, 256, (1300, 1700, 3), dtype=np.uint8)
    two = demo.resize_image(demo.cut_image(Image.fromarray(big), (100, 150, 1500, 1250)))
    pv, grid = hf_live.hf_preprocess([two], 3136, 200704)
    meta.append({"flow": "demo two-step", "bbox": [100, 150, 1500, 1250], "two_step_size": list(two.size),
                 "two_step_sha256": hashlib.sha256(np.asarray(two).tobytes()).hexdigest(), "grid": grid[0].tolist(),
                 "pv_sha256": hashlib.sha256(pv.numpy().tobytes()).hexdigest()})
    px["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "pixels.npz"), **px)

    # ---------------------------------------------------------------- tower (live HF, 2 blocks)
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(5, cfg)
    grid = torch.tensor([[1, 6, 10], [1, 8, 8], [1, 2, 4]])
    S = int((grid[:, 1] * grid[:, 2]).sum())
    pv = torch.randn(S, 1176, generator=torch.Generator().manual_seed(11))
    model = hf_live.hf_tower(sd, cfg)
    out = hf_live.hf_tower_forward(model, pv, grid)
    widx, cu = model.get_window_index(grid)
    np.savez_compressed(os.path.join(HERE, "tower.npz"), grid=grid.numpy(), pixel_values=pv.numpy().astype(np.float32),
                        out=out.numpy(), window_index=widx.numpy(), cu_window_seqlens=np.array(cu, np.int64),
                        rot_pos_emb=model.rot_pos_emb(grid).numpy())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
