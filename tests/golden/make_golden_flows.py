"""Generates tests/golden/flows.json from the reference's OWN cut_image / resize_image variants, in the build container.

    python tests/golden/make_golden_flows.py

The four call sites of the reference feed the processor different images:
  infer    src/eval/infer.py:41-85          cut_image(min 512) -> resize_image(512), returns (image, 1/scale)
  sft      src/train/SFT.py:76-125          resize_image(1024) ALWAYS resizes; cut_image's else-branch resizes the crop to
                                            min side 512 and centre-crops 512 x 512
  custom   src/train/RL/.../open_r1/custom/customized_funcs.py:37-85   cut_image returns the image when len(bbox) != 4;
                                            resize_image(512) with scale = max(30 / min(w, h), 512 / max(w, h))
(demo.py = infer's arithmetic with max_size 1024; pinned in geometry.json already.)  Those modules import accelerate /
trl / nltk, which are not installed here, so the functions are cut out of their source files with `ast` and executed
as-is (same trick as make_golden_handoff.py) - on stub images that record sizes (geometry cases) and on real Pillow images
(pixel cases: sha256 of the uint8 result).  The reference tree does not exist on the GPU box; this fixture travels.
"""
import ast
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"
FILES = {"infer": f"{REF}/eval/infer.py", "sft": f"{REF}/train/SFT.py",
         "custom": f"{REF}/train/RL/src/open-r1-multimodal/src/open_r1/custom/customized_funcs.py"}


def extract(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"Image": Image, "np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]


class Stub:
    """Records what the reference asks Pillow to do instead of doing it."""
    def __init__(self, w, h, ops=None):
        self.width, self.height, self.size, self.ops = w, h, (w, h), list(ops or [])

    def crop(self, box):
        box = tuple(int(round(v)) for v in box)
        return Stub(box[2] - box[0], box[3] - box[1], self.ops + [["crop", list(box)]])

    def resize(self, size, resample=None):
        return Stub(size[0], size[1], self.ops + [["resize", list(size)]])


def main():
    fn = {k: extract(p, ["resize_image", "cut_image"]) for k, p in FILES.items()}
    rng = np.random.default_rng(77)
    out = {"resize_image": [], "cut_image": [], "pixels": []}
    sizes = [(5000, 5000), (1300, 900), (512, 512), (511, 300), (1024, 1024), (1025, 40), (40, 3000), (60, 60), (29, 700)]
    sizes += [(int(rng.integers(20, 7000)), int(rng.integers(20, 7000))) for _ in range(120)]
    for w, h in sizes:
        for variant, ms in (("infer", 512), ("sft", 1024), ("custom", 512)):
            r = fn[variant][0](Stub(w, h), ms)
            img = r[0] if isinstance(r, tuple) else r
            out["resize_image"].append({"variant": variant, "w": w, "h": h, "max_size": ms, "size": list(img.size),
                                        "inv_scale": (r[1] if isinstance(r, tuple) else None)})
    boxes = [((5000, 5000), b) for b in [(1000, 1200, 2300, 2100), (100.7, 50.2, 300.9, 260.1), (4900, 4950, 4990, 4999),
                                         (-20, -30, 100, 90), (2000, 2000, 2600, 2300), (0, 0, 5000, 5000), (10, 10, 522, 522),
                                         (10, 10, 521, 900), (4800, 100, 5400, 900), (100, 100, 1636, 612)]]
    for _ in range(150):
        w, h = int(rng.integers(600, 6000)), int(rng.integers(600, 6000))
        x1, y1 = rng.uniform(-50, w - 100), rng.uniform(-50, h - 100)
        boxes.append(((w, h), (float(x1), float(y1), float(x1 + rng.uniform(1, 2500)), float(y1 + rng.uniform(1, 2500)))))
    for (w, h), b in boxes:
        for variant in ("sft", "custom"):
            r = fn[variant][1](Stub(w, h), b)
            out["cut_image"].append({"variant": variant, "w": w, "h": h, "bbox": list(b), "size": list(r.size), "ops": r.ops})
    r = fn["custom"][1](Stub(900, 700), [1, 2, 3])              # len(bbox) != 4 -> the image itself
    out["cut_image"].append({"variant": "custom", "w": 900, "h": 700, "bbox": [1, 2, 3], "size": list(r.size), "ops": r.ops})

    # pixel cases on a real image (regenerated from the seed by the tests)
    img = np.random.default_rng(303).integers(0, 256, (1500, 2100, 3), dtype=np.uint8)
    pil = Image.fromarray(img)
    sha = lambda im: hashlib.sha256(np.asarray(im).tobytes()).hexdigest()
    for variant, ms in (("infer", 512), ("sft", 1024), ("custom", 512), ("infer", 1024)):
        r = fn[variant][0](pil, ms)
        r = r[0] if isinstance(r, tuple) else r
        out["pixels"].append({"flow": "resize_image", "variant": variant, "max_size": ms, "size": list(r.size), "sha256": sha(r)})
    for b in [(100, 150, 1500, 1250), (300.5, 200.2, 700.9, 650.0), (1900, 1300, 2090, 1490), (-30, -20, 800, 900),
              (50, 60, 1100, 640), (1800, 100, 2400, 1200)]:
        for variant, ms in (("infer", 512), ("sft", 1024), ("custom", 512)):
            cut = fn[variant][1](pil, b)
            out["pixels"].append({"flow": "cut_image", "variant": variant, "bbox": list(b), "size": list(cut.size), "sha256": sha(cut)})
            r = fn[variant][0](cut, ms)
            r = r[0] if isinstance(r, tuple) else r
            out["pixels"].append({"flow": "resize_image(cut_image)", "variant": variant, "bbox": list(b), "max_size": ms,
                                  "size": list(r.size), "sha256": sha(r)})
    json.dump(out, open(os.path.join(HERE, "flows.json"), "w"))
    print({k: len(v) for k, v in out.items()}, "->", os.path.join(HERE, "flows.json"))


if __name__ == "__main__":
    main()
