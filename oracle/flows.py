"""Oracle: the reference's cut_image / resize_image at its four call sites, on NumPy uint8 images.  Test infrastructure only.

Follows ``src/eval/infer.py:41-85`` (= ``src/demo.py:30-93`` with max_size 1024), ``src/train/SFT.py:76-125`` and
``src/train/RL/src/open-r1-multimodal/src/open_r1/custom/customized_funcs.py:37-85``; the Pillow calls inside them are the
restatements in ``oracle.resample``.  Pinned against tests/golden/flows.json (made by running the reference's own
functions, tests/golden/make_golden_flows.py).
"""
from . import geometry, resample


def apply_ops(img, ops):
    for kind, arg in ops:
        if kind == "crop":
            img = resample.crop_u8(img, arg)
        else:
            img = resample.resize_u8(img, arg[0], arg[1])
    return img


def cut_image(img, bbox, min_size=512, variant="infer"):
    h, w, _ = img.shape
    return apply_ops(img, geometry.cut_ops(w, h, bbox, min_size, variant))


def resize_image(img, max_size=512, variant="infer"):
    """-> (image, 1/scale)."""
    h, w, _ = img.shape
    nw, nh, inv = geometry.resize_dims_ex(w, h, max_size, variant)
    if (nw, nh) != (w, h) or variant == "sft":
        img = resample.resize_u8(img, nw, nh)
    return img, inv
