"""Oracle: zoom geometry (integer / fp64 host arithmetic).  Test infrastructure only.

Follows reference ``src/eval/infer.py``:
  * ``extract_bbox``  infer.py:20-32
  * ``cut_box``       infer.py:41-76   (the box arithmetic of ``cut_image``)
  * ``resize_dims``   infer.py:78-85   (the size arithmetic of ``resize_image``)
and HF ``models/qwen2_vl/image_processing_pil_qwen2_vl.py:57-83`` (``smart_resize``).
"""
import math
import re

_BBOX_RE = re.compile(r'"bbox_2d"\s*:\s*\[(.*?)\]', re.DOTALL)


def extract_bbox(text, scale):
    """infer.py:20-32 - all ``"bbox_2d": [...]`` lists, floats, times ``scale``."""
    out = []
    for m in _BBOX_RE.findall(text):
        try:
            nums = [float(x.strip()) for x in m.split(",")]
        except ValueError:
            continue
        out.append([n * scale for n in nums])
    return out


def cut_box(img_w, img_h, bbox, min_size=512):
    """infer.py:41-76 - the (left, upper, right, lower) box ``cut_image`` hands to ``Image.crop``."""
    x1, y1, x2, y2 = (int(v) for v in bbox)          # truncation toward zero
    width, height = x2 - x1, y2 - y1
    if width < min_size or height < min_size:
        cx = (x1 + x2) // 2
        cy = (y1 + y2) // 2
        nx1 = cx - min_size // 2
        ny1 = cy - min_size // 2
        nx2 = nx1 + min_size
        ny2 = ny1 + min_size
        if nx1 < 0:
            nx2 += -nx1
            nx1 = 0
        if ny1 < 0:
            ny2 += -ny1
            ny1 = 0
        if nx2 > img_w:
            nx1 -= nx2 - img_w
            nx2 = img_w
        if ny2 > img_h:
            ny1 -= ny2 - img_h
            ny2 = img_h
        nx1 = max(0, nx1)
        ny1 = max(0, ny1)
        nx2 = min(img_w, nx1 + min_size)
        ny2 = min(img_h, ny1 + min_size)
        return int(nx1), int(ny1), int(nx2), int(ny2)
    return x1, y1, x2, y2


def resize_dims(w, h, max_size):
    """infer.py:78-85 - (new_w, new_h, 1/scale); unchanged dims when scale >= 1."""
    scale = max_size / max(w, h)
    if scale < 1:
        return int(w * scale), int(h * scale), 1 / scale
    return w, h, 1 / scale


def smart_resize(height, width, factor=28, min_pixels=56 * 56, max_pixels=14 * 14 * 4 * 1280):
    """HF image_processing_pil_qwen2_vl.py:57-83 (Python round = half-to-even)."""
    if max(height, width) / min(height, width) > 200:
        raise ValueError(
            f"absolute aspect ratio must be smaller than 200, got {max(height, width) / min(height, width)}"
        )
    h_bar = round(height / factor) * factor
    w_bar = round(width / factor) * factor
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


# ---------------------------------------------------------------- the other call sites' variants
RESIZE_MODE = {"infer": 0, "demo": 0, "sft": 1, "custom": 2}


def resize_dims_ex(w, h, max_size, variant="infer"):
    """(new_w, new_h, 1/scale) of ``resize_image`` at each call site: ``infer`` / ``demo`` infer.py:78-85, demo.py:86-93
    (resize only when scale < 1); ``sft`` SFT.py:76-81 (always resizes, may upscale); ``custom`` customized_funcs.py:76-85
    (scale = max(30 / min(w, h), max_size / max(w, h)), resize only when < 1)."""
    scale = max_size / max(w, h)
    if variant == "custom":
        scale = max(30 / min(w, h), scale)
    if variant == "sft" or scale < 1:
        return int(w * scale), int(h * scale), 1 / scale
    return w, h, 1 / scale


def cut_ops(img_w, img_h, bbox, min_size=512, variant="infer"):
    """The Pillow operations ``cut_image`` performs at each call site, as a list of ["crop", box] / ["resize", (w, h)]:
    ``infer`` infer.py:41-76; ``custom`` customized_funcs.py:37-74 (the image itself when len(bbox) != 4); ``sft``
    SFT.py:83-125 (boxes with both sides >= min_size: crop, resize to min side = min_size, centre crop)."""
    if variant == "custom" and len(bbox) != 4:
        return []
    box = cut_box(img_w, img_h, bbox, min_size)
    ops = [["crop", list(box)]]
    if variant == "sft":
        x1, y1, x2, y2 = (int(v) for v in bbox)
        if not (x2 - x1 < min_size or y2 - y1 < min_size):
            w, h = box[2] - box[0], box[3] - box[1]
            scale = min_size / min(w, h)
            nw, nh = int(w * scale), int(h * scale)
            left, top = (nw - min_size) // 2, (nh - min_size) // 2
            ops += [["resize", [nw, nh]], ["crop", [left, top, left + min_size, top + min_size]]]
    return ops
