"""Oracle: the image branch of the HF Qwen2-VL processor, PIL backend.  Test infrastructure only.

Follows HF ``models/qwen2_vl/image_processing_pil_qwen2_vl.py:143-224``
(``_preprocess``), ``image_transforms.py:89-124`` (rescale, f64 -> f32) and
``image_transforms.py:384-442`` (normalize in f32), with the resize replaced by
the NumPy restatement in ``oracle.resample``.
"""
import numpy as np

from . import geometry, resample

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
PATCH, MERGE, TEMPORAL = 14, 2, 2


def normalize_lut(mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD, rescale=1 / 255):
    """(3, 256) f32 table: f32(f64(v) * rescale) then (x - f32(mean)) / f32(std), every step rounded to f32."""
    v = (np.arange(256, dtype=np.float64) * rescale).astype(np.float32)
    m = np.array(mean, dtype=np.float32)
    s = np.array(std, dtype=np.float32)
    return ((v[None, :] - m[:, None]) / s[:, None]).astype(np.float32)


def patchify(img_f32_chw):
    """(3, H, W) f32 -> (gh*gw, 1176) rows in merge-group raster order, columns (c, t, py, px)."""
    c, h, w = img_f32_chw.shape
    gh, gw = h // PATCH, w // PATCH
    p = np.stack([img_f32_chw] * TEMPORAL, axis=0)[None]          # (1, 2, 3, H, W): frame repeated
    p = p.reshape(1, 1, TEMPORAL, c, gh // MERGE, MERGE, PATCH, gw // MERGE, MERGE, PATCH)
    p = p.transpose(0, 1, 4, 7, 5, 8, 3, 2, 6, 9)
    return np.ascontiguousarray(p.reshape(gh * gw, c * TEMPORAL * PATCH * PATCH)), (1, gh, gw)


def preprocess_u8(images_hwc, min_pixels=56 * 56, max_pixels=28 * 28 * 1280):
    """List of (H, W, 3) uint8 arrays -> (pixel_values (S,1176) f32, image_grid_thw (N,3) i64, resized u8 list)."""
    lut = normalize_lut()
    rows, grids, resized = [], [], []
    for img in images_hwc:
        h, w, _ = img.shape
        rh, rw = geometry.smart_resize(h, w, PATCH * MERGE, min_pixels, max_pixels)
        r = resample.resize_u8(img, rw, rh)
        resized.append(r)
        f = np.stack([lut[c][r[:, :, c]] for c in range(3)], axis=0)
        pv, g = patchify(f)
        rows.append(pv)
        grids.append(g)
    return np.concatenate(rows, 0), np.array(grids, dtype=np.int64), resized


def zoom_step_u8(image_hwc, bbox, min_size=512, min_pixels=56 * 56, max_pixels=128 * 128 * 28 * 28):
    """One zoom step of the fused path: cut_image box -> crop -> processor (single resample)."""
    h, w, _ = image_hwc.shape
    box = geometry.cut_box(w, h, bbox, min_size)
    crop = resample.crop_u8(image_hwc, box)
    pv, grid, _ = preprocess_u8([crop], min_pixels, max_pixels)
    return box, pv, grid
