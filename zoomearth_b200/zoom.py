"""The fused zoom step: crop box -> patches (bf16, window order) -> vision tower -> merged embeddings.

This is the fast path behind the reference's zoom loop (``src/eval/infer.py:236-247``: re-open image ->
``cut_image`` -> ``resize_image`` -> processor -> ``model.visual``): the decoded image is uploaded once and
stays resident on the GPU as uint8; every zoom step is two library calls (``zv_preprocess`` straight out of
the resident pixels, ``zv_visual_forward``) with no host round trip of pixel data.  Crops are independent, so
``encode`` takes any number of (image, box) pairs as one ragged batch.
"""
import numpy as np
import torch

from .processor import FusedImageProcessor
from .visual import FusedVisual


class ZoomEncoder:
    def __init__(self, visual: FusedVisual, processor: FusedImageProcessor = None, max_pixels=128 * 128 * 28 * 28,
                 min_pixels=56 * 56):
        self.visual = visual
        self.processor = processor or FusedImageProcessor(min_pixels=min_pixels, max_pixels=max_pixels,
                                                          device=visual.device)
        self.last_launches = 0

    def upload(self, image):
        """PIL / (H, W, 3) uint8 array or tensor -> resident uint8 CUDA tensor (one H2D copy per image)."""
        from .processor import _to_u8_hwc
        t = _to_u8_hwc(image)
        if t.device.type != "cuda":
            t = t.pin_memory().to(self.visual.device, non_blocking=True)
        return t

    @torch.no_grad()
    def encode(self, images_dev, boxes=None, image_index=None, apply_cut_image=True, return_patches=False):
        """images_dev: resident images; boxes (n, 4) in image pixels (None = global view of every image).
        Returns (embeddings (T, out_hidden), image_grid_thw (n, 3), crop boxes (n, 4))."""
        pv, grid, crop = self.processor.preprocess_crops(
            images_dev, boxes, out_dtype=self.visual.operand_dtype, window_order=True, image_index=image_index,
            apply_cut_image=apply_cut_image and boxes is not None)
        k1 = self.processor.last_launches
        emb = self.visual(pv, grid, window_order=True)
        self.last_launches = k1 + self.visual.last_launches
        if return_patches:
            return emb, grid, crop, pv
        return emb, grid, crop

    def tokens_per_crop(self, grid_thw):
        g = np.asarray(grid_thw)
        return (g[:, 0] * g[:, 1] * g[:, 2]) // (self.visual.spatial_merge_unit)
