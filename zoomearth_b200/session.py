"""ZoomSession - the zoom loop's vision side with resident images and a global-view embedding cache.

The reference's loop re-opens and re-decodes the image for stage 2 (``src/eval/infer.py:215,237``) and re-encodes the
global view inside every stage-2 call (``infer.py:242-247`` passes ``[images[i], image_bbox]``; the GRPO rollout
encodes it up to three times, ``grpo_trainer.py:607,651-656``).  Here each image is uploaded once, its global-view
embeddings are computed once and cached, and the stage-2 crops of many questions are encoded as ONE ragged batch.
The LM side is untouched: ``stage1`` / ``stage2`` return exactly the per-image embedding blocks and ``image_grid_thw``
rows that ``get_image_features`` would have produced, in the order the reference feeds them.
"""
import numpy as np
import torch

from .geometry import cut_box, resize_dims_ex
from .zoom import ZoomEncoder


class ZoomSession:
    def __init__(self, encoder: ZoomEncoder, global_max_size=None, variant="infer", store=None):
        """global_max_size: if set (512 in infer.py, 1024 in demo.py), every image the tower sees goes through the
        reference's ``resize_image`` first - the global view (infer.py:215) AND each zoom crop
        (``resize_image(cut_image(...))``, infer.py:239) - as a device-side uint8 resample (``zv_resize_u8``), so grids,
        token counts and pixels are those of the unmodified loop and ``scale()`` is the factor that maps boxes the model
        draws on the resized view back to source pixels.  None = the fused single-resample fast path, scale 1.
        ``variant``: which call site's ``resize_image`` / ``cut_image`` ("infer", "demo", "sft", "custom").
        ``store``: an ``ingest.ImageStore`` for ``add_image(key, path)`` (created on first use otherwise)."""
        self.enc = encoder
        self.store = store             # optional ingest.ImageStore: add_image(key, path) decodes each file once
        self.global_max_size = global_max_size
        self.variant = variant
        self._pre = None if global_max_size is None else (variant, int(global_max_size))
        self._images = {}          # key -> resident uint8 tensor
        self._global = {}          # key -> (embeddings (T, D), grid_thw row)

    def add_image(self, key, image):
        """image: PIL / uint8 array / tensor, or a file path (decoded once through ``store``, infer.py:215,237)."""
        if key not in self._images:
            if isinstance(image, (str, bytes)) or hasattr(image, "__fspath__"):
                if self.store is None:
                    from .ingest import ImageStore
                    self.store = ImageStore(device=self.enc.visual.device)
                self._images[key] = self.store.get(image)
            else:
                self._images[key] = self.enc.upload(image)
        return self._images[key]

    def drop(self, key):
        self._images.pop(key, None)
        self._global.pop(key, None)

    def scale(self, key):
        """The factor the reference multiplies model boxes by (``extract_bbox(text, scale)``, infer.py:226)."""
        t = self._images[key]
        if self.global_max_size is None:
            return 1.0
        return resize_dims_ex(int(t.shape[1]), int(t.shape[0]), self.global_max_size, self.variant)[2]

    def stage1(self, keys):
        """Global-view embeddings for ``keys`` (cached).  Returns (list of (T_i, D) tensors, grid_thw (n, 3))."""
        missing = [k for k in keys if k not in self._global]
        if missing:
            emb, grid, _ = self.enc.encode([self._images[k] for k in missing], None, pre_resize=self._pre)
            tokens = self.enc.tokens_per_crop(grid.numpy())
            off = 0
            for k, n, g in zip(missing, tokens, grid):
                self._global[k] = (emb[off:off + int(n)], g.clone())
                off += int(n)
        return [self._global[k][0] for k in keys], torch.stack([self._global[k][1] for k in keys])

    def stage2(self, keys, bboxes):
        """One ragged batch of zoom crops (``cut_image`` rule applied) for many questions.  Returns, per question,
        ([global embeddings, crop embeddings], grid_thw (2, 3)) - the two images infer.py:242-247 hands the model -
        plus the crop boxes used."""
        g_emb, g_grid = self.stage1(keys)
        uniq = {k: i for i, k in enumerate(dict.fromkeys(keys))}
        imgs = [self._images[k] for k in uniq]
        emb, grid, crop = self.enc.encode(imgs, np.asarray(bboxes, np.float64), image_index=[uniq[k] for k in keys],
                                          pre_resize=self._pre)
        tokens = self.enc.tokens_per_crop(grid.numpy())
        out, off = [], 0
        for i, n in enumerate(tokens):
            out.append(([g_emb[i], emb[off:off + int(n)]], torch.stack([g_grid[i], grid[i]])))
            off += int(n)
        return out, crop

    def crop_box(self, key, bbox, min_size=512):
        t = self._images[key]
        return cut_box(int(t.shape[1]), int(t.shape[0]), bbox, min_size)
