"""install(model, processor): swap the two reference-facing objects for the fused ones, in place.

After this, the reference's own scripts run unchanged on top of the drop-in: ``processor(text=..., images=...)``
(``src/eval/infer.py:102-107``, ``src/demo.py:7-12``) lands on FusedImageProcessor, and the LM forward's
``self.visual(pixel_values, grid_thw=...)`` (HF modeling_qwen2_5_vl.py:1172) lands on FusedVisual.  The module
path is ``model.visual`` on transformers 4.49 (what the reference pins) and ``model.model.visual`` on 5.x.
"""
import torch

from .processor import FusedImageProcessor
from .visual import FusedVisual


def _find_visual(model):
    for owner in (model, getattr(model, "model", None)):
        if owner is not None and hasattr(owner, "visual"):
            return owner
    raise AttributeError("no `.visual` / `.model.visual` on this model")


def install(model=None, processor=None, device=None, dtype=None):
    """Returns (fused_visual or None, fused_image_processor or None)."""
    fv = fp = None
    if model is not None:
        owner = _find_visual(model)
        hf_visual = owner.visual
        import transformers
        returns_struct = int(transformers.__version__.split(".")[0]) >= 5
        p = next(hf_visual.parameters())
        fv = FusedVisual.from_hf(hf_visual, device=device or (p.device if p.device.type == "cuda" else None),
                                 dtype=dtype or (p.dtype if p.dtype in (torch.float32, torch.bfloat16, torch.float16) else torch.float16),
                                 return_pooling_output=returns_struct)
        owner.visual = fv
    if processor is not None:
        old = processor.image_processor
        fp = FusedImageProcessor(size=dict(getattr(old, "size", None) or {}) or None,
                                 patch_size=getattr(old, "patch_size", 14),
                                 temporal_patch_size=getattr(old, "temporal_patch_size", 2),
                                 merge_size=getattr(old, "merge_size", 2),
                                 image_mean=getattr(old, "image_mean", None), image_std=getattr(old, "image_std", None),
                                 device=device)
        processor.image_processor = fp
    return fv, fp
