"""Oracle: the *installed* third-party code the reference executes, called live.  Test infrastructure only.

The reference's hot path is ``PIL.Image.crop/resize`` (``src/eval/infer.py:72-84``)
-> HF image processor (``infer.py:102-107``) -> ``model.visual`` (HF
``modeling_qwen2_5_vl.py:455-518``).  None of that lives under /root/reference;
Pillow and transformers are site-packages of this image (here and on the GPU
box), so they can be run as the reference arm.  ``cut_image``/``resize_image``
are taken from ``oracle.geometry`` (restated from infer.py:41-85; the reference
tree does not travel to the GPU box).
"""
import numpy as np
import torch

from . import geometry
from .tower import CFG


def pil_cut_image(image_hwc, bbox, min_size=512):
    """reference cut_image (infer.py:41-76) executed with real Pillow; returns (PIL.Image, box)."""
    from PIL import Image
    im = Image.fromarray(image_hwc)
    box = geometry.cut_box(im.width, im.height, bbox, min_size)
    return im.crop(box), box


def pil_resize_image(pil_image, max_size=512):
    """reference resize_image (infer.py:78-85) with real Pillow."""
    from PIL import Image
    w, h = pil_image.size
    nw, nh, inv = geometry.resize_dims(w, h, max_size)
    if (nw, nh) != (w, h):
        pil_image = pil_image.resize((nw, nh), Image.BICUBIC)
    return pil_image, inv


def pil_processor(min_pixels=56 * 56, max_pixels=28 * 28 * 1280):
    """HF PIL-backend image processor = the slow processor of transformers 4.49 the reference pins."""
    from transformers.models.qwen2_vl.image_processing_pil_qwen2_vl import Qwen2VLImageProcessorPil
    return Qwen2VLImageProcessorPil(min_pixels=min_pixels, max_pixels=max_pixels)


def hf_preprocess(images, min_pixels=56 * 56, max_pixels=28 * 28 * 1280):
    """images: list of PIL images or (H,W,3) u8 arrays -> (pixel_values f32 tensor, image_grid_thw i64 tensor)."""
    from PIL import Image
    ims = [Image.fromarray(i) if isinstance(i, np.ndarray) else i for i in images]
    out = pil_processor(min_pixels, max_pixels)(images=ims, return_tensors="pt")
    return out["pixel_values"], out["image_grid_thw"]


def hf_vision_config(cfg=CFG):
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLVisionConfig
    return Qwen2_5_VLVisionConfig(
        depth=cfg["depth"], hidden_size=cfg["hidden"], intermediate_size=cfg["inter"], num_heads=cfg["heads"],
        out_hidden_size=cfg["out_hidden"], patch_size=cfg["patch"], spatial_merge_size=cfg["merge"],
        temporal_patch_size=cfg["temporal"], window_size=cfg["window"],
        fullatt_block_indexes=list(cfg["fullatt"]), hidden_act="silu", in_channels=cfg["in_ch"])


def hf_tower(state_dict=None, cfg=CFG, dtype=torch.float32, seed=0):
    """Qwen2_5_VisionTransformerPretrainedModel (sdpa, eval); loads ``state_dict`` if given else seeded init."""
    from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VisionTransformerPretrainedModel
    vc = hf_vision_config(cfg)
    vc._attn_implementation = "sdpa"
    torch.manual_seed(seed)
    m = Qwen2_5_VisionTransformerPretrainedModel(vc).eval()
    if state_dict is not None:
        missing, unexpected = m.load_state_dict(state_dict, strict=False)
        assert not unexpected and all("inv_freq" in k for k in missing), (missing, unexpected)
    return m.to(dtype)


@torch.no_grad()
def hf_tower_forward(model, pixel_values, grid_thw):
    out = model(pixel_values.to(model.dtype), grid_thw=grid_thw)
    return out.pooler_output if hasattr(out, "pooler_output") else out
