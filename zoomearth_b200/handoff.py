"""LM hand-off: what sits between ``model.visual`` and the language model in the reference.

  get_rope_index    reference src/train/RL/src/open-r1-multimodal/src/open_r1/model/modeling_qwen2_vl.py:967-1114
                    (same signature; by ``hf5_semantics`` the transformers 5.x variant, HF modeling_qwen2_5_vl.py:1024-1135)
  embed_images      reference .../modeling_qwen2_vl.py:1191-1207 / HF modeling_qwen2_5_vl.py:1301-1307:
                    ``inputs_embeds.masked_scatter(image_mask, visual(pixel_values, grid_thw))`` - here the tower's last
                    GEMM writes every embedding row straight into its placeholder row of ``inputs_embeds``
                    (``zv_visual_forward_into``), so the (T, 2048) tensor never makes a round trip through HBM.

The position index is host integer work in libzoomvit (``zv_rope_index``); both are thin ctypes calls.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

IMAGE_TOKEN_ID, VIDEO_TOKEN_ID, VISION_START_TOKEN_ID = 151655, 151656, 151652      # Qwen2.5-VL tokenizer


def get_rope_index(input_ids, image_grid_thw=None, video_grid_thw=None, attention_mask=None, *, spatial_merge_size=2,
                   image_token_id=IMAGE_TOKEN_ID, video_token_id=VIDEO_TOKEN_ID,
                   vision_start_token_id=VISION_START_TOKEN_ID, hf5_semantics=False):
    """-> (position_ids (3, B, L) int64, mrope_position_deltas (B, 1) int64), on ``input_ids``' device."""
    if video_grid_thw is not None:
        raise NotImplementedError("get_rope_index covers images only (the reference's zoom loop feeds no video)")
    dev = input_ids.device
    ids = np.ascontiguousarray(input_ids.detach().cpu().numpy().astype(np.int64))
    if ids.ndim != 2:
        raise ValueError("input_ids must be (batch, seq_len)")
    B, L = ids.shape
    mask = None if attention_mask is None else np.ascontiguousarray(attention_mask.detach().cpu().numpy().astype(np.int64))
    grid = None if image_grid_thw is None else np.ascontiguousarray(
        np.asarray(image_grid_thw.detach().cpu() if isinstance(image_grid_thw, torch.Tensor) else image_grid_thw,
                   dtype=np.int64).reshape(-1, 3))
    pos = np.empty((3, B, L), np.int64)
    delta = np.empty((B, 1), np.int64)
    _lib.check(_lib.lib().zv_rope_index(
        ids.ctypes.data, None if mask is None else mask.ctypes.data, B, L, None if grid is None else grid.ctypes.data,
        0 if grid is None else grid.shape[0], image_token_id, video_token_id, vision_start_token_id, spatial_merge_size,
        1 if hf5_semantics else 0, pos.ctypes.data, delta.ctypes.data))
    return torch.from_numpy(pos).to(dev), torch.from_numpy(delta).to(dev)


def placeholder_rows(input_ids, image_token_id=IMAGE_TOKEN_ID, expected=None):
    """Flattened row indices of the image placeholders (int64 tensor on ``input_ids``' device), in masked_scatter
    order.  ``expected`` = the number of embeddings: a mismatch raises the ValueError HF raises (one host sync)."""
    flat = input_ids.reshape(-1)
    if flat.device.type == "cpu":
        ids = np.ascontiguousarray(flat.numpy().astype(np.int64))
        n = _lib.check(_lib.lib().zv_placeholder_rows(ids.ctypes.data, ids.size, image_token_id, None, 0))
        rows = np.empty(n, np.int64)
        _lib.check(_lib.lib().zv_placeholder_rows(ids.ctypes.data, ids.size, image_token_id, rows.ctypes.data, n))
        rows = torch.from_numpy(rows)
    else:
        rows = torch.nonzero(flat == image_token_id).reshape(-1)
    if expected is not None and rows.numel() != expected:
        raise ValueError(f"Image features and image tokens do not match, tokens: {rows.numel()}, features: {expected}")
    return rows


@torch.no_grad()
def embed_images(visual, inputs_embeds, input_ids, pixel_values, image_grid_thw, image_token_id=IMAGE_TOKEN_ID,
                 window_order=False):
    """In place: ``inputs_embeds`` (B, L, D) gets the tower's embeddings at its image-placeholder rows.  ``visual`` is
    a ``FusedVisual``; ``pixel_values`` as for ``visual.forward`` (HF-order fp32/16-bit patches, or the fused
    preprocess output with ``window_order=True``).  Returns ``inputs_embeds``."""
    plan = visual.plan_for(image_grid_thw)
    rows = placeholder_rows(input_ids, image_token_id, expected=plan.num_tokens).to(visual.device)
    return visual.forward_into(inputs_embeds, rows, pixel_values, image_grid_thw, window_order=window_order)
