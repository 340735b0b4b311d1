"""TEST INFRASTRUCTURE - CPU restatement of the LM hand-off, NumPy integer work.

    get_rope_index   reference src/train/RL/src/open-r1-multimodal/src/open_r1/model/modeling_qwen2_vl.py:967-1114
                     (transformers 5.x variant: HF models/qwen2_5_vl/modeling_qwen2_5_vl.py:1024-1135)
    masked_scatter   reference .../modeling_qwen2_vl.py:1191-1207; HF modeling_qwen2_5_vl.py:1179-1218,1301-1307

Pinned by tests/test_oracle_pinning.py against (a) fixtures produced by executing the REFERENCE's own
get_rope_index source (tests/golden/make_golden_handoff.py) and (b) the installed transformers 5.5 implementation.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np

IMAGE_TOKEN, VIDEO_TOKEN, VISION_START = 151655, 151656, 151652


def get_rope_index(input_ids, image_grid_thw, attention_mask=None, merge=2, image_token_id=IMAGE_TOKEN,
                   vision_start_token_id=VISION_START, hf5_semantics=False):
    """input_ids (B, L) int64; image_grid_thw (N, 3) or None -> (position_ids (3, B, L), deltas (B, 1))."""
    ids = np.asarray(input_ids, np.int64)
    B, L = ids.shape
    mask = np.ones_like(ids) if attention_mask is None else np.asarray(attention_mask, np.int64)
    if image_grid_thw is None:                                   # reference :1091-1112
        if attention_mask is None:
            pos = np.broadcast_to(np.arange(L, dtype=np.int64), (3, B, L)).copy()
            return pos, np.zeros((B, 1), np.int64)
        pos = np.cumsum(mask, -1) - 1
        pos[mask == 0] = 1
        pos = np.broadcast_to(pos, (3, B, L)).copy()
        return pos, pos.max(0).max(-1, keepdims=True) + 1 - L
    grid = np.asarray(image_grid_thw, np.int64).reshape(-1, 3)
    pos = np.full((3, B, L), 0 if hf5_semantics else 1, np.int64)
    deltas = np.zeros((B, 1), np.int64)
    image_index = 0
    for b in range(B):
        keep = mask[b] == 1
        tok = ids[b][keep].tolist()
        starts = [i for i, t in enumerate(tok) if t == vision_start_token_id]
        n_img = sum(1 for i in starts if tok[i + 1] == image_token_id)
        chunks, st = [], 0
        for _ in range(n_img):
            ed = tok.index(image_token_id, st)
            t, h, w = grid[image_index]
            image_index += 1
            gh, gw = int(h) // merge, int(w) // merge
            text_len = ed - st
            st_idx = int(chunks[-1].max()) + 1 if chunks else 0
            chunks.append(np.broadcast_to(np.arange(text_len), (3, text_len)) + st_idx)
            ti = np.repeat(np.arange(int(t)), gh * gw)
            hi = np.tile(np.repeat(np.arange(gh), gw), int(t))
            wi = np.tile(np.arange(gw), int(t) * gh)
            chunks.append(np.stack([ti, hi, wi]) + text_len + st_idx)
            st = ed + int(t) * gh * gw
        if st < len(tok):
            st_idx = int(chunks[-1].max()) + 1 if chunks else 0
            chunks.append(np.broadcast_to(np.arange(len(tok) - st), (3, len(tok) - st)) + st_idx)
        llm = np.concatenate(chunks, 1)
        pos[:, b, keep] = llm
        deltas[b, 0] = llm.max() + 1 - (len(tok) if hf5_semantics else L)
    return pos, deltas


def placeholder_rows(input_ids, image_token_id=IMAGE_TOKEN):
    return np.flatnonzero(np.asarray(input_ids).reshape(-1) == image_token_id).astype(np.int64)


def masked_scatter(inputs_embeds, input_ids, image_embeds, image_token_id=IMAGE_TOKEN):
    """inputs_embeds (B, L, D) with the k-th image-placeholder row replaced by image_embeds[k]."""
    out = np.array(inputs_embeds, copy=True)
    rows = placeholder_rows(input_ids, image_token_id)
    if len(rows) != len(image_embeds):
        raise ValueError(f"Image features and image tokens do not match: tokens: {len(rows)}, features {len(image_embeds)}")
    out.reshape(-1, out.shape[-1])[rows] = image_embeds
    return out
