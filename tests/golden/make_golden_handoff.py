"""Generates tests/golden/handoff_rope_index.npz by EXECUTING the reference's own get_rope_index
(/root/reference/src/train/RL/src/open-r1-multimodal/src/open_r1/model/modeling_qwen2_vl.py:967-1114): the function's
source is cut out of the reference file with `ast` and run as-is against a stand-in `self` that carries only the
config fields it reads.  (The module itself cannot be imported here: it needs accelerate/trl)
Run in the build container only; /root/reference does not exist on the GPU box.

    python tests/golden/make_golden_handoff.py
"""
import ast
import os
import textwrap
import types
from typing import Optional  # noqa: F401  (used by the executed source)

import numpy as np
import torch

REF = "/root/reference/src/train/RL/src/open-r1-multimodal/src/open_r1/model/modeling_qwen2_vl.py"
IMG, VID, VSTART, VEND, PAD = 151655, 151656, 151652, 151653, 151643


def reference_get_rope_index():
    src = open(REF).read()
    tree = ast.parse(src)
    best = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "get_rope_index":
            seg = ast.get_source_segment(src, node)
            if "vision_start_token_id" in seg and (best is None or len(seg) > len(best)):
                best = seg
    ns = {"torch": torch, "Optional": Optional}
    exec(textwrap.dedent(best), ns)
    return ns["get_rope_index"]


def sample(rng, n_images, lens, total, left_pad):
    """One row: [pad...] text <vs> <img>*T <ve> text ... ; returns (ids, mask, grids)."""
    ids, grids = [], []
    for k in range(n_images):
        ids += rng.integers(0, 1000, int(lens[k])).tolist()
        gh, gw = 2 * int(rng.integers(1, 9)), 2 * int(rng.integers(1, 9))
        grids.append((1, gh, gw))
        ids += [VSTART] + [IMG] * (gh * gw // 4) + [VEND]
    ids += rng.integers(0, 1000, int(lens[n_images])).tolist()
    return ids, grids


def main():
    fn = reference_get_rope_index()
    me = types.SimpleNamespace(config=types.SimpleNamespace(
        vision_config=types.SimpleNamespace(spatial_merge_size=2), image_token_id=IMG, video_token_id=VID,
        vision_start_token_id=VSTART))
    rng = np.random.default_rng(7)
    out = {}
    cases = [  # (images per row, padding side)
        ([1], "none"), ([2, 1], "left"), ([1, 2, 0], "right"), ([2, 2, 1, 1], "left")]
    for c, (imgs, side) in enumerate(cases):
        rows, grids = [], []
        for n in imgs:
            lens = rng.integers(0, 12, n + 1)
            if c == 1:
                lens[0] = 0                      # an image at position 0 and back-to-back images
            ids, g = sample(rng, n, lens, None, side)
            rows.append(ids)
            grids += g
        L = max(len(r) for r in rows)
        ids = np.full((len(rows), L), PAD, np.int64)
        mask = np.zeros((len(rows), L), np.int64)
        for i, r in enumerate(rows):
            sl = slice(L - len(r), L) if side == "left" else slice(0, len(r))
            ids[i, sl] = r
            mask[i, sl] = 1
        grid = np.asarray(grids, np.int64).reshape(-1, 3)
        am = None if side == "none" else torch.from_numpy(mask)
        pos, delta = fn(me, torch.from_numpy(ids), torch.from_numpy(grid) if len(grid) else None, None, am)
        out[f"ids{c}"], out[f"mask{c}"], out[f"grid{c}"] = ids, (mask if side != "none" else np.zeros((0, 0), np.int64)), grid
        out[f"pos{c}"], out[f"delta{c}"] = pos.numpy(), delta.numpy()
    # text-only branch, with and without a mask
    ids = rng.integers(0, 1000, (2, 9)).astype(np.int64)
    mask = np.array([[0, 0, 1, 1, 1, 1, 1, 1, 1], [1] * 9], np.int64)
    pos, delta = fn(me, torch.from_numpy(ids), None, None, torch.from_numpy(mask))
    out["ids_t"], out["mask_t"], out["pos_t"], out["delta_t"] = ids, mask, pos.numpy(), delta.numpy()
    pos, delta = fn(me, torch.from_numpy(ids), None, None, None)
    out["pos_t2"], out["delta_t2"] = pos.numpy(), delta.numpy()
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "handoff_rope_index.npz"), **out)
    print("wrote handoff_rope_index.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
