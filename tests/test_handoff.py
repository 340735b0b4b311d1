"""LM hand-off, host side: get_rope_index / placeholder rows.  The oracle (oracle/handoff.py) and the
C ABI (zv_rope_index, zv_placeholder_rows) against fixtures produced by executing the reference's own get_rope_index
(tests/golden/make_golden_handoff.py), against the installed transformers 5.x implementation, and against each other
on random batches.  Integer work: everything is compared with array_equal."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import handoff as OH

GOLD = os.path.join(os.path.dirname(__file__), "golden", "handoff_rope_index.npz")
IMG, VID, VSTART, VEND, PAD = 151655, 151656, 151652, 151653, 151643


def _cases():
    d = np.load(GOLD)
    for c in range(int(d["n_cases"])):
        mask = d[f"mask{c}"]
        yield d[f"ids{c}"], (None if mask.size == 0 else mask), d[f"grid{c}"], d[f"pos{c}"], d[f"delta{c}"]


def _random_batch(seed, left_pad):
    rng = np.random.default_rng(seed)
    rows, grids = [], []
    for _ in range(int(rng.integers(1, 5))):
        ids = []
        for _ in range(int(rng.integers(0, 4))):
            ids += rng.integers(0, 1000, int(rng.integers(0, 9))).tolist()
            gh, gw = 2 * int(rng.integers(1, 12)), 2 * int(rng.integers(1, 12))
            grids.append((1, gh, gw))
            ids += [VSTART] + [IMG] * (gh * gw // 4) + [VEND]
        ids += rng.integers(0, 1000, int(rng.integers(1, 9))).tolist()
        rows.append(ids)
    L = max(len(r) for r in rows)
    ids = np.full((len(rows), L), PAD, np.int64)
    mask = np.zeros((len(rows), L), np.int64)
    for i, r in enumerate(rows):
        sl = slice(L - len(r), L) if left_pad else slice(0, len(r))
        ids[i, sl], mask[i, sl] = r, 1
    return ids, mask, np.asarray(grids, np.int64).reshape(-1, 3)


def test_oracle_rope_index_matches_the_reference_fixtures():
    for ids, mask, grid, pos, delta in _cases():
        p, d = OH.get_rope_index(ids, grid, mask)
        assert np.array_equal(p, pos) and np.array_equal(d, delta)
    g = np.load(GOLD)
    p, d = OH.get_rope_index(g["ids_t"], None, g["mask_t"])
    assert np.array_equal(p, g["pos_t"]) and np.array_equal(d, g["delta_t"])
    p, d = OH.get_rope_index(g["ids_t"], None, None)
    assert np.array_equal(p, g["pos_t2"]) and np.array_equal(d, g["delta_t2"])


def test_cabi_rope_index_matches_the_reference_fixtures():
    from zoomearth_b200 import get_rope_index
    for ids, mask, grid, pos, delta in _cases():
        p, d = get_rope_index(torch.from_numpy(ids), torch.from_numpy(grid), None,
                              None if mask is None else torch.from_numpy(mask))
        assert p.dtype == torch.int64 and p.shape == pos.shape
        assert np.array_equal(p.numpy(), pos) and np.array_equal(d.numpy(), delta)
    g = np.load(GOLD)
    p, d = get_rope_index(torch.from_numpy(g["ids_t"]), None, None, torch.from_numpy(g["mask_t"]))
    assert np.array_equal(p.numpy(), g["pos_t"]) and np.array_equal(d.numpy(), g["delta_t"])
    p, d = get_rope_index(torch.from_numpy(g["ids_t"]), None, None, None)
    assert np.array_equal(p.numpy(), g["pos_t2"]) and np.array_equal(d.numpy(), g["delta_t2"])


@pytest.mark.parametrize("seed", range(12))
def test_cabi_rope_index_equals_oracle_on_random_batches(seed):
    from zoomearth_b200 import get_rope_index
    ids, mask, grid = _random_batch(seed, left_pad=seed % 2 == 0)
    for hf5 in (False, True):
        po, do = OH.get_rope_index(ids, grid if len(grid) else None, mask, hf5_semantics=hf5) if len(grid) else \
            OH.get_rope_index(ids, None, mask)
        p, d = get_rope_index(torch.from_numpy(ids), torch.from_numpy(grid) if len(grid) else None, None,
                              torch.from_numpy(mask), hf5_semantics=hf5)
        if not len(grid) and hf5:
            continue                     # transformers 5.x has no text-only branch (it returns None)
        assert np.array_equal(p.numpy(), po) and np.array_equal(d.numpy(), do)


@pytest.mark.parametrize("seed", range(4))
def test_hf5_semantics_match_installed_transformers(seed):
    """The 5.x variant (mm_token_type_ids based) on the same batches: positions 0 at pads, delta vs unpadded length.
    tokens_per_second is 1 here: 5.5 multiplies an image's temporal START position by tokens_per_second
    (HF modeling_qwen2_5_vl.py:1017-1018), which neither the reference's copy nor transformers 4.49 (the version the
    reference pins) does for images; that scaling is deliberately not reproduced."""
    mod = pytest.importorskip("transformers.models.qwen2_5_vl.modeling_qwen2_5_vl")
    from zoomearth_b200 import get_rope_index
    ids, mask, grid = _random_batch(100 + seed, left_pad=True)
    if not len(grid):
        pytest.skip("no image in this batch")
    model_cls = mod.Qwen2_5_VLModel
    me = types.SimpleNamespace(config=types.SimpleNamespace(
        vision_config=types.SimpleNamespace(spatial_merge_size=2, tokens_per_second=1)))
    me.get_vision_position_ids = types.MethodType(model_cls.get_vision_position_ids, me)
    tt = torch.from_numpy((ids == IMG).astype(np.int32))
    ref_pos, ref_delta = model_cls.get_rope_index(me, torch.from_numpy(ids), tt, torch.from_numpy(grid),
                                                  attention_mask=torch.from_numpy(mask))
    p, d = get_rope_index(torch.from_numpy(ids), torch.from_numpy(grid), None, torch.from_numpy(mask), hf5_semantics=True)
    assert torch.equal(p, ref_pos) and torch.equal(d, ref_delta.to(torch.int64))
    po, do = OH.get_rope_index(ids, grid, mask, hf5_semantics=True)
    assert np.array_equal(po, ref_pos.numpy()) and np.array_equal(do, ref_delta.numpy())


def test_placeholder_rows_and_masked_scatter_semantics():
    from zoomearth_b200 import placeholder_rows
    ids, mask, grid = _random_batch(3, left_pad=False)
    rows = placeholder_rows(torch.from_numpy(ids), IMG)
    assert np.array_equal(rows.numpy(), OH.placeholder_rows(ids, IMG))
    T = int((grid[:, 1] * grid[:, 2]).sum() // 4)
    assert rows.numel() == T
    emb = torch.randn(ids.shape[0], ids.shape[1], 8)
    feats = torch.randn(T, 8)
    ref = emb.masked_scatter((torch.from_numpy(ids) == IMG).unsqueeze(-1).expand_as(emb), feats)
    assert np.array_equal(OH.masked_scatter(emb.numpy(), ids, feats.numpy(), IMG), ref.numpy())
    with pytest.raises(ValueError, match="do not match"):
        placeholder_rows(torch.from_numpy(ids), IMG, expected=T + 1)


def test_rope_index_error_behaviour():
    from zoomearth_b200 import get_rope_index
    from zoomearth_b200._lib import ZoomVitError
    ids = torch.tensor([[5, VSTART, IMG, IMG, IMG, IMG, VEND, 7]])
    ok = get_rope_index(ids, torch.tensor([[1, 4, 4]]))[0]
    assert ok[:, 0].tolist() == [[0, 1, 2, 2, 2, 2, 4, 5], [0, 1, 2, 2, 3, 3, 4, 5], [0, 1, 2, 3, 2, 3, 4, 5]]
    with pytest.raises(ZoomVitError, match="does not match"):       # 4 placeholders, grid says 16 tokens
        get_rope_index(ids, torch.tensor([[1, 8, 8]]))
    with pytest.raises(ZoomVitError, match="video"):
        get_rope_index(torch.tensor([[VSTART, VID, VEND]]), torch.tensor([[1, 4, 4]]))
    with pytest.raises(NotImplementedError):
        get_rope_index(ids, torch.tensor([[1, 4, 4]]), video_grid_thw=torch.tensor([[1, 4, 4]]))
