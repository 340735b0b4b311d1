"""CPU: the cut_image / resize_image variants of the reference's four call sites - oracle and C ABI - against
tests/golden/flows.json, which tests/golden/make_golden_flows.py produced by executing the reference's own functions
(src/eval/infer.py, src/train/SFT.py, open_r1/custom/customized_funcs.py) with real Pillow."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import flows as OF, geometry as OG

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def flows():
    return json.load(open(os.path.join(GOLD, "flows.json")))


def test_oracle_resize_dims_variants(flows):
    for c in flows["resize_image"]:
        w, h, inv = OG.resize_dims_ex(c["w"], c["h"], c["max_size"], c["variant"])
        assert [w, h] == c["size"], c
        if c["inv_scale"] is not None:
            assert inv == c["inv_scale"], c


def test_oracle_cut_image_variants(flows):
    for c in flows["cut_image"]:
        assert OG.cut_ops(c["w"], c["h"], c["bbox"], 512, c["variant"]) == c["ops"], c


def test_c_abi_variants_match_the_reference(flows, lib):
    from zoomearth_b200 import geometry as G
    for c in flows["resize_image"]:
        w, h, inv = G.resize_dims_ex(c["w"], c["h"], c["max_size"], c["variant"])
        assert [w, h] == c["size"], c
        if c["inv_scale"] is not None:
            assert inv == c["inv_scale"], c
    for c in flows["cut_image"]:
        assert G.cut_ops(c["w"], c["h"], c["bbox"], 512, c["variant"]) == c["ops"], c
    assert G.resize_dims_ex(5000, 5000, None, "demo")[:2] == (1024, 1024)
    with pytest.raises(ValueError):
        G.cut_box(100, 100, [1, 2, 3])
    with pytest.raises(ValueError):
        G.cut_box(100, 100, [1, float("nan"), 3, 4])
    with pytest.raises(OverflowError):
        G.cut_box(100, 100, [1, float("inf"), 3, 4])


def test_oracle_pixel_flows(flows):
    img = np.random.default_rng(303).integers(0, 256, (1500, 2100, 3), dtype=np.uint8)
    for c in flows["pixels"]:
        if c["flow"] == "resize_image":
            r, _ = OF.resize_image(img, c["max_size"], c["variant"])
        elif c["flow"] == "cut_image":
            r = OF.cut_image(img, c["bbox"], 512, c["variant"])
        else:
            r, _ = OF.resize_image(OF.cut_image(img, c["bbox"], 512, c["variant"]), c["max_size"], c["variant"])
        assert [r.shape[1], r.shape[0]] == c["size"], c
        assert hashlib.sha256(np.ascontiguousarray(r).tobytes()).hexdigest() == c["sha256"], c
