"""CPU: decode-once ImageStore - cache / prefetch / eviction logic with host tensors; pixels bit-identical to
the reference's own decode, Image.open(path).convert("RGB") (infer.py:215,237)."""
import os
import time

import numpy as np
import pytest
import torch
from PIL import Image


@pytest.fixture()
def files(tmp_path):
    rng = np.random.default_rng(5)
    out = {}
    for name, fmt, mode in (("a.tif", "TIFF", "RGB"), ("b.png", "PNG", "RGB"), ("c.jpg", "JPEG", "RGB"), ("d.png", "PNG", "L"),
                            ("e.png", "PNG", "RGBA")):
        shape = (60, 90) if mode == "L" else (60, 90, 4) if mode == "RGBA" else (60, 90, 3)
        p = str(tmp_path / name)
        Image.fromarray(rng.integers(0, 256, shape, dtype=np.uint8), mode=mode).save(p, format=fmt)
        out[name] = p
    return out


def test_decode_once_and_pixels_match_the_reference_decode(files):
    from zoomearth_b200.ingest import ImageStore
    store = ImageStore(device="cpu")
    for p in files.values():
        ref = np.asarray(Image.open(p).convert("RGB"))
        t = store.get(p)
        assert t.dtype == torch.uint8 and tuple(t.shape) == ref.shape and np.array_equal(t.numpy(), ref)
    n = store.decodes
    assert n == len(files)
    for p in files.values():                               # stage 2 of every question: served from the store
        store.get(p)
    assert store.decodes == n and store.hits == len(files)
    store.close()


def test_changed_file_is_decoded_again_and_drop(files):
    from zoomearth_b200.ingest import ImageStore
    store = ImageStore(device="cpu")
    p = files["b.png"]
    a = store.get(p).clone()
    time.sleep(0.01)
    Image.fromarray(np.full((10, 12, 3), 7, np.uint8)).save(p, format="PNG")
    b = store.get(p)
    assert store.decodes == 2 and tuple(b.shape) == (10, 12, 3) and not torch.equal(a[:10, :12], b)
    store.drop(p)
    assert p not in store and store.resident_bytes() == 0
    store.close()


def test_prefetch_overlaps_and_budget_evicts(files):
    from zoomearth_b200.ingest import ImageStore
    calls = []

    def slow_decoder(path):
        calls.append(path)
        time.sleep(0.2)
        return np.zeros((100, 100, 3), np.uint8)

    store = ImageStore(device="cpu", workers=4, decoder=slow_decoder, budget_bytes=2 * 100 * 100 * 3)
    paths = [files[k] for k in ("a.tif", "b.png", "c.jpg")]
    t0 = time.perf_counter()
    store.prefetch(paths)
    assert time.perf_counter() - t0 < 0.1                  # returns immediately
    for p in paths:
        store.get(p)
    assert time.perf_counter() - t0 < 0.45 and len(calls) == 3          # decoded concurrently, once each
    assert store.resident_bytes() <= 2 * 100 * 100 * 3 and paths[0] not in store and paths[2] in store   # LRU under the budget
    store.close()
