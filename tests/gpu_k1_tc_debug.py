"""(not a pytest file; lives under tests/ because it uses the oracle as its checker)
GPU diagnosis of K1's tensor-core route: a few crops through zv_resize_u8 / zv_preprocess against the oracle, with the
mismatches broken down by position so that a layout error (swizzle, lane mapping, digit order) shows its signature.

    python tests/gpu_k1_tc_debug.py [--bench N]    # --bench: also time zv_preprocess alone on N 5000x5000 images
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
from oracle import processor as OP, resample as OR   # noqa: E402  (checker only)
from zoomearth_b200 import FusedImageProcessor, _lib   # noqa: E402


def img(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def report(name, got, ref):
    bad = np.argwhere(got != ref)
    print(f"[{name}] shape {got.shape} mismatches {len(bad)} of {got.size}")
    if len(bad) == 0:
        return True
    y, xb = bad[:, 0], bad[:, 1] * (got.shape[2] if got.ndim == 3 else 1) + (bad[:, 2] if got.ndim == 3 else 0)
    print("   first:", bad[:6].tolist())
    g, r = got[tuple(bad[:6].T)], ref[tuple(bad[:6].T)]
    print("   got", g.tolist(), "ref", r.tolist())
    for label, v, m in (("y%32", y, 32), ("xbyte%32", xb, 32), ("xbyte%4", xb, 4), ("y//32", y, None), ("xbyte//128", xb // 128, None)):
        vals = v % m if m else v
        u, c = np.unique(vals, return_counts=True)
        print(f"   by {label}:", dict(zip(u.tolist()[:40], c.tolist()[:40])))
    return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bench", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    fp = FusedImageProcessor(min_pixels=3136, max_pixels=12845056, device=dev)
    ok = True
    cases = [(64, 64, (0, 0, 64, 64), (56, 56)), (512, 512, (0, 0, 512, 512), (504, 504)), (600, 800, (0, 0, 800, 600), (420, 560)),
             (900, 1300, (100, 200, 700, 648), (308, 420)), (1200, 1600, (3, 5, 1599, 1197), (224, 308)),
             (2000, 2000, (0, 0, 2000, 2000), (196, 196))]
    for h, w, box, (oh, ow) in cases:
        a = img(h + w, h, w)
        t = torch.from_numpy(a).to(dev)
        try:
            out = fp.resize_u8([t], [box], [(ow, oh)])[0]
            torch.cuda.synchronize()
        except Exception as e:       # noqa: BLE001
            print(f"[resize {h}x{w}->{oh}x{ow}] FAILED: {e}")
            return 1
        ref = OR.resize_u8(OR.crop_u8(a, box), ow, oh)
        ok &= report(f"resize {h}x{w} box {box} -> {oh}x{ow}", out.cpu().numpy(), ref)
    # patches, config-2 shape
    a = img(0, 5000, 5000)
    t = torch.from_numpy(a).to(dev)
    fp2 = FusedImageProcessor(min_pixels=3136, max_pixels=1280 * 28 * 28, device=dev)
    pv, grid, _ = fp2.preprocess_crops([t], None, torch.float32)
    torch.cuda.synchronize()
    from PIL import Image
    r = np.asarray(Image.fromarray(a).resize((980, 980), Image.BICUBIC))
    lut = OP.normalize_lut()
    ref, _ = OP.patchify(np.stack([lut[c][r[:, :, c]] for c in range(3)], 0))
    same = np.array_equal(pv.cpu().numpy(), ref)
    print("[patches 5000->980] equal:", same, "launches", fp2.last_launches)
    ok &= same
    if args.bench:
        lib = _lib.lib()
        images = [t] + [torch.from_numpy(img(i + 1, 5000, 5000)).to(dev) for i in range(min(args.bench, 8) - 1)]
        images = [images[i % len(images)] for i in range(args.bench)]
        out = torch.empty((args.bench * 4900, 1176), dtype=torch.float16, device=dev)
        for _ in range(2):
            fp2.preprocess_crops(images, None, torch.float16, window_order=True, out=out)
        torch.cuda.synchronize()
        lib.zv_timing_reset(); lib.zv_timing_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fp2.preprocess_crops(images, None, torch.float16, window_order=True, out=out)
        e1.record()
        torch.cuda.synchronize()
        lib.zv_timing_enable(0)
        ms = e0.elapsed_time(e1) / 5
        cls = {}
        for name, cid in (("k1_hpass", 0), ("k1_vpass", 1)):
            tt, n = C.c_double(), C.c_int64()
            lib.zv_timing_read(cid, C.byref(tt), C.byref(n))
            cls[name] = tt.value / 5
        nbytes = args.bench * (5000 * 5000 * 3 + 4900 * 1176 * 2)
        print(json.dumps({"k1_alone_ms": ms, "images": args.bench, "GBps": nbytes / ms / 1e6, "classes_ms": cls}))
    print("ALL OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
