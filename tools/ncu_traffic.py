#!/usr/bin/env python
"""Turns the ncu CSVs of tools/gpu_ncu_traffic.sh (metrics dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum per launch, bench workload at 64 images) into profiles/ncu_traffic.json, the file bench.py reads
for its `roofline.traffic` fields - so the numbers in the bench line are a capture of the tree they are printed by
(`commit` in the file), not constants typed into bench.py.

    python tools/ncu_traffic.py gpurun_out/traffic_gemm_64img.csv gpurun_out/traffic_other_64img.csv [--images 64]
"""
import argparse
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(path):
    """-> list of launches in order: dict(kernel, read, write, ns, tensor_pct)."""
    rows = [l for l in open(path, newline="") if l.startswith('"')]
    launches = {}
    for r in csv.DictReader(rows):
        d = launches.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"], "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        name, unit = r["Metric Name"], r["Metric Unit"]
        if name == "dram__bytes_read.sum":
            d["read"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        elif name == "dram__bytes_write.sum":
            d["write"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        elif name == "gpu__time_duration.sum":
            d["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        elif name.startswith("sm__pipe_tensor"):
            d["tensor_pct"] = v
    return [launches[k] for k in sorted(launches)]


def short(kernel):
    k = kernel.replace("void ", "").replace("zv::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return k.split("(")[0].strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv", nargs="+")
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ncu_traffic.json"))
    args = ap.parse_args()
    per = defaultdict(list)
    for p in args.csv:
        for l in read(p):
            if "read" in l and "write" in l:
                per[short(l["kernel"])].append(l)
    summary = {}
    for k, ls in per.items():
        n = len(ls)
        summary[k] = {"launches": n, "dram_bytes_per_launch": sum(l["read"] + l["write"] for l in ls) / n,
                      "read": sum(l["read"] for l in ls) / n, "write": sum(l["write"] for l in ls) / n,
                      "us": sum(l.get("ns", 0) for l in ls) / n / 1e3,
                      "tensor_pct": (sum(l.get("tensor_pct", 0) for l in ls) / n) if any("tensor_pct" in l for l in ls) else None}
    gemm = [v for k, v in summary.items() if k.startswith("gemm_tc")]
    k1 = [v for k, v in summary.items() if k.startswith("k1_")]
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    out = {"commit": commit, "images": args.images, "sources": [os.path.basename(p) for p in args.csv], "kernels": summary}
    if gemm:
        n = sum(v["launches"] for v in gemm)
        out["gemm_bytes_per_launch"] = sum(v["dram_bytes_per_launch"] * v["launches"] for v in gemm) / n
        out["gemm_note"] = (f"mean DRAM bytes (read + write) per launch over {n} captured gemm_tc launches of one block "
                            f"(QKV, proj, SwiGLU, down) at {args.images} images, ncu at commit {commit} "
                            f"({', '.join(out['sources'])})")
    if k1:
        # the captured K1 launches of one step together cover the whole batch (two passes of k1_resample_tc + the patchify)
        out["k1_bytes_per_image"] = sum(v["dram_bytes_per_launch"] * v["launches"] for v in k1) / args.images
        out["k1_note"] = (f"DRAM bytes (read + write) of the K1 launches of one step / {args.images} images, ncu at commit "
                          f"{commit}; algorithmic 86.52 MB per image")
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    sys.exit(main())
