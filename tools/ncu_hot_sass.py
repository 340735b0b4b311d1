#!/usr/bin/env python
"""Top stall-sample SASS instructions of a kernel in an .ncu-rep (source page, SASS view) with their dominant stall reason."""
import csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
I = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[2:]):
    try:
        s = float(r[I["# Samples"]] or 0)
    except Exception:
        continue
    st = sorted(((float(r[I[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    data.append((s, n, r[I["Source"]].strip(), r[I["Instructions Executed"]], st))
tot = sum(d[0] for d in data) or 1
print(f"total samples {tot:.0f}")
for s, n, src, ex, st in sorted(data, reverse=True)[:top]:
    print(f"{100 * s / tot:5.1f}%  #{n:5d} exec {ex:>9s}  {src[:90]:90s} {st[0][1]}={st[0][0]:.0f} {st[1][1]}={st[1][0]:.0f}")
