#!/usr/bin/env python
"""Top stalled SASS instructions from `ncu -i X.ncu-rep --page source --csv [--launch-skip n --launch-count 1]`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = next(r for r in rows if r and r[0] == "Address")
I = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[I["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
for r in sorted(data, key=lambda r: -int(r[I["# Samples"]] or 0))[:n]:
    s = int(r[I["# Samples"]])
    st = sorted(((h, int(r[I[h]] or 0)) for h in stalls), key=lambda kv: -kv[1])[:2]
    print(f"{s:6d} {100*s/tot:5.1f}%  exec={r[I['Instructions Executed']]:>8s}  {r[I['Source']].strip()[:64]:64s} {st}")
agg = {h: sum(int(r[I[h]] or 0) for r in data) for h in stalls}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
