#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-native paths in libzoomvit.so (runs on the CPU box: cuobjdump only).

    python tools/sass_summary.py > profiles/r02_sass.txt

For every kernel in the shared object: counts of the mnemonics that prove which hardware path it runs on
(B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UBLKCP = TMA,
HMMA = mma.sync (legacy tensor path), IDP = dp4a, LDGSTS = cp.async, plus registers per thread from the ELF.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "zoomearth_b200", "libzoomvit.so")
PAT = collections.OrderedDict([
    ("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCIMMA", r"\bUTCIMMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
    ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.MULTICAST", r"\bUTMALDG\S*MULTICAST"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"),
    ("SYNCS(mbarrier)", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("IMMA", r"\bIMMA"), ("IDP.4A", r"\bIDP\.4A"), ("LDGSTS", r"\bLDGSTS"), ("MUFU.EX2", r"\bMUFU\.EX2"),
    ("ST.E(peer/global)", r"\bST\.E|\bSTG")])


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return r.stdout.splitlines()


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    regs = {}
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            regs[cur] = int(m.group(1))
    blocks = re.split(r"\n\s*Function : ", sass)[1:]
    rows = []
    for b in blocks:
        name = b.split("\n", 1)[0].strip()
        counts = [len(re.findall(p, b)) for p in PAT.values()]
        n_inst = len(re.findall(r"/\*[0-9a-f]{4}\*/", b))
        rows.append((name, n_inst, regs.get(name), counts))
    names = demangle([r[0] for r in rows])
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); columns = instruction counts")
    print("# kernel | instructions | registers | " + " | ".join(PAT))
    for (raw, n_inst, rg, counts), nm in sorted(zip(rows, names), key=lambda t: t[1]):
        nm = re.sub(r"\(anonymous namespace\)::|zv::|void ", "", nm)
        nm = re.sub(r"\(.*", "", nm)
        print(f"{nm} | {n_inst} | {rg} | " + " | ".join(str(c) for c in counts))


if __name__ == "__main__":
    main()
