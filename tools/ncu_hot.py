#!/usr/bin/env python
"""Stall-reason totals and SASS hot regions of an ncu report (source page):  ncu_hot.py <report.ncu-rep> [kernel index] [bucket]"""
import csv
import re
import subprocess
import sys
from collections import Counter


def main(path, which=1, B=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, k, data = None, 0, []
    for r in rows:
        if r and r[0] == "Kernel Name":
            k += 1
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and k == which and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {s: sum(int(d[s] or 0) for d in data) for s in stalls}
    S = sum(tot.values())
    print("instructions", len(data), "samples", S)
    for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
        print(f"  {s:26s}{v:7d} {v / S:.3f}")
    for i in range(0, len(data), B):
        ch = data[i:i + B]
        s = sum(int(d["# Samples"] or 0) for d in ch)
        ex = max(int(d["Instructions Executed"] or 0) for d in ch)
        ops = [(d["Source"].split()[1] if d["Source"].strip().startswith("@") else d["Source"].split()[0]) for d in ch]
        marks = Counter(o for o in ops if re.match(r"LDTM|STTM|LDG|STG|SYNCS|UTC|UTMA|BAR|ERRBAR|MEMBAR|STS|LDS|WARPSYNC|UCGABAR|STL|LDL|ST\b|CCTL|FENCE|STAS", o))
        if s or ex:
            print(f"{i:5d} samples {s:4d} exec {ex:7d}  " + " ".join(f"{k}:{v}" for k, v in marks.most_common(7)))
    hot = sorted(data, key=lambda d: -int(d["# Samples"] or 0))[:25]
    print()
    for d in hot:
        top = max(stalls, key=lambda s: int(d[s] or 0))
        print(f'{d["# Samples"]:>6} {d["Instructions Executed"]:>8} {top:18s} {d["Source"].strip()[:100]}')


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
