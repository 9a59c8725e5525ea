#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares (markdown)."""
import collections
import csv
import re
import sys


def main(path, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("egoego::", "")[:80]
        a = agg.setdefault((name, row["Grid Size"], row["Block Size"]), [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"# {title}\n\nsource: `{path}` ({sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time; "
          "ncu per-launch times are cold-cache and serialised: compare SHARES)\n")
    print("| share | launches | avg us | kernel | grid | block |\n|---:|---:|---:|---|---|---|")
    for (name, grid, block), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {t / tot * 100:.2f}% | {n} | {t / n / 1e3:.1f} | `{name}` | {grid} | {block} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "kernel launch list")
