#!/usr/bin/env python
"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised launches:
compare SHARES with the in-step CUDA-event numbers, not absolutes).  usage: summarize_launches.py launches.csv [title]"""
import collections
import csv
import re
import sys


def main(path, title=""):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    total, n = 0.0, 0
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    if title:
        print(title)
    print(f"total kernel time in window: {total:.1f} ms over {n} launches\n")
    for name, (k, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:10.2f} ms  {100 * ms / total:5.1f}%  n={k:5d}  {name[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
