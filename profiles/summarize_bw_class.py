#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
the bandwidth-class kernels of one training step: per kernel (template instance + grid), launches, total time, DRAM
bytes moved and achieved DRAM GB/s against the measured HBM peak (MEASURED_PEAKS.json).  ncu serialises launches with
cold caches, so these are per-launch figures in isolation, not in-step times."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path):
    peak = 6552.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(lines)
    launches = collections.OrderedDict()
    for r in rd:
        k = launches.setdefault(r["ID"], dict(name=r["Kernel Name"], grid=r["Grid Size"]))
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            k["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            k[r["Metric Name"]] = v * mult
    agg = collections.OrderedDict()
    for k in launches.values():
        name = re.sub(r"\(.*", "", k["name"]).replace("void ", "")
        a = agg.setdefault(name, dict(n=0, ns=0.0, rd=0.0, wr=0.0))
        a["n"] += 1
        a["ns"] += k.get("ns", 0.0)
        a["rd"] += k.get("dram__bytes_read.sum", 0.0)
        a["wr"] += k.get("dram__bytes_write.sum", 0.0)
    print(f"{'kernel':58s} {'n':>5s} {'total ms':>9s} {'read GB':>8s} {'write GB':>8s} {'GB/s':>7s} {'of peak':>7s}   (HBM peak {peak:.0f} GB/s)")
    tot = dict(ns=0.0, b=0.0)
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        gbs = (a["rd"] + a["wr"]) / a["ns"] if a["ns"] else 0.0
        print(f"{name[:58]:58s} {a['n']:5d} {a['ns'] / 1e6:9.3f} {a['rd'] / 1e9:8.3f} {a['wr'] / 1e9:8.3f} {gbs:7.0f} {gbs / peak:7.2f}")
        tot["ns"] += a["ns"]
        tot["b"] += a["rd"] + a["wr"]
    print(f"{'TOTAL':58s} {sum(a['n'] for a in agg.values()):5d} {tot['ns'] / 1e6:9.3f} {tot['b'] / 1e9:17.3f} {tot['b'] / tot['ns']:7.0f} {tot['b'] / tot['ns'] / peak:7.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
