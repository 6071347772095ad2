#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total, mean, share."""
import csv
import sys
from collections import OrderedDict


def main(path, header):
    rows = [l for l in open(path) if not l.startswith("==")]
    agg = OrderedDict()
    for r in csv.DictReader(rows):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit.replace("usecond", "us").replace("nsecond", "ns").replace("msecond", "ms").replace("second", "s")]
        name = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print("# " + header)
    print("# (ncu serialises the launches; per-launch times are cold-cache) kernel, launches, total ms, ms/launch, share")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:46s} {n:4d} {ms:9.3f} {ms / n:8.3f} {ms / tot:6.3f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
