#!/usr/bin/env python
"""Summarise an ncu `--csv` launch list: per kernel the launch count, mean duration, share of the total, and (when they
were collected) the mean DRAM bytes read / written per launch."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        metric = row["Metric Name"]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        if metric.startswith("gpu__time_duration"):
            a["n"] += 1
            a["us"] += v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        elif metric.startswith("dram__bytes"):
            mb = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6) * v
            a["rd" if "read" in metric else "wr"] += mb
    total = sum(a["us"] for a in agg.values())
    print(f"{'kernel':60s} {'n':>5s} {'mean us':>9s} {'share':>7s} {'rd MB':>8s} {'wr MB':>8s}")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = max(a["n"], 1)
        print(f"{name[:60]:60s} {a['n']:5d} {a['us'] / n:9.1f} {100 * a['us'] / total:6.1f}% {a['rd'] / n:8.1f} {a['wr'] / n:8.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
