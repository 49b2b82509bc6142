"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name → share of device time."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*$", "", name)
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "ns")
        if unit in ("us", "usecond"):
            v *= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e6
        agg[name][0] += 1
        agg[name][1] += v
        total += v
    print(f"total device time {total / 1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / total * 100:6.2f}%  {t / 1e6:9.3f} ms  {n:5d}x  {name[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
