#!/usr/bin/env python
"""Per-kernel table (launches, total ms, share, mean ms) from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`).   python scripts/summarize_launches.py X.csv"""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r["Kernel Name"]
        name = name.split("(")[0]
        rows.append((name, ms))
    tot = sum(ms for _, ms in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, ms in rows:
        agg[n][0] += 1
        agg[n][1] += ms
    print("%d launches, %.1f ms under ncu\n" % (len(rows), tot))
    print("| kernel | launches | total ms | share | mean ms |")
    print("|---|---|---|---|---|")
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% | %.4f |" % (n[:70], c, ms, 100 * ms / tot, ms / c))


if __name__ == "__main__":
    main(sys.argv[1])
