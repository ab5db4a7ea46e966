#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid).
usage: tools/launch_summary.py gpurun_out/launches_TAG.csv [forwards]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    fw = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")
        k = (name.split("(")[0][:44], row["Grid Size"], row["Block Size"])
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print("| kernel | grid | block | launches/fwd | avg us | us/fwd | share % |")
    print("|---|---|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %s | %s | %.1f | %.1f | %.0f | %.1f |" % (k[0], k[1], k[2], n / fw, t / n / 1e3, t / fw / 1e3, 100 * t / tot))
    print("total us/fwd: %.0f" % (tot / fw / 1e3))


if __name__ == "__main__":
    main()
