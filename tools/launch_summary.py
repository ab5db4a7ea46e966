#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`
launch list by (kernel, grid).
usage: tools/launch_summary.py gpurun_out/launches_TAG.csv [forwards]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    fw = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot, dram_tot = collections.OrderedDict(), 0.0, 0.0
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}
    for row in csv.DictReader(lines):
        m = row.get("Metric Name")
        if m not in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        name = row["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")
        k = (name.split("(")[0][:44], row["Grid Size"], row["Block Size"])
        v = float(row["Metric Value"].replace(",", "")) * unit_scale.get(row.get("Metric Unit", ""), 1.0)
        a = agg.setdefault(k, [0, 0.0, 0.0])
        if m == "gpu__time_duration.sum":
            a[0] += 1
            a[1] += v
            tot += v
        else:
            a[2] += v
            dram_tot += v
    print("| kernel | grid | block | launches/fwd | avg us | us/fwd | share % | DRAM MB/fwd |")
    print("|---|---|---|---|---|---|---|---|")
    for k, (n, t, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %s | %s | %.1f | %.1f | %.0f | %.1f | %.1f |" % (k[0], k[1], k[2], n / fw, t / n / 1e3, t / fw / 1e3, 100 * t / tot, d / fw / 1e6))
    print("total us/fwd: %.0f   DRAM MB/fwd: %.1f" % (tot / fw / 1e3, dram_tot / fw / 1e6))


if __name__ == "__main__":
    main()
