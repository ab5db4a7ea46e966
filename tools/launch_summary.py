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
    # HRNet conv stack only (what bench.py's `roofline` describes): launches, serialised time and DRAM bytes per forward
    hr = [(k, v) for k, v in agg.items() if k[0].startswith(("conv_umma", "stem1", "head_kernel", "upsample_add"))]
    n_l = sum(v[0] for _, v in hr) / fw
    t_l = sum(v[1] for _, v in hr) / fw
    d_l = sum(v[2] for _, v in hr) / fw
    print("HRNet stack: %.0f launches/fwd, %.0f us/fwd serialised, DRAM %.1f MB/fwd (%.2f MB per launch)" % (n_l, t_l / 1e3, d_l / 1e6, d_l / 1e6 / max(n_l, 1)))
    if len(sys.argv) > 3:
        import json
        json.dump({"source": path, "forwards": fw, "launches_per_forward": n_l, "dram_bytes_per_forward": d_l,
                   "dram_bytes_per_launch": d_l / max(n_l, 1), "serialised_us_per_forward": t_l / 1e3,
                   "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over bench.py; HRNet kernels only"},
                  open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
