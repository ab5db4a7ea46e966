#!/usr/bin/env python
"""Compact table of the metrics that matter from `ncu --set full` reports.
usage: tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [more.ncu-rep ...]  > profiles/rNN_ncu_summary.md"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (active SMs)"),
    ("sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "UTCHMMA fp16 % of peak (all SMs, elapsed)"),
    ("sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.max.pct_of_peak_sustained_elapsed", "UTCHMMA fp16 % of peak (busiest SM)"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor memory (TMA) pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
]


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        print("### %s\n" % path.split("/")[-1])
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].replace("void <unnamed>::", "").split("(")[0]
            print("**%s** grid %s block %s\n" % (name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            print("| metric | value |\n|---|---|")
            for k, label in KEYS:
                if k in hdr and r[hdr.index(k)] not in ("", "n/a"):
                    print("| %s (`%s`) | %s %s |" % (label, k, r[hdr.index(k)], units[hdr.index(k)]))
            print()


if __name__ == "__main__":
    main()
