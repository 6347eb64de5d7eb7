#!/usr/bin/env python
"""One line per captured kernel of an .ncu-rep: duration, occupancy, issue rate, instructions, DRAM bytes.
usage: tools/ncu_summary.py REPORT.ncu-rep"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "ms", 1e-6), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("launch__registers_per_thread", "regs", 1), ("smsp__inst_executed.sum", "Ginst", 1e-9), ("dram__bytes_read.sum", "rdGB", None),
        ("dram__bytes_write.sum", "wrGB", None), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst", 1), ("launch__grid_size", "grid", 1),
        ("launch__occupancy_limit_shared_mem", "occ_smem", 1), ("launch__occupancy_limit_registers", "occ_regs", 1), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1)]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        parts = [r[h.index("Kernel Name")][:44].ljust(44)]
        for name, label, scale in WANT:
            if name not in h:
                continue
            i = h.index(name)
            v = float(r[i].replace(",", "")) if r[i] else 0.0
            if scale is None:
                u = units[i]
                v *= {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1.0)
                parts.append(f"{label}={v:.3f}")
            elif name == "gpu__time_duration.sum":
                u = units[i]
                v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
                parts.append(f"{label}={v:.3f}")
            else:
                parts.append(f"{label}={v * scale:.2f}")
        print("  ".join(parts))


if __name__ == "__main__":
    main()
