#!/usr/bin/env python
"""Exports what profiles/ keeps from one `ncu --set full` report and one launch list:
    tools/ncu_profile_export.py REPORT.ncu-rep LAUNCHES.csv PREFIX
writes PREFIX_kernels_raw.csv (selected metrics, one row per captured kernel), PREFIX_traffic.json (DRAM bytes per
kernel; bench.py reads roofline.traffic from it) and prints the per-kernel launch averages of the launch list."""
import collections
import csv
import json
import re
import subprocess
import sys

KEEP = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, launches, prefix = sys.argv[1:4]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(h)}
    keep = [k for k in KEEP if k in ix]
    with open(prefix + "_kernels_raw.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([u[ix[k]] for k in keep])
        for r in rows[2:]:
            w.writerow([r[ix[k]] for k in keep])
    per = {}
    for r in rows[2:]:
        name = re.match(r"(?:void )?(?:ga::)?(\w+)", r[ix["Kernel Name"]]).group(1)
        tm = re.search(r"<([^>]*)>", r[ix["Kernel Name"]])
        if tm:                                                          # template instantiations are different kernels
            name += "<" + tm.group(1).replace(" ", "") + ">"
        per[name] = [float(r[ix[k]]) * UNIT[u[ix[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")]
        print(f"{name[:44]:44s} {float(r[ix['gpu__time_duration.sum']]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[u[ix['gpu__time_duration.sum']]]:8.1f} us  read {per[name][0] / 1e9:6.3f} GB  write {per[name][1] / 1e9:6.3f} GB  "
              f"inst {float(r[ix['smsp__inst_executed.sum']]) / 1e6:7.1f} M  issue {float(r[ix['sm__issue_active.avg.pct_of_peak_sustained_elapsed']]):5.1f} %  "
              f"warps {float(r[ix['sm__warps_active.avg.pct_of_peak_sustained_active']]):5.1f} %  regs {r[ix['launch__registers_per_thread']]}  "
              f"grid {r[ix['launch__grid_size']]} x {r[ix['launch__block_size']]}")
    total = sum(a + b for a, b in per.values())
    json.dump({"workload": sys.argv[4] if len(sys.argv) > 4 else "chr1-30x-50k", "windows": 50000,
               "source": prefix.split("/")[-1] + "_kernels_raw.csv under profiles/ (ncu --set full, one launch per kernel)",
               "dram_bytes_per_pass": int(total), "per_kernel": per}, open(prefix + "_traffic.json", "w"), indent=1)
    print(f"DRAM bytes per pass: {total / 1e9:.3f} GB")
    lrows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hh = lrows[0]
    kn, mv = hh.index("Kernel Name"), hh.index("Metric Value")
    agg = collections.OrderedDict()
    for r in lrows[1:]:
        if r[kn].startswith("ga::") or "ga::" in r[kn]:
            agg.setdefault(r[kn], []).append(float(r[mv]))
    for k, v in agg.items():
        print(f"{len(v):3d} x {sum(v) / len(v) / 1e3:9.1f} us  {k[:60]}")


if __name__ == "__main__":
    main()
