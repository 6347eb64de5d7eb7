#!/usr/bin/env python
"""Instruction / stall-sample share per source region of an .ncu-rep.

usage: tools/ncu_regions.py REPORT.ncu-rep file:lo-hi=name [...]   (lines not covered are grouped by file)
"""
import csv
import os
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    regions = []
    for a in sys.argv[2:]:
        loc, name = a.split("=")
        f, rng = loc.split(":")
        lo, hi = rng.split("-")
        regions.append((f, int(lo), int(hi), name))
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"]
    if os.environ.get("KERNEL"):                                          # base-name regex of the kernel to look at
        cmd += ["--kernel-name", "regex:" + os.environ["KERNEL"]]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    fname, hdr = None, None
    inst, thr, smp = defaultdict(int), defaultdict(int), defaultdict(int)
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            n = len(hdr)
            try:
                i, t, s = int(r[hdr.index("Instructions Executed") - n] or 0), int(r[hdr.index("Thread Instructions Executed") - n] or 0), int(r[hdr.index("# Samples") - n] or 0)
            except ValueError:
                continue
            ln = int(r[0])
            key = fname
            for f, lo, hi, name in regions:
                if f == fname and lo <= ln <= hi:
                    key = name
                    break
            inst[key] += i; thr[key] += t; smp[key] += s
    ti, ts = sum(inst.values()) or 1, sum(smp.values()) or 1
    print(f"warp instructions {ti}, samples {ts}")
    for k in sorted(inst, key=lambda k: -inst[k]):
        print(f"  inst {100 * inst[k] / ti:5.1f}%  samples {100 * smp[k] / ts:5.1f}%  active lanes {thr[k] / max(1, inst[k]):5.1f}  {k}")


if __name__ == "__main__":
    main()
