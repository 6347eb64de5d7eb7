#!/usr/bin/env python
"""Top source lines by warp-stall samples from an .ncu-rep (needs -lineinfo and --import-source on).

usage: tools/ncu_hotlines.py REPORT.ncu-rep [N] [KERNEL_REGEX] [samples|inst]
"""
import csv
import os
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    kern = sys.argv[3] if len(sys.argv) > 3 else None
    by_inst = len(sys.argv) > 4 and sys.argv[4] == "inst"
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"]
    if kern:
        cmd += ["--kernel-name", "regex:" + kern]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr, data = None, None, []
    want = os.environ.get("FUNC_SUBSTR")                  # e.g. "(int)768": one template instantiation (ncu matches base names only)
    take = True
    for r in rows:
        if not r:
            continue
        if r[0] == "Function Name":
            take = want is None or want in r[1]
        if not take:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            si = hdr.index("# Samples") - len(hdr)          # count from the end: quotes inside source text split the row
            ie = hdr.index("Instructions Executed") - len(hdr)
            try:
                data.append((int(r[si] or 0), int(r[ie] or 0), fname, int(r[0]), r[1].strip()[:120]))
            except ValueError:
                pass
    tot = sum(d[0] for d in data) or 1
    toti = sum(d[1] for d in data) or 1
    print(f"total samples {tot}, warp instructions {toti}")
    for n, ie, f, ln, src in sorted(data, key=(lambda d: d[1]) if by_inst else (lambda d: d[0]), reverse=True)[:top]:
        print(f"{n:7d} {100 * n / tot:5.1f}%  inst {100 * ie / toti:5.1f}%  {f}:{ln:<4d} {src}")


if __name__ == "__main__":
    main()
