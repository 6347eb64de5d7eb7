#!/usr/bin/env python
"""Prints the timing fields of a bench.py JSON line: tools/show_stage.py gpurun_out/b.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(f"step {d['ms_per_step']:.3f} ms  pass {r['kernel_ms']:.3f} ms  frac {r['frac']:.3f}  " +
      "  ".join(f"{k} {v:.3f}" for k, v in r["stage_ms"].items()) + f"  parity {d.get('parity_vs_oracle_on_sample')}")
