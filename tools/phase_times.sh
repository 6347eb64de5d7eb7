#!/bin/bash
# Cumulative session-kernel time when sessions stop after phase k (GA_STOP_AFTER profiling knob).
for k in ${PHASES:-1 2 3 4 5 6 7 0}; do
  GA_STOP_AFTER=$k timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 "$@" 2>&1 | tail -1 > /tmp/pt.json
  python -c "import json; d=json.load(open('/tmp/pt.json')); print('stop_after', $k, 'session_ms', round(d['roofline']['session_kernel_ms'],3), 'emit_ms', round(d['roofline']['emit_kernel_ms'],3))"
done
