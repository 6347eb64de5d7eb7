#!/bin/bash
# Per-kernel durations (us) of one masking pass at full size: ncu launch list of the engine's kernels, last step.
# (cold-cache, serialised: use for shares, not for absolute numbers)  usage: tools/kernel_times.sh [bench.py args]
out=${OUT:-gpurun_out/launches_tmp.csv}
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"scan_kernel|resolve|emit|session_kernel|assign_sessions|clear_kernel" -c 80 --csv --log-file $out \
  python bench.py --steps 1 --no-e2e --no-cpu-baseline --no-fastq "$@" > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$out")) if len(r)>5]
h=rows[0]
for r in rows[-8:]:
    print(f"{float(r[h.index('Metric Value')])/1e3:9.1f} us  {r[h.index('Kernel Name')][:48]}")
PY
