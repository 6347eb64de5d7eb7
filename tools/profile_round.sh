#!/bin/bash
# One GPU call's worth of evidence for a round: launch lists and ncu --set full captures of one masking pass
# (after three warm-up passes) for the headline workload and for cigar-stress.  usage: tools/profile_round.sh TAG
tag=${1:-rXX}
K='regex:scan_kernel|resolve|emit|session_kernel|assign_sessions|clear_kernel'
lean="--no-e2e --no-cpu-baseline --no-fastq --no-bam --no-strong --others none"
for wl in chr1-30x-50k cigar-stress; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 72 --csv \
    --log-file gpurun_out/${tag}_launches_${wl}.csv python bench.py --steps 3 --warmup 3 --workload $wl $lean \
    > gpurun_out/${tag}_launches_${wl}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 36 -c 12 \
    -f -o gpurun_out/prof_${tag}_${wl} python bench.py --steps 1 --warmup 3 --workload $wl $lean \
    > gpurun_out/${tag}_ncu_${wl}.log 2>&1
done
