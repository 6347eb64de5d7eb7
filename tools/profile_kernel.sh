#!/bin/bash
# ncu --set full capture of ONE kernel (base-name regex) of one masking pass after three warm-up passes.
# usage: tools/profile_kernel.sh TAG WORKLOAD KERNEL_REGEX [SKIP]
tag=$1; wl=$2; k=$3; skip=${4:-3}
lean="--no-e2e --no-cpu-baseline --no-fastq --no-bam --no-strong --others none"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$k" --launch-skip $skip -c 1 \
  -f -o gpurun_out/prof_${tag}_${wl} python bench.py --steps 1 --warmup 3 --workload $wl $lean > gpurun_out/${tag}_ncu_${wl}.log 2>&1
