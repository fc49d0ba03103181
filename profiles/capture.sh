#!/bin/bash
# usage (under gpurun): bash profiles/capture.sh <tag> <kernel-regex> <skip> <bench args...>
# plain run first (must exit 0), then one ncu --set full capture of the named kernel
tag=$1; regex=$2; skip=$3; shift 3
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu $*"
timeout 300 $B > gpurun_out/plain_$tag.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 \
    -o gpurun_out/prof_$tag -f $B > gpurun_out/ncu_$tag.log 2>&1
echo "$tag rc=$?"
