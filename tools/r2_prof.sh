#!/bin/bash
# round 2: one ncu --set full capture of a kernel on the bench workload.  usage: tools/r2_prof.sh <tag> <codec> <op> <kernel-regex> <blocks> [skip]
tag=$1; codec=$2; op=$3; rx=$4; blocks=$5; skip=${6:-3}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/prof_r2_$tag \
    python bench.py --profile --codec $codec --op $op --steps 1 --warmup 3 --blocks $blocks > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-160
