#!/bin/bash
# One GPU call: plain run first, then the ncu launch list and one --set full capture of the hot kernels.
# usage: gpu_profile.sh <tag> <kernel-regex> [bench args...]
TAG=${1:-poisson}; REGEX=${2:-'k_(num|sym)_merge|k_flop_count|k_scan'}; shift; shift
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s 12 -c 4 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
