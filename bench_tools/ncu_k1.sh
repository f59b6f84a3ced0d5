#!/bin/bash
# usage (under gpurun): bash bench_tools/ncu_k1.sh <tag> [quick_k1.py shape]   -> gpurun_out/<tag>.ncu-rep
tag=$1; shift
ncu --set full --clock-control none --import-source on -k regex:level_step -s 5 -c 1 -f -o gpurun_out/$tag python bench_tools/quick_k1.py "$@" > gpurun_out/$tag.log 2>&1
tail -2 gpurun_out/$tag.log
