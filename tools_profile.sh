#!/bin/bash
# usage (under gpurun): bash tools_profile.sh <tag>
# 1) plain run, 2) launch list of the same command, 3) ncu --set full of the fused MLP kernel
set -x
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 3 --no-train --no-cpu"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 2 -o gpurun_out/${TAG}_mlp_tc $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_ncu2.log
