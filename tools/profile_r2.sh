#!/bin/bash
# usage (under gpurun, ONE GPU): bash tools/profile_r2.sh <tag>
#   1) plain runs (must exit 0), 2) ncu launch list of the render step and of one training step (with DRAM bytes),
#   3) ncu --set full of the fused MLP kernel (render) and of the three training tensor-core kernels.
set -x
TAG=${1:-r2}
R="python bench.py --steps 1 --warmup 3 --no-train --no-cpu --no-extra"
T="python tools/train_profile.py 1024 1"
$R > gpurun_out/${TAG}_render_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_render_plain.log; exit 1; }
$T > gpurun_out/${TAG}_train_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_train_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_render.csv $R > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_train.csv $T > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_tc7_kernel -s 8 -c 2 -o gpurun_out/${TAG}_mlp_tc7 $R > gpurun_out/${TAG}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:mlp_tc_kernel|bwd_tc_kernel|dw_grouped_kernel" -s 6 -c 6 -o gpurun_out/${TAG}_train_tc $T > gpurun_out/${TAG}_ncu4.log 2>&1
tail -3 gpurun_out/${TAG}_ncu3.log gpurun_out/${TAG}_ncu4.log
