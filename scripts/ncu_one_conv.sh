#!/bin/bash
# ncu --set full capture of ONE convolution launch (op index $1, frames $2) via the trace entry point.
# usage: scripts/ncu_one_conv.sh <op> <frames> <tag> [kernel-regex]
op=$1; n=$2; tag=$3; rx=${4:-conv_raster_kernel}
WARM=0 RASTER=${RASTER-1} ncu --set full --clock-control none --import-source on --kernel-name regex:$rx --launch-skip 1 --launch-count 1 \
  -o gpurun_out/${tag} -f python scripts/trace_conv.py $n $op > gpurun_out/${tag}.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
