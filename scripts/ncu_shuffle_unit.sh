#!/bin/bash
# ncu --set full capture of the fused ShuffleNetV2 unit kernels of one eager 64-frame replay
# (launch-skip past the warm-up replays; captures the seven units of one replay)
tag=${1:-r2_shuffle_unit}
ncu --set full --clock-control none --import-source on --kernel-name "regex:shuffle_unit_kernel" --launch-skip 21 --launch-count 7 \
  -o gpurun_out/${tag} -f python scripts/profile_ops_any.py 64 shufflenetv2-pose > gpurun_out/${tag}.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
