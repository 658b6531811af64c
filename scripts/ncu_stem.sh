#!/bin/bash
# ncu --set full capture of the stem kernel of one eager replay of $FR frames
tag=${1:-r2_stem}
ncu --set full --clock-control none --import-source on --kernel-name "regex:stem_bayer2x_kernel|stem_kernel" --launch-skip 2 --launch-count 1 \
  -o gpurun_out/${tag} -f python scripts/profile_replay.py ${FR:-256} 1 > gpurun_out/${tag}.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
