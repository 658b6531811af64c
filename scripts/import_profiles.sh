#!/bin/bash
# Copies the artefacts scripts/collect_profiles.sh left in gpurun_out/ into profiles/ (only the files that script
# writes) and rebuilds the summary: scripts/import_profiles.sh [tag]
T=${1:-r2}
FR=${FR:-256}
for n in bench_launches replay${FR}_metrics raster_h0_raw raster_h0_source gather_m5_raw gather_m5_source stem_raw stem_source \
         decode_kernel_raw decode_kernel_source nms_kernel_raw nms_kernel_source pnp_kernel_raw pnp_kernel_source \
         dwconv_raw dwconv_source armors_raw armors_source shuffle_unit_raw shuffle_unit_source; do
  cp gpurun_out/${T}_${n}.csv profiles/ || exit 1
done
cp gpurun_out/${T}_ops.json gpurun_out/${T}_src_hash.txt profiles/
cp gpurun_out/${T}_ops_shuffle64.log profiles/${T}_ops_shuffle64.txt
cp gpurun_out/${T}_ops_shuffle64_unfused.log profiles/${T}_ops_shuffle64_unfused.txt
python - <<PY
p='profiles/${T}_shuffle_unit_source.csv'
lines=open(p).read().split('\n')
idx=[i for i,l in enumerate(lines) if l.startswith('"Kernel Name"')]
if len(idx) > 1: open(p,'w').write('\n'.join(lines[:idx[1]])+'\n')      # keep the first unit's source page only
PY
python scripts/summarize_profiles.py ${T} > /dev/null && tail -3 profiles/${T}_traffic.json
