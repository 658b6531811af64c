#!/bin/bash
# Run on the GPU box (gpurun): every ncu artefact that profiles/ is built from.  Each capture runs
# only after the same command has exited 0 without ncu.  usage: scripts/collect_profiles.sh [tag]
set -u
T=${1:-r2}
FR=${FR:-256}            # frames per replay = the bench's (bench.py replay_split)
KRX='regex:^(conv_raster_kernel|conv_tc_kernel|decode_kernel|nms_kernel|pnp_kernel|quads_from_dets_kernel|set_src_kernel|sppf_pool_kernel|stem_kernel|stem_bayer2x_kernel|dwconv3x3_kernel|shuffle_unit_kernel|kpts_from_dets_kernel|quads_from_kpts_kernel)$'
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --clock-warmup-s 0"
# 1. launch list of the bench command: 3 warm-up steps + the timed step (one replay each), then the end-to-end loop
$BENCH > gpurun_out/${T}_bench_plain.json 2> gpurun_out/${T}_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KRX" -c 560 --csv \
    --log-file gpurun_out/${T}_bench_launches.csv $BENCH > gpurun_out/${T}_bench_ncu.log 2>&1
# 2. per-launch counters of one eager replay of the bench's size
scripts/ncu_replay_metrics.sh $FR ${T}_replay${FR}_metrics
cp gpurun_out/ops.json gpurun_out/${T}_ops.json
# 3. full captures: top GEMM (Detect P3 box.0|cls.0), the gather kernel's largest launch (m5), the stem,
#    decode, NMS, PnP
OP=$(python -c "import sys, json; sys.path.insert(0, 'scripts'); from analyze_launches import name_ops; print(1 + [n[0] for n in name_ops(json.load(open('gpurun_out/${T}_ops.json'))) if n[0] != 'POOL'].index('h0.01'))")
scripts/ncu_one_conv.sh $OP $FR ${T}_raster_h0
OP2=$(python -c "import sys, json; sys.path.insert(0, 'scripts'); from analyze_launches import name_ops; print(1 + [n[0] for n in name_ops(json.load(open('gpurun_out/${T}_ops.json'))) if n[0] != 'POOL'].index('m5'))")
RASTER= scripts/ncu_one_conv.sh $OP2 $FR ${T}_gather_m5 conv_tc_kernel
FR=$FR scripts/ncu_stem.sh ${T}_stem
for k in decode_kernel nms_kernel pnp_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:^${k}$ --launch-skip 2 --launch-count 1 \
    -o gpurun_out/${T}_${k} -f python scripts/profile_replay.py $FR 1 > gpurun_out/${T}_${k}.log 2>&1
  ncu -i gpurun_out/${T}_${k}.ncu-rep --page raw --csv > gpurun_out/${T}_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${T}_${k}.ncu-rep --page source --csv > gpurun_out/${T}_${k}_source.csv 2>/dev/null
done
# 4. ShuffleNetV2 variant at 64 frames: per-op times fused / unfused, the fused unit kernels (seven launches of one
#    replay), and the depthwise 3x3 kernel of the one-launch-per-conv path (largest launch: d1.b1.dw, 16 channels,
#    320x320 -> 160x160)
python scripts/profile_ops_any.py 64 shufflenetv2-pose > gpurun_out/${T}_ops_shuffle64.log 2>&1 || exit 1
python scripts/profile_ops_any.py 64 shufflenetv2-pose unfused > gpurun_out/${T}_ops_shuffle64_unfused.log 2>&1 || exit 1
scripts/ncu_shuffle_unit.sh ${T}_shuffle_unit
ncu --set full --clock-control none --import-source on --kernel-name regex:dwconv3x3_kernel --launch-skip 39 --launch-count 1 \
  -o gpurun_out/${T}_dwconv -f python scripts/profile_ops_any.py 64 shufflenetv2-pose unfused > gpurun_out/${T}_dwconv.log 2>&1
ncu -i gpurun_out/${T}_dwconv.ncu-rep --page raw --csv > gpurun_out/${T}_dwconv_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_dwconv.ncu-rep --page source --csv > gpurun_out/${T}_dwconv_source.csv 2>/dev/null
# 5. light-bar / armor extraction on its own workload
python scripts/bench_armors.py 64 10 > gpurun_out/${T}_armors_plain.json 2> gpurun_out/${T}_armors_plain.err || exit 1
ncu --set full --clock-control none --import-source on --kernel-name regex:extract_armors_kernel --launch-skip 4 --launch-count 1 \
  -o gpurun_out/${T}_armors -f python scripts/bench_armors.py 64 10 > gpurun_out/${T}_armors.log 2>&1
ncu -i gpurun_out/${T}_armors.ncu-rep --page raw --csv > gpurun_out/${T}_armors_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_armors.ncu-rep --page source --csv > gpurun_out/${T}_armors_source.csv 2>/dev/null
python -c "
import sys; sys.path.insert(0, '.')
import bench; open('gpurun_out/${T}_src_hash.txt', 'w').write(bench.source_hash())"
rm -f gpurun_out/*.ncu-rep
