#!/bin/bash
# Run on the GPU box (gpurun): every ncu artefact that profiles/ is built from.  Each capture runs
# only after the same command has exited 0 without ncu.
set -u
KRX='regex:^(conv_raster_kernel|conv_tc_kernel|decode_kernel|nms_kernel|pnp_kernel|quads_from_dets_kernel|set_src_kernel|sppf_pool_kernel|stem_kernel)$'
mkdir -p gpurun_out
# 1. launch list of the bench command (timed region = the last 132 launches of our kernels)
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_plain.json 2> gpurun_out/r1_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KRX" -s 394 -c 420 --csv \
    --log-file gpurun_out/r1_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_ncu.log 2>&1
# 2. per-launch counters of one eager 128-frame replay
scripts/ncu_replay_metrics.sh 128 r1_replay128_metrics
# 3. full captures: top GEMM (Detect P3 box.0|cls.0, network op 45) and the stem
cp gpurun_out/ops.json gpurun_out/r1_ops.json
OP=$(python -c "import sys, json; sys.path.insert(0, 'scripts'); from analyze_launches import name_ops; print(1 + [n[0] for n in name_ops(json.load(open('gpurun_out/r1_ops.json'))) if n[0] != 'POOL'].index('h0.01'))")
scripts/ncu_one_conv.sh $OP 128 r1_raster_h0
scripts/ncu_stem.sh r1_stem
rm -f gpurun_out/*.ncu-rep
