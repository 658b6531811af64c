#!/bin/bash
# Run on the GPU box (gpurun): every ncu artefact that profiles/ is built from.  Each capture runs
# only after the same command has exited 0 without ncu.
set -u
KRX='regex:^(conv_raster_kernel|conv_tc_kernel|decode_kernel|nms_kernel|pnp_kernel|quads_from_dets_kernel|set_src_kernel|sppf_pool_kernel|stem_kernel)$'
mkdir -p gpurun_out
# 1. launch list of the bench command (timed region = the last 132 launches of our kernels)
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_plain.json 2> gpurun_out/r1_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KRX" -s 374 -c 420 --csv \
    --log-file gpurun_out/r1_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_ncu.log 2>&1
# 2. per-launch counters of one eager 128-frame replay
scripts/ncu_replay_metrics.sh 128 r1_replay128_metrics
# 3. full captures: top GEMM (Detect P3 box.0|cls.0, network op 45) and the stem
cp gpurun_out/ops.json gpurun_out/r1_ops.json
OP=$(python -c "import sys, json; sys.path.insert(0, 'scripts'); from analyze_launches import name_ops; print(1 + [n[0] for n in name_ops(json.load(open('gpurun_out/r1_ops.json'))) if n[0] != 'POOL'].index('h0.01'))")
scripts/ncu_one_conv.sh $OP 128 r1_raster_h0
scripts/ncu_stem.sh r1_stem
# 4. light-bar / armor extraction on its own workload
python scripts/bench_armors.py 64 10 > gpurun_out/r1_armors_plain.json 2> gpurun_out/r1_armors_plain.err || exit 1
ncu --set full --clock-control none --import-source on --kernel-name regex:extract_armors_kernel --launch-skip 4 --launch-count 1 \
  -o gpurun_out/r1_armors -f python scripts/bench_armors.py 64 10 > gpurun_out/r1_armors.log 2>&1
ncu -i gpurun_out/r1_armors.ncu-rep --page raw --csv > gpurun_out/r1_armors_raw.csv 2>/dev/null
ncu -i gpurun_out/r1_armors.ncu-rep --page source --csv > gpurun_out/r1_armors_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
