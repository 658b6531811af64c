#!/bin/bash
# per-launch metrics of one eager replay of $1 frames (ncu launch list with a few counters)
n=${1:-128}; tag=${2:-r1_replay${n}_metrics}
python scripts/profile_replay.py $n 1 > gpurun_out/${tag}_plain.log 2>&1 || exit 1
L=$(grep -o "launches per replay [0-9]*" gpurun_out/${tag}_plain.log | grep -o "[0-9]*$")
# the script runs 2 enqueue replays + 1 profile_stages replay; capture the last one (only irmv:: kernels are counted)
SKIP=$(( 2 * L ))   # L counts every irmv:: kernel of a replay, set_src included
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__shared_mem_per_block_dynamic,launch__registers_per_thread,launch__grid_size,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --kernel-name "regex:^(conv_raster_kernel|conv_tc_kernel|decode_kernel|nms_kernel|pnp_kernel|quads_from_dets_kernel|set_src_kernel|sppf_pool_kernel|stem_kernel|stem_bayer2x_kernel)$" -s $SKIP -c $L --csv --log-file gpurun_out/${tag}.csv python scripts/profile_replay.py $n 1 > gpurun_out/${tag}_ncu.log 2>&1
