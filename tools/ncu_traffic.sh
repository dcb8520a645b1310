#!/bin/bash
# DRAM traffic + duration of every kernel of ONE forward + write_results (YOLOv3-416, batch 64, bench.py's weights,
# autotuned plan, stream launches) under ncu; summarised into profiles/ by tools/ncu_summary.py.
# Run on the GPU box AFTER the same command has exited 0 without ncu:   bash tools/ncu_traffic.sh [tag]
cd "$(dirname "$0")/.."
tag=${1:-r2}
mkdir -p gpurun_out
timeout 600 python tools/gpu_probe.py ncufwd 64 > gpurun_out/${tag}_ncufwd_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_ncufwd_plain.log; exit 1; }
timeout 900 ncu --profile-from-start off --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,lts__t_bytes.sum \
    --csv --log-file gpurun_out/${tag}_traffic.csv python tools/gpu_probe.py ncufwd 64 > gpurun_out/${tag}_ncufwd_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_traffic.csv gpurun_out/${tag}_traffic_summary
