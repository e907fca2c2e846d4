#!/bin/bash
# Final profiling recipe of round 2 (run under gpurun, one GPU): per-class CUDA-event timing, ncu
# launch list of one step, whole-step DRAM traffic (16 signals in groups of 8: the launches of the
# timed run), full ncu capture of the hot kernels.  Usage: bash tools/profile_final.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out
python tools/profile_kinds.py 160000 300 float32 morlet 1 > gpurun_out/${TAG}_kinds_c2.txt 2>&1
python tools/profile_kinds.py 160000 300 float32 gmw 8 > gpurun_out/${TAG}_kinds_c4b8.txt 2>&1
python tools/profile_kinds.py 1048576 512 float64 gmw 1 > gpurun_out/${TAG}_kinds_c5.txt 2>&1
export SSQB_GROUP=8
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/traffic_step.py 16 > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --profile-from-start off --cache-control none --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv \
    --log-file gpurun_out/${TAG}_traffic_raw.csv python tools/traffic_step.py 16 > gpurun_out/${TAG}_traffic.log 2>&1
python tools/traffic_sum.py gpurun_out/${TAG}_traffic_raw.csv 16 gpurun_out/${TAG}_traffic.json > gpurun_out/${TAG}_traffic_sum.txt 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"grid_interp|sblk_rows|grid_dec|sblk_fwd" -c 12 \
    -o gpurun_out/${TAG}_hot python tools/traffic_step.py 16 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
