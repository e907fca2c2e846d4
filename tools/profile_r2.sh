#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU): per-kind CUDA-event timing, ncu launch
# list of one C2 step, full ncu capture of the hot kernels.  Usage: bash tools/profile_r2.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out
python tools/profile_kinds.py 160000 300 float32 morlet 1 > gpurun_out/${TAG}_kinds_c2.txt 2>&1
python tools/profile_kinds.py 160000 300 float32 gmw 8 > gpurun_out/${TAG}_kinds_c4b8.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/profile_kinds.py 160000 300 float32 morlet 1 \
    > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"grid_interp|cwt_rows_kernel|cwt_pass1f|grid_dec" \
    -s 40 -c 16 -o gpurun_out/${TAG}_hot python tools/profile_kinds.py 160000 300 float32 morlet 1 \
    > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/
