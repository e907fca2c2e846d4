#!/bin/bash
# One-shot profiling recipe (run under gpurun): launch list + full capture of the hot kernels.
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --steps 10 --warmup 3 --batch 8 > gpurun_out/bench_r1_b8.json 2>> gpurun_out/bench_r1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cwt_rows_kernel|cwt_pass1f" -c 12 \
    -o gpurun_out/r1_hot python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
