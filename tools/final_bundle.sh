#!/bin/bash
# Round-end measurement bundle on one B200: the GPU test suite, smoke(), the bench line with its
# reference arm, the other BASELINE configurations, then the profiling recipe.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
( time python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python tools/time_cases.py C1 C2 C3 C4 C5 > gpurun_out/r2_cases.txt 2>&1
bash tools/profile_final.sh r2 > gpurun_out/f_profile.log 2>&1
tail -n 4 gpurun_out/f_tests.log; cat gpurun_out/f_smoke.log | tail -n 2; cut -c1-300 gpurun_out/r2_bench_n1.json; cat gpurun_out/r2_cases.txt; cat gpurun_out/r2_kinds_c4b8.txt; tail -n 12 gpurun_out/r2_traffic_sum.txt
