mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_tests.log 2>&1
python tools/profile_kinds.py 160000 300 float32 gmw 8 > gpurun_out/c1_kinds_c4b8.txt 2>&1
python tools/profile_kinds.py 160000 300 float32 morlet 1 > gpurun_out/c1_kinds_c2.txt 2>&1
python tools/profile_kinds.py 1048576 512 float64 gmw 1 > gpurun_out/c1_kinds_c5.txt 2>&1
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/c1_bench_n1.json 2> gpurun_out/c1_bench_n1.err
( time python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/c1_bench_ref.json 2> gpurun_out/c1_bench_ref.err
tail -3 gpurun_out/c1_tests.log; cat gpurun_out/c1_kinds_c4b8.txt
