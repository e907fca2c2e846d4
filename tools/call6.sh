mkdir -p gpurun_out
( SSQB_INTERP_PPK=8 timeout 600 python -m pytest tests/test_gpu_shapes.py -x -q -k "single_signal or C4" ) > gpurun_out/c6_tests_ppk8.log 2>&1
SSQB_INTERP_PPK=8 python tools/time_groups.py 160000 300 float32 gmw 64 8 > gpurun_out/c6_b64_ppk8.txt 2>&1
SSQB_INTERP_PPK=4 python tools/time_groups.py 160000 300 float32 gmw 64 8 > gpurun_out/c6_b64_ppk4.txt 2>&1
SSQB_INTERP_PPK=8 python tools/time_groups.py 160000 300 float32 gmw 8 4 > gpurun_out/c6_b8_ppk8.txt 2>&1
SSQB_INTERP_PPK=4 python tools/time_groups.py 160000 300 float32 gmw 8 4 > gpurun_out/c6_b8_ppk4.txt 2>&1
SSQB_INTERP_PPK=8 python tools/time_groups.py 160000 300 float32 morlet 1 0 > gpurun_out/c6_c2_ppk8.txt 2>&1
SSQB_INTERP_PPK=4 python tools/time_groups.py 160000 300 float32 morlet 1 0 > gpurun_out/c6_c2_ppk4.txt 2>&1
tail -n 2 gpurun_out/c6_tests_ppk8.log; cat gpurun_out/c6_b*.txt gpurun_out/c6_c2*.txt
