mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_sblk.py tests/test_gpu_groups.py -x -q ) > gpurun_out/c5_tests.log 2>&1
( SSQB_SBLK_TWP=1 timeout 600 python -m pytest tests/test_gpu_sblk.py -x -q ) > gpurun_out/c5_tests_twp.log 2>&1
SSQB_SBLK_TWP=0 python tools/time_groups.py 160000 300 float32 gmw 64 8,16 > gpurun_out/c5_b64_twp0.txt 2>&1
SSQB_SBLK_TWP=1 python tools/time_groups.py 160000 300 float32 gmw 64 8 > gpurun_out/c5_b64_twp1.txt 2>&1
SSQB_LANES=0 python tools/time_groups.py 160000 300 float32 gmw 64 8 > gpurun_out/c5_b64_nolanes.txt 2>&1
SSQB_SBLK_TWP=0 python tools/time_groups.py 160000 300 float32 gmw 8 4 > gpurun_out/c5_b8_twp0.txt 2>&1
SSQB_SBLK_TWP=1 python tools/time_groups.py 160000 300 float32 gmw 8 4 > gpurun_out/c5_b8_twp1.txt 2>&1
SSQB_SBLK_TWP=1 python tools/time_groups.py 1048576 512 float64 gmw 2 1 > gpurun_out/c5_c5_twp1.txt 2>&1
SSQB_SBLK_TWP=0 python tools/time_groups.py 1048576 512 float64 gmw 2 1 > gpurun_out/c5_c5_twp0.txt 2>&1
for f in gpurun_out/c5_tests.log gpurun_out/c5_tests_twp.log; do tail -n 2 $f; done; cat gpurun_out/c5_b*.txt gpurun_out/c5_c5*.txt
