mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_sblk.py tests/test_gpu_groups.py -x -q ) > gpurun_out/c4_tests.log 2>&1
( SSQB_RIDGE_CS=1 timeout 300 python -m pytest tests/test_ridges.py -x -q -m gpu ) > gpurun_out/c4_tests_ridge1.log 2>&1
( timeout 300 python -m pytest tests/test_ridges.py -x -q -m gpu ) > gpurun_out/c4_tests_ridge8.log 2>&1
SSQB_SBLK_PREF=1 python tools/time_groups.py 160000 300 float32 gmw 64 8,0 > gpurun_out/c4_b64_pref1.txt 2>&1
SSQB_SBLK_PREF=0 python tools/time_groups.py 160000 300 float32 gmw 64 8 > gpurun_out/c4_b64_pref0.txt 2>&1
SSQB_SBLK_PREF=1 python tools/time_groups.py 160000 300 float32 gmw 8 4,0 > gpurun_out/c4_b8_pref1.txt 2>&1
SSQB_SBLK_PREF=0 python tools/time_groups.py 160000 300 float32 gmw 8 4 > gpurun_out/c4_b8_pref0.txt 2>&1
SSQB_SBLK_PREF=1 python tools/time_groups.py 160000 300 float32 morlet 1 0 > gpurun_out/c4_c2_pref1.txt 2>&1
SSQB_SBLK_PREF=1 python tools/time_groups.py 1048576 512 float64 gmw 2 1 > gpurun_out/c4_c5_pref1.txt 2>&1
SSQB_SBLK_PREF=0 python tools/time_groups.py 1048576 512 float64 gmw 2 1 > gpurun_out/c4_c5_pref0.txt 2>&1
SSQB_RIDGE_CS=1 python tools/time_ridges.py 8 > gpurun_out/c4_ridges_b8_cs1.txt 2>&1
SSQB_RIDGE_CS=8 python tools/time_ridges.py 8 > gpurun_out/c4_ridges_b8_cs8.txt 2>&1
python tools/time_ridges.py 32 > gpurun_out/c4_ridges_b32.txt 2>&1
tail -3 gpurun_out/c4_tests.log gpurun_out/c4_tests_ridge1.log gpurun_out/c4_tests_ridge8.log; cat gpurun_out/c4_b*.txt gpurun_out/c4_c*.txt gpurun_out/c4_ridges*.txt
