mkdir -p gpurun_out
( SSQB_INTERP_PPK=16 timeout 600 python -m pytest tests/test_gpu_shapes.py -x -q -k "single_signal or C4" ) > gpurun_out/c7_tests_ppk16.log 2>&1
( SSQB_INTERP_PPK64=4 timeout 600 python -m pytest tests/test_gpu_shapes.py -x -q -k "C5" ) > gpurun_out/c7_tests_ppk64.log 2>&1
SSQB_INTERP_PPK=16 python tools/time_groups.py 160000 300 float32 gmw 64 8 > gpurun_out/c7_b64_ppk16.txt 2>&1
SSQB_INTERP_PPK=8 python tools/time_groups.py 160000 300 float32 gmw 64 8 > gpurun_out/c7_b64_ppk8.txt 2>&1
SSQB_INTERP_PPK=16 python tools/time_groups.py 160000 300 float32 gmw 8 4 > gpurun_out/c7_b8_ppk16.txt 2>&1
SSQB_INTERP_PPK=8 python tools/time_groups.py 160000 300 float32 gmw 8 4 > gpurun_out/c7_b8_ppk8.txt 2>&1
SSQB_INTERP_PPK=16 python tools/time_groups.py 160000 300 float32 morlet 1 0 > gpurun_out/c7_c2_ppk16.txt 2>&1
SSQB_INTERP_PPK=16 python tools/time_groups.py 10000 300 float32 gmw 1 0 > gpurun_out/c7_c1_ppk16.txt 2>&1
SSQB_INTERP_PPK=8 python tools/time_groups.py 10000 300 float32 gmw 1 0 > gpurun_out/c7_c1_ppk8.txt 2>&1
SSQB_INTERP_PPK64=4 python tools/time_groups.py 1048576 512 float64 gmw 2 1 > gpurun_out/c7_c5_ppk4.txt 2>&1
SSQB_INTERP_PPK64=2 python tools/time_groups.py 1048576 512 float64 gmw 2 1 > gpurun_out/c7_c5_ppk2.txt 2>&1
for f in gpurun_out/c7_tests_ppk16.log gpurun_out/c7_tests_ppk64.log; do tail -n 2 $f; done
for f in gpurun_out/c7_b*.txt gpurun_out/c7_c*.txt; do echo "$f: $(cat $f)"; done
