mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_groups.py tests/test_gpu_sblk.py::test_ssq_cwt_cut_rows_reassignment -x -q ) > gpurun_out/c3_tests.log 2>&1
python tools/time_groups.py 160000 300 float32 gmw 8 0,4,2,1 > gpurun_out/c3_groups_b8.txt 2>&1
SSQB_ZERO_CTAS=4 python tools/time_groups.py 160000 300 float32 gmw 8 0,2 > gpurun_out/c3_groups_b8_z4.txt 2>&1
SSQB_ZERO_CTAS=2 python tools/time_groups.py 160000 300 float32 gmw 8 0,2 > gpurun_out/c3_groups_b8_z2.txt 2>&1
python tools/time_groups.py 160000 300 float32 gmw 64 0,16,8,4,2 > gpurun_out/c3_groups_b64.txt 2>&1
python tools/time_groups.py 160000 300 float32 morlet 1 0 > gpurun_out/c3_c2.txt 2>&1
python tools/time_groups.py 1048576 512 float64 gmw 2 0,1 > gpurun_out/c3_c5.txt 2>&1
tail -3 gpurun_out/c3_tests.log; cat gpurun_out/c3_groups_b8.txt gpurun_out/c3_groups_b8_z4.txt gpurun_out/c3_groups_b8_z2.txt gpurun_out/c3_groups_b64.txt gpurun_out/c3_c2.txt gpurun_out/c3_c5.txt
