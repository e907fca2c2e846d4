mkdir -p gpurun_out
( SSQB_DEC_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_shapes.py -x -q -k "single_signal or C4" ) > gpurun_out/c8_tests_split.log 2>&1
SSQB_DEC_SPLIT=1 python tools/time_groups.py 160000 300 float32 gmw 64 8 > gpurun_out/c8_b64_split1.txt 2>&1
SSQB_DEC_SPLIT=0 python tools/time_groups.py 160000 300 float32 gmw 64 8 > gpurun_out/c8_b64_split0.txt 2>&1
SSQB_DEC_SPLIT=1 python tools/time_groups.py 160000 300 float32 gmw 8 4 > gpurun_out/c8_b8_split1.txt 2>&1
SSQB_DEC_SPLIT=0 python tools/time_groups.py 160000 300 float32 gmw 8 4 > gpurun_out/c8_b8_split0.txt 2>&1
SSQB_DEC_SPLIT=1 python tools/time_groups.py 160000 300 float32 morlet 1 0 > gpurun_out/c8_c2_split1.txt 2>&1
SSQB_DEC_SPLIT=0 python tools/time_groups.py 160000 300 float32 morlet 1 0 > gpurun_out/c8_c2_split0.txt 2>&1
tail -n 2 gpurun_out/c8_tests_split.log
for f in gpurun_out/c8_b*.txt gpurun_out/c8_c*.txt; do echo "$f: $(cat $f)"; done
