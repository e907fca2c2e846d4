mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/c2_launches.csv python tools/traffic_step.py 8 > gpurun_out/c2_ncu_list.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"sblk_rows|grid_interp|cwt_pass1_kernel|cwt_pass2_kernel|sblk_fwd|grid_dec" -c 16 \
    -o gpurun_out/c2_hot python tools/traffic_step.py 8 > gpurun_out/c2_ncu_full.log 2>&1
SSQB_GRAPH=1 python tools/profile_kinds.py 160000 300 float32 gmw 8 2>&1 | head -1 > gpurun_out/c2_graph.txt
SSQB_NO_SBLK=1 python tools/profile_kinds.py 160000 300 float32 gmw 8 > gpurun_out/c2_nosblk.txt 2>&1
SSQB_LANES=0 python tools/profile_kinds.py 160000 300 float32 gmw 8 2>&1 | head -1 > gpurun_out/c2_nolanes.txt
ls -la gpurun_out | tail -8
