mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_multi_gpu.py -q ) > gpurun_out/g2_tests.log 2>&1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --gather ) > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -n 3 gpurun_out/g2_tests.log; cut -c1-700 gpurun_out/r2_bench_n2.json; tail -n 5 gpurun_out/r2_bench_n2.err
