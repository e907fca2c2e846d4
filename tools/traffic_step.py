"""One ssq_cwt step of the bench workload (GMW(12,3), 300 scales, N = 160 000, B signals) between
cudaProfilerStart/Stop, for a whole-step DRAM-traffic capture:

  ncu --profile-from-start off --cache-control none --clock-control none \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --csv --log-file gpurun_out/r2_traffic_raw.csv python tools/traffic_step.py 8

With `--cache-control none` and single-pass metrics every kernel runs once, in order, with the
caches in their natural state: the write-backs of one kernel's dirty lines are counted in the
kernels that evict them, so the SUM over the step is the step's DRAM traffic (up to the
<= 126 MB still dirty in L2 at the end; B = 8 moves 6.1 GB).  `tools/traffic_sum.py` adds it up."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = sys.argv[2] if len(sys.argv) > 2 else 'C4'
w = bench.Workload(cfg, B, 0)
for _ in range(3):
    w.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
w.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("step done: B=%d algorithmic bytes %.1f MB" % (B, w.bytes_per_step / 1e6))
