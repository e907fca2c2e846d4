"""Time extract_ridges on ssq_cwt planes that stay on the device (B signals of 160 000 samples, 300 scales)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ssqueezepy_b200 as S
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = bench.Workload('C4', B, 0)
w.step(); torch.cuda.synchronize()
for nr in (1, 2):
    S.extract_ridges(w.Tx, w.ssq_freqs, penalty=2., n_ridges=nr, bw=4); torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx = S.extract_ridges(w.Tx, w.ssq_freqs, penalty=2., n_ridges=nr, bw=4); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("extract_ridges B=%d n_ridges=%d: %.1f ms  (%.2f us per time step per ridge)" % (B, nr, dt * 1e3, dt * 1e6 / 160000 / nr))
