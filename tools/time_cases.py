"""Time the BASELINE.json configurations (other than the bench workload) through the
public API on one GPU.  Usage: python tools/time_cases.py [case ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_diag as D

def _c3():
    """C3 (ssq_stft, n_fft 512, hop 128, N = 160 000: 5.8 MB of outputs): a launch-latency-bound
    call; reports the wall time per call of a back-to-back loop through the public API (host
    cost, the GPU work overlaps the next call's host work), the same with a synchronise after
    every call (latency), and the device time of one call (CUDA events)."""
    import time, torch, ssqueezepy_b200 as S
    from oracle import ssq_oracle as O
    x = torch.as_tensor(O.chirp(160000), device='cuda')
    f = lambda: S.ssq_stft(x, n_fft=512, hop_len=128, dtype='float32')
    for _ in range(20):
        f()
    torch.cuda.synchronize()
    n = 500
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e6
    t0 = time.perf_counter()
    for _ in range(200):
        f(); torch.cuda.synchronize()
    lat = (time.perf_counter() - t0) / 200 * 1e6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev = []
    for _ in range(20):
        torch.cuda.synchronize(); e0.record(); f(); e1.record(); torch.cuda.synchronize()
        dev.append(e0.elapsed_time(e1) * 1e3)
    return ("%.1f us/call back-to-back (%.0f Msamples/s), %.1f us/call with a sync after each, "
            "device %.1f us (min of 20)" % (wall, 160000 / wall, lat, min(dev)))


CASES = {
    'C3':    _c3,
    'C1':    lambda: D.timing(10000, 300, 1, 'float32', 'gmw', iters=20),
    'C2':    lambda: D.timing(160000, 300, 1, 'float32', 'morlet', iters=20),
    'C2f64': lambda: D.timing(160000, 300, 1, 'float64', 'morlet', iters=5),
    'C4':    lambda: D.timing(160000, 300, 8, 'float32', 'gmw', iters=3),
    'C5':    lambda: D.timing(1 << 20, 512, 1, 'float64', 'gmw', iters=2),
    'C5f32': lambda: D.timing(1 << 20, 512, 1, 'float32', 'gmw', iters=3),
    'L18':   lambda: D.timing(1 << 18, 300, 1, 'float32', 'morlet', iters=5),
    'L19':   lambda: D.timing(1 << 19, 300, 1, 'float32', 'morlet', iters=5),
}

if __name__ == '__main__':
    for name in (sys.argv[1:] or list(CASES)):
        D.stage(name, CASES[name])
