"""Time the BASELINE.json configurations (other than the bench workload) through the
public API on one GPU.  Usage: python tools/time_cases.py [case ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_diag as D

def _c3():
    import torch, ssqueezepy_b200 as S
    from oracle import ssq_oracle as O
    x = torch.as_tensor(O.chirp(160000), device='cuda')
    for _ in range(3):
        S.ssq_stft(x, n_fft=512, hop_len=128, dtype='float32')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        S.ssq_stft(x, n_fft=512, hop_len=128, dtype='float32')
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    return "%.3f ms/call  %.1f Msamples/s" % (ms, 160000 / ms / 1e3)


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
