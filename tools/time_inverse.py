"""Time the inverse reductions at the BASELINE sizes (C2 plane, C3 STFT) on one GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ssqueezepy_b200 as S
from oracle import ssq_oracle as O


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


N, na = 160_000, 300
wav = S.Wavelet('morlet')
scales = O.bench_scales(O.OracleWavelet('morlet', 'float32'), N, na)
x = torch.as_tensor(O.chirp(N), device='cuda')
Tx, Wx, _, sc = S.ssq_cwt(x, wav, scales=scales)
S.issq_cwt(Tx, wav); S.icwt(Wx, wav, scales=sc)            # warm the host caches
ms = timeit(lambda: S.issq_cwt(Tx, wav))
print("issq_cwt  [300 x 160000] c64: %.3f ms/call, %.0f GB/s of the %d MB plane"
      % (ms, Tx.numel() * 8 / ms / 1e6, Tx.numel() * 8 >> 20))
ms = timeit(lambda: S.algos.colsum_real(Tx, scale=2.0))
print("colsum_real kernel only     : %.3f ms/call, %.0f GB/s" % (ms, Tx.numel() * 8 / ms / 1e6))
ms = timeit(lambda: S.icwt(Wx, wav, scales=sc))
print("icwt      [300 x 160000] c64: %.3f ms/call" % ms)
Sx = S.stft(x, n_fft=512, hop_len=128, dtype='float32')
ms = timeit(lambda: S.istft(Sx, n_fft=512, hop_len=128, N=N))
print("istft C3 (n_fft 512, hop 128): %.3f ms/call" % ms)
Sx1 = S.stft(x, n_fft=512, hop_len=1, dtype='float32')
ms = timeit(lambda: S.istft(Sx1, n_fft=512, hop_len=1, N=N), iters=5)
print("istft hop 1 (%d MB of Sx)   : %.3f ms/call, %.0f GB/s" % (Sx1.numel() * 8 >> 20, ms, Sx1.numel() * 8 / ms / 1e6))
