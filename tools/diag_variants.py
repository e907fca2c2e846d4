"""Which rows of cwt(x) differ between derivative=False and derivative=True (must be none)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ssqueezepy_b200 as S
from ssqueezepy_b200._cwt import _band_limits, _time_supports
from oracle import ssq_oracle as O
dtype = sys.argv[1] if len(sys.argv) > 1 else 'float64'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
wav = S.Wavelet(('morlet', {'dtype': dtype})); owav = O.OracleWavelet('morlet', dtype)
na = 80
scales = O.bench_scales(owav, N, na)
x = O.chirp(N, 3, dtype)
W0, sc = S.cwt(x, wav, scales=scales, fs=2.)
W1, _, dW1 = S.cwt(x, wav, scales=scales, fs=2., derivative=True)
W0b, _ = S.cwt(x, wav, scales=scales, fs=2.)
W0, W1, W0b = [t.cpu().numpy() for t in (W0, W1, W0b)]
n_up = S.utils.p2up(N)[0]
lo, ln = _band_limits(wav, np.asarray(scales, dtype=dtype), n_up)
ts = _time_supports(wav, np.asarray(scales, dtype=dtype))
print("n_up", n_up, "repeatable W0:", np.array_equal(W0, W0b))
for a in range(na):
    d = np.abs(W0[a] - W1[a]); nz = np.flatnonzero(d)
    if nz.size:
        print("row %2d scale %9.3f L %6d ts %6d: %6d points differ, max %.3e (|row| max %.3e), first at %d last at %d, stride hist %s"
              % (a, scales[a], ln[a], ts[a], nz.size, d.max(), np.abs(W0[a]).max(), nz[0], nz[-1],
                 np.unique(np.diff(nz))[:6]))
