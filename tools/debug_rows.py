"""Per-row error of cwt() against the oracle with the routing-relevant quantities of each row
(band length L, time support S); run under gpurun.  Usage: python tools/debug_rows.py N na wavelet [dtype]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ssqueezepy_b200 as S
from ssqueezepy_b200._cwt import _band_limits, _time_supports
from ssqueezepy_b200.utils import p2up
from oracle import ssq_oracle as O

N, na, name = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
dtype = sys.argv[4] if len(sys.argv) > 4 else 'float32'
opts = {'dtype': dtype}
if name == 'gmw':
    opts.update(beta=12, gamma=3)
wav = S.Wavelet((name, opts))
ow = O.OracleWavelet(name, dtype, **{k: v for k, v in opts.items() if k != 'dtype'})
scales = O.bench_scales(ow, N, na)
x = O.chirp(N, 0, dtype)
n_up, n1, _ = p2up(N)
lo, ln = _band_limits(wav, np.asarray(scales, dtype=dtype), n_up)
ts = _time_supports(wav, np.asarray(scales, dtype=dtype))
Wo, _, dWo = O.cwt(x, ow, scales)
Wx, sc, dWx = S.cwt(x, wav, scales=scales, derivative=True)
Wx, dWx = Wx.cpu().numpy(), dWx.cpu().numpy()
nr = np.linalg.norm(Wo, axis=1); nd = np.linalg.norm(dWo, axis=1)
eW = np.linalg.norm(Wx - Wo, axis=1) / (nr + 1e-6 * nr.max())
eD = np.linalg.norm(dWx - dWo, axis=1) / (nd + 1e-6 * nd.max())
P = 4096; ratio = n_up // P
print("N=%d n_up=%d na=%d %s %s" % (N, n_up, na, name, dtype))
bad = 0
for a in range(na):
    flag = '' if max(eW[a], eD[a]) < 1e-5 else '  <-- BAD'
    bad += bool(flag)
    if flag or a % 10 == 0:
        print("row %3d L=%7d S=%7d  Lb~%5d  eW %.2e eD %.2e%s" % (a, ln[a], ts[a], ln[a] // ratio + 3, eW[a], eD[a], flag))
print("bad rows:", bad)
