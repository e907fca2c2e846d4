# -*- coding: utf-8 -*-
"""Short-block rows (csrc/cwt_sblk.cuh): small scales run as overlap-save blocks of 4096 /
2048 samples, and scales whose spectrum is CUT at Nyquist (scale * pi inside the wavelet's
support; the reference samples psih on [0, pi] only, wavelets.py:86-95, 473-484) are factored
into the analytic part of the signal times a smooth tapered spectrum.  Checked against the
oracle's whole-signal transform (_cwt.py:167-177) row by row, for every padding rule the block
loader re-implements, batches, and signal lengths that put the block grid at odd offsets."""
import numpy as np
import pytest

from conftest import relerr
from oracle import ssq_oracle as O

pytestmark = pytest.mark.gpu
TOL = {'float32': 2e-6, 'float64': 2e-13}
EPS = {'float32': 1.2e-7, 'float64': 2.3e-16}


@pytest.fixture(scope='module')
def S():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ssqueezepy_b200 as S_
    return S_


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)


def _rows_close(got, ref, dtype):
    """per-row bound of test_gpu_shapes: TOL * |row| + 10 eps * max |row|"""
    rn = np.sqrt((np.abs(ref) ** 2).sum(-1))
    err = np.sqrt((np.abs(got - ref) ** 2).sum(-1))
    bound = TOL[dtype] * rn + 10 * EPS[dtype] * rn.max()
    bad = np.flatnonzero(err > bound)
    assert bad.size == 0, (bad[:8], (err / np.maximum(rn, 1e-300))[bad[:8]])


# log-spaced (the reference accepts only linear / exponential scale arrays): 0.42 .. 1.0 cut at
# Nyquist (gmw(12, 3) has support up to w ~ 3.3 / 4.0), 1.4 .. 9 smooth and short, the largest
# ones long -- they stay on the other routes
SCALES = 2 ** np.linspace(-1.25, 5.35, 12)


@pytest.mark.parametrize('name', ['gmw', 'morlet'])
@pytest.mark.parametrize('dtype', ['float32', 'float64'])
@pytest.mark.parametrize('N,padtype', [(20000, 'reflect'), (23456, 'zero'), (17001, 'symmetric'),
                                       (40000, 'replicate'), (16500, 'wrap')])
def test_cwt_short_and_cut_rows(S, name, dtype, N, padtype):
    opts = {'dtype': dtype}
    okw = {}
    if name == 'gmw':
        opts.update(beta=12, gamma=3); okw = dict(beta=12, gamma=3)
    wav = S.Wavelet((name, opts)); owav = O.OracleWavelet(name, dtype, **okw)
    scales = SCALES if name == 'gmw' else SCALES * 4.2      # morlet(mu = 13.4): support near 13.4 / scale
    rng = np.random.default_rng(N)
    x = (O.chirp(N, 1, dtype) + 0.1 * rng.standard_normal(N)).astype(dtype)
    Wr, _, dWr = O.cwt(x, owav, scales, padtype=padtype)
    Wx, sc, dWx = S.cwt(x, wav, scales=scales, padtype=padtype, derivative=True)
    _rows_close(_np(Wx), Wr, dtype)
    _rows_close(_np(dWx), dWr, dtype)


def test_cut_rows_batch_and_time_support_sign(S):
    """The Python side marks Nyquist-cut scales with a NEGATIVE time support (include/ssq_b200.h,
    tsupport_host) and the batch dimension goes through the same block grid."""
    from ssqueezepy_b200._cwt import _time_supports
    wav = S.Wavelet(('gmw', {'beta': 12, 'gamma': 3, 'dtype': 'float32'}))
    ts = _time_supports(wav, SCALES.astype('float32'))
    ncut = int((SCALES * np.pi <= 3.29).sum())
    assert ncut >= 3 and (ts[:ncut] < 0).all() and (ts[ncut + 1:] > 0).all()
    assert (np.abs(ts[:ncut]) < 200).all()
    owav = O.OracleWavelet('gmw', 'float32', beta=12, gamma=3)
    N = 30000
    x = np.stack([O.chirp(N, b, 'float32') for b in range(3)])
    Wx, _ = S.cwt(x, wav, scales=SCALES, padtype='reflect')
    for b in range(3):
        Wr, _ = O.cwt(x[b], owav, SCALES, padtype='reflect')[:2]
        _rows_close(_np(Wx)[b], Wr, 'float32')


@pytest.mark.parametrize('dtype', ['float32', 'float64'])
def test_ssq_cwt_cut_rows_reassignment(S, dtype):
    """Tx = the oracle's reassignment of the CUDA (Wx, dWx) -- every point in the same bin -- on
    scales that take the short-block route (the fused epilogue is the same on every route)."""
    wav = S.Wavelet(('gmw', {'beta': 12, 'gamma': 3, 'dtype': dtype}))
    owav = O.OracleWavelet('gmw', dtype, beta=12, gamma=3)
    N = 20000
    scales = 2 ** (np.arange(-10, 30) / 8.)                 # 0.42 .. 12.3, log-spaced
    x = np.stack([O.chirp(N, b, dtype) for b in range(2)])
    Tx, Wx, freqs, sc, dWx = S.ssq_cwt(x, wav, scales=scales, get_dWx=True)
    To, Wo, fo, so, dWo = O.ssq_cwt(x, owav, scales, get_dWx=True)
    assert np.array_equal(np.asarray(freqs), fo)
    _rows_close(_np(Wx)[0], Wo[0], dtype)
    st, nv = O.infer_scaletype(_np(sc))
    const = O.cwt_const(_np(sc), st, nv)
    gamma = 10 * (O.EPS64 if dtype == 'float64' else O.EPS32)
    for b in range(2):
        Tref = O.ssqueeze_fused(_np(Wx[b]), _np(dWx[b]), fo[::-1], const, True, True, gamma)
        # same bins (identical non-zero pattern); the sums inside a bin are accumulated by
        # atomics in no fixed order
        assert np.array_equal(_np(Tx[b]) != 0, Tref != 0)
        assert relerr(_np(Tx[b]), Tref) < (2e-6 if dtype == 'float32' else 1e-14)
