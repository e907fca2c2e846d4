# -*- coding: utf-8 -*-
"""Transform lengths that are not powers of two (csrc/gfft.cuh: mixed-radix Stockham in
shared memory, two-pass splitting, Bluestein): `padtype=None` on any signal length
(ssqueezepy/utils/common.py:131-156, _cwt.py:261-271) and STFT frames of any `n_fft`
(the reference's own benchmark uses 598 = 2 * 13 * 23, examples/benchmarks.py:82).
Checked against the oracle (pocketfft, like the reference) and NumPy's float64 FFT."""
import numpy as np
import pytest

from conftest import relerr
from oracle import ssq_oracle as O

pytestmark = pytest.mark.gpu
TOL = {'float32': 1e-5, 'float64': 1e-12}


@pytest.fixture(scope='module')
def S():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ssqueezepy_b200 as S_
    return S_


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)


def _pair(S, dtype):
    return S.Wavelet(('morlet', {'dtype': dtype})), O.OracleWavelet('morlet', dtype)


# 1000 = 2^3 5^3 (one shared-memory transform), 1009 prime (Bluestein, M = 2048),
# 4099 prime (Bluestein through a two-pass power of two), 25000 = 2^3 5^5 and
# 160000 = 2^8 5^4 (two passes, 400 x 400), 30030 = 2 3 5 7 11 13, 598 = 2 13 23
@pytest.mark.parametrize('n,dtype', [(1000, 'float32'), (1009, 'float64'), (4099, 'float32'),
                                     (25000, 'float64'), (160000, 'float32'),
                                     (30030, 'float32'), (598, 'float64'), (97, 'float32')])
def test_forward_fft_any_length(S, n, dtype):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((2, n)).astype(dtype)
    wav, _ = _pair(S, dtype)
    plan = S.CwtPlan.get(wav, np.array([4., 8.]), n, n, 0, 'zero', 1.)
    xh = _np(plan.debug_xh(x))
    ref = np.fft.fft(x.astype(np.float64), axis=-1) / n
    assert relerr(xh, ref) < (3e-6 if dtype == 'float32' else 1e-14)


@pytest.mark.parametrize('N,dtype,na', [(1000, 'float32', 24), (1009, 'float64', 16),
                                        (25000, 'float32', 16), (160000, 'float32', 6)])
def test_cwt_padtype_none_any_length(S, N, dtype, na):
    wav, owav = _pair(S, dtype)
    scales = 2 ** np.linspace(2.5, 7.5, na)
    x = O.chirp(N, 2, dtype)
    Wr, _, dWr = O.cwt(x, owav, scales, padtype=None)
    Wx, sc, dWx = S.cwt(x, wav, scales=scales, padtype=None, derivative=True)
    assert tuple(Wx.shape) == (na, N)
    assert relerr(_np(Wx), Wr) < TOL[dtype]
    assert relerr(_np(dWx), dWr) < TOL[dtype]
    W2, _ = S.cwt(x, wav, scales=scales, padtype=None, l1_norm=False)
    assert relerr(_np(W2), _np(Wx) * np.sqrt(_np(sc).astype(dtype))[:, None]) < 1e-6


@pytest.mark.parametrize('N,dtype', [(3000, 'float32'), (1250, 'float64')])
def test_ssq_cwt_padtype_none(S, N, dtype):
    """Same column-owner operator as `ssqueeze`: Tx equals the oracle's ordered reassignment of
    the CUDA (Wx, dWx) bit for bit."""
    wav, owav = _pair(S, dtype)
    na = 40
    scales = 2 ** np.linspace(2.5, 7.0, na)
    x = np.stack([O.chirp(N, b, dtype) for b in range(2)])
    Tx, Wx, freqs, sc, dWx = S.ssq_cwt(x, wav, scales=scales, padtype=None, get_dWx=True)
    To, Wo, fo, so, dWo = O.ssq_cwt(x, owav, scales, padtype=None, get_dWx=True)
    assert np.array_equal(np.asarray(freqs), fo)
    assert relerr(_np(Wx), Wo) < TOL[dtype]
    st, nv = O.infer_scaletype(_np(sc))
    const = O.cwt_const(_np(sc), st, nv)
    gamma = 10 * (O.EPS64 if dtype == 'float64' else O.EPS32)
    for b in range(2):
        Tref = O.ssqueeze_fused(_np(Wx[b]), _np(dWx[b]), fo[::-1], const, True, True, gamma)
        assert np.array_equal(_np(Tx[b]), Tref)
        assert relerr(_np(Tx[b]).sum(0), To[b].sum(0)) < (5e-5 if dtype == 'float32' else 1e-11)


@pytest.mark.parametrize('n_fft,hop,dtype', [(598, 128, 'float32'), (97, 5, 'float64'),
                                            (1031, 64, 'float32'), (600, 7, 'float64')])
def test_stft_any_n_fft(S, n_fft, hop, dtype):
    N = 20_000
    x = O.chirp(N, 1, dtype)
    Sr, dSr = O.stft(x, None, n_fft, None, hop, 1., dtype=dtype)
    Tx, Sx, freqs, Sfs, dSx = S.ssq_stft(x, n_fft=n_fft, hop_len=hop, dtype=dtype, get_dWx=True)
    assert tuple(Sx.shape) == Sr.shape
    assert relerr(_np(Sx), Sr) < TOL[dtype] and relerr(_np(dSx), dSr) < TOL[dtype]
    sfs = _np(Sfs)
    gamma = 10 * (O.EPS64 if dtype == 'float64' else O.EPS32)
    Tref = O.ssqueeze_fused(_np(Sx), _np(dSx), sfs, sfs[1] - sfs[0], False, False, gamma, Sfs=sfs)
    assert relerr(_np(Tx), Tref) < (2e-6 if dtype == 'float32' else 1e-14)
    assert np.array_equal(_np(Tx) != 0, Tref != 0)
