# -*- coding: utf-8 -*-
"""Inverse transforms (issq_cwt, icwt, istft, issq_stft).

CPU part: the oracle restatement against tests/golden/inverse.npz (outputs of the real
reference on the stored forward fixtures).  GPU part: the CUDA reductions, through the
public API and the C ABI, against the same vectors.  Given identical inputs the column
sums add rows in the reference's order, so `issq_*` / `icwt` agree to the rounding of
the admissibility constant (its float32 `exp` differs between numba and numpy at 3e-9);
`istft` additionally carries the FFT's rounding."""
import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle import ssq_oracle as O

GM = dict(beta=12, gamma=3)


def _owav(tag):
    if 'gmw_f64' in tag:
        return O.OracleWavelet('gmw', 'float64', **GM)
    if 'gmw_f32' in tag:
        return O.OracleWavelet('gmw', 'float32', **GM)
    if 'piecewise' in tag:
        return O.OracleWavelet('gmw', 'float32')
    return O.OracleWavelet('morlet', 'float32')


# ---- oracle vs the reference's outputs (CPU) ----------------------------------------
def test_oracle_admissibility_constants():
    adm = load_golden('inverse')['adm']
    got = [O.adm_ssq(O.OracleWavelet('morlet')), O.adm_ssq(O.OracleWavelet('gmw', **GM)),
           O.adm_ssq(O.OracleWavelet('gmw', 'float64', **GM))]
    assert np.allclose(got, adm[:3], rtol=2e-8, atol=0)
    assert abs(got[2] / adm[2] - 1) < 1e-13          # float64 wavelet: same doubles


def test_oracle_issq_cwt_and_components():
    inv = load_golden('inverse')
    g = load_golden('cwt_morlet_f32')
    x = O.issq_cwt(g['Tx'], _owav('morlet'))
    assert x.dtype == inv['issq_morlet_f32'].dtype == np.float32
    assert relerr(x, inv['issq_morlet_f32']) < 2e-7
    xc = O.issq_cwt(g['Tx'], _owav('morlet'), inv['cc'], inv['cw'])
    assert xc.shape == inv['issq_morlet_f32_comp'].shape == (4, g['Tx'].shape[1])
    assert relerr(xc, inv['issq_morlet_f32_comp']) < 2e-7
    g = load_golden('cwt_gmw_f64')
    assert relerr(O.issq_cwt(g['Tx'], _owav('gmw_f64')), inv['issq_gmw_f64']) < 1e-13


@pytest.mark.parametrize('tag,key,kw', [
    ('cwt_morlet_f32', 'Wx', {}), ('cwt_morlet_f32', 'Wx_l2', dict(l1_norm=False, x_mean=0.25)),
    ('cwt_gmw_f64', 'Wx', {}), ('cwt_lin_f32', 'Wx', {}),
    ('cwt_piecewise_f32', 'Wx', dict(x_mean=0.5)), ('cwt_gmw_f32_batch', 'Wx', {})])
def test_oracle_icwt(tag, key, kw):
    inv = load_golden('inverse')
    g = load_golden(tag)
    ref = inv['icwt_' + tag[4:] + ('_l2' if key == 'Wx_l2' else '')]
    x = O.icwt(g[key], _owav(tag), g['scales_in'], **kw)
    assert x.dtype == ref.dtype and x.shape == ref.shape
    assert relerr(x, ref) < (1e-13 if 'f64' in tag else 2e-7)


ISTFT_CASES = [
    ('stft_f32', 'istft_f32', dict(n_fft=128, hop_len=16, N=3000)),
    ('stft_f32', 'istft_f32_exp0', dict(n_fft=128, hop_len=16, N=3000, win_exp=0)),
    ('stft_f32', 'istft_f32_defN', dict(n_fft=128, hop_len=16)),
    ('stft_f64_odd', 'istft_f64_odd', dict(n_fft=97, hop_len=5, N=1111)),
    ('stft_f32_batch', 'istft_f32_winlen_b0', dict(n_fft=64, win_len=48, hop_len=8, N=900)),
    ('stft_f32_nomod', 'istft_f32_nomod', dict(n_fft=64, hop_len=8, N=800, modulated=False)),
]


@pytest.mark.parametrize('tag,key,kw', ISTFT_CASES)
def test_oracle_istft(tag, key, kw):
    inv = load_golden('inverse')
    Sx = load_golden(tag)['Sx']
    Sx = Sx[0] if Sx.ndim == 3 else Sx
    x = O.istft(Sx, **kw)
    assert x.dtype == inv[key].dtype and x.shape == inv[key].shape
    assert relerr(x, inv[key]) < (1e-13 if 'f64' in key else 1e-6)
    if key == 'istft_f32':                 # and the round trip itself (reference test
        x0 = load_golden(tag)['x']         # tests/reconstruction_test.py:160-179)
        assert relerr(x, x0) < 2e-3        # right-most hop is imprecise in float32 (NOLA note)
        assert relerr(x[:2900], x0[:2900]) < 1e-5


def test_oracle_issq_stft():
    inv = load_golden('inverse')
    assert relerr(O.issq_stft(inv['sq_Tx'], n_fft=64), inv['issq_stft_f32']) < 2e-7
    xc = O.issq_stft(inv['sq_Tx'], cc=inv['sq_cc'], cw=inv['sq_cw'], n_fft=64)
    assert relerr(xc, inv['issq_stft_f32_comp']) < 2e-7
    assert relerr(O.issq_stft(inv['sq_Tx64'], n_fft=48, win_len=32), inv['issq_stft_f64']) < 1e-13
    # reconstruction quality of the reference's own scheme on this signal
    assert relerr(inv['issq_stft_f32'], inv['sq_x']) < 0.05


# ---- product host logic that needs no GPU -------------------------------------------
def test_product_admissibility_constants():
    import ssqueezepy_b200 as S
    adm = load_golden('inverse')['adm']
    got = [S.adm_ssq('morlet'), S.adm_ssq(('gmw', GM)),
           S.adm_ssq(('gmw', dict(GM, dtype='float64'))), S.adm_cwt('morlet'), S.adm_ssq('gmw')]
    assert np.allclose(got[:4], adm[:4], rtol=2e-8, atol=0)
    assert abs(got[4] / adm[4] - 1) < 1e-6    # beta=60 in float32: libm powf/expf noise
    assert abs(got[2] / adm[2] - 1) < 1e-13


def test_product_inverse_argument_errors():
    import ssqueezepy_b200 as S
    Tx = np.zeros((33, 10), dtype=np.complex64)
    with pytest.raises(ValueError):
        S.issq_stft(Tx, hop_len=2)
    with pytest.raises(ValueError):
        S.issq_stft(Tx, modulated=False)
    with pytest.raises(NotImplementedError):
        S.icwt(Tx, 'morlet', scales=np.arange(1., 34.), one_int=False)


# ---- CUDA path ------------------------------------------------------------------------
@pytest.fixture(scope='module')
def S():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import ssqueezepy_b200 as S_
    return S_


def _pwav(tag):
    if 'gmw_f64' in tag:
        return ('gmw', dict(GM, dtype='float64'))
    if 'gmw_f32' in tag:
        return ('gmw', dict(GM))
    if 'piecewise' in tag:
        return 'gmw'
    return 'morlet'


@pytest.mark.gpu
def test_gpu_issq_cwt(S):
    import torch
    inv = load_golden('inverse')
    g = load_golden('cwt_morlet_f32')
    x = S.issq_cwt(g['Tx'], 'morlet')
    assert isinstance(x, np.ndarray) and x.dtype == np.float32
    assert relerr(x, inv['issq_morlet_f32']) < 2e-7
    # same sums as numpy's, row after row: identical up to the one product with 2 / Css
    Css = S.adm_ssq('morlet')
    want = g['Tx'].real.sum(axis=0)
    want *= (2 / Css)
    assert np.array_equal(x, want)
    xt = S.issq_cwt(torch.as_tensor(g['Tx']).cuda(), 'morlet')
    assert xt.is_cuda and np.array_equal(xt.cpu().numpy(), x)
    xc = S.issq_cwt(g['Tx'], 'morlet', inv['cc'], inv['cw'])
    assert xc.dtype == np.float64 and xc.shape == inv['issq_morlet_f32_comp'].shape
    assert relerr(xc, inv['issq_morlet_f32_comp']) < 2e-8
    assert np.array_equal(xc, O.invert_components(g['Tx'], inv['cc'], inv['cw']) * (2 / Css))
    g = load_golden('cwt_gmw_f64')
    assert relerr(S.issq_cwt(g['Tx'], _pwav('cwt_gmw_f64')), inv['issq_gmw_f64']) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize('tag,key,kw', [
    ('cwt_morlet_f32', 'Wx', {}), ('cwt_morlet_f32', 'Wx_l2', dict(l1_norm=False, x_mean=0.25)),
    ('cwt_gmw_f64', 'Wx', {}), ('cwt_lin_f32', 'Wx', {}),
    ('cwt_piecewise_f32', 'Wx', dict(x_mean=0.5)), ('cwt_gmw_f32_batch', 'Wx', {})])
def test_gpu_icwt(S, tag, key, kw):
    inv = load_golden('inverse')
    g = load_golden(tag)
    ref = inv['icwt_' + tag[4:] + ('_l2' if key == 'Wx_l2' else '')]
    x = S.icwt(g[key], _pwav(tag), scales=g['scales_in'], **kw)
    assert x.dtype == ref.dtype and x.shape == ref.shape
    assert relerr(x, ref) < (1e-13 if 'f64' in tag else 2e-7)


@pytest.mark.gpu
def test_gpu_cwt_icwt_round_trip(S):
    """cwt -> icwt on the device (reference tests/reconstruction_test.py:60-95 design)."""
    N = 4096
    t = np.arange(N) / N
    x = (np.cos(2 * np.pi * 60 * t) + np.cos(2 * np.pi * (100 * t + 40 * t**2))).astype('float32')
    Wx, scales = S.cwt(x, 'gmw', scales='log', nv=32)
    xr = S.icwt(Wx, 'gmw', scales=scales, nv=32)
    assert xr.is_cuda
    xr = xr.cpu().numpy()
    mid = slice(N // 8, -N // 8)
    assert relerr(xr[mid], x[mid]) < 0.03


@pytest.mark.gpu
@pytest.mark.parametrize('tag,key,kw', ISTFT_CASES)
def test_gpu_istft(S, tag, key, kw):
    inv = load_golden('inverse')
    Sx = load_golden(tag)['Sx']
    Sx = Sx[0] if Sx.ndim == 3 else Sx
    x = S.istft(Sx, **kw)
    assert x.dtype == inv[key].dtype and x.shape == inv[key].shape
    assert relerr(x, inv[key]) < (1e-12 if 'f64' in key else 1e-5)
    assert relerr(x, O.istft(Sx, **kw)) < (1e-12 if 'f64' in key else 1e-5)


@pytest.mark.gpu
def test_gpu_stft_istft_round_trip_and_batch(S):
    import torch
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 5000))
    for hop, n_fft in ((1, 128), (32, 256)):
        Sx = S.stft(x, n_fft=n_fft, hop_len=hop, dtype='float64')
        xr = S.istft(Sx, n_fft=n_fft, hop_len=hop, N=5000)
        assert xr.is_cuda and tuple(xr.shape) == (3, 5000)
        assert relerr(xr.cpu().numpy(), x) < 1e-12
        x1 = S.istft(Sx[1], n_fft=n_fft, hop_len=hop, N=5000)
        assert torch.equal(x1, xr[1])


@pytest.mark.gpu
def test_gpu_issq_stft(S):
    inv = load_golden('inverse')
    x = S.issq_stft(inv['sq_Tx'], n_fft=64)
    assert x.dtype == np.float32 and relerr(x, inv['issq_stft_f32']) < 2e-7
    xc = S.issq_stft(inv['sq_Tx'], cc=inv['sq_cc'], cw=inv['sq_cw'], n_fft=64)
    assert relerr(xc, inv['issq_stft_f32_comp']) < 2e-7
    assert relerr(S.issq_stft(inv['sq_Tx64'], n_fft=48, win_len=32), inv['issq_stft_f64']) < 1e-13


@pytest.mark.gpu
def test_gpu_colsum_c_abi_full_size(S):
    """C2-sized plane through the C entry point: column sums vs torch, and the
    flip-invariant identity sum_k Tx[k, j] == const * sum_a active Wx[a, j]."""
    import torch
    N, na = 160_000, 300
    wav, ow = S.Wavelet('morlet'), O.OracleWavelet('morlet', 'float32')
    scales = O.bench_scales(ow, N, na)
    Tx, Wx, _, sc = S.ssq_cwt(O.chirp(N), wav, scales=scales)
    x = S.issq_cwt(Tx, wav)
    ref = Tx.real.double().sum(0) * (2 / S.adm_ssq(wav))
    assert float((x.double() - ref).norm() / ref.norm()) < 1e-6
    xi = S.icwt(Wx, wav, scales=sc)
    assert relerr(xi[N // 8:-N // 8].cpu().numpy(), O.chirp(N)[N // 8:-N // 8]) < 0.02
