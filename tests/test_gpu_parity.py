# -*- coding: utf-8 -*-
"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the
golden vectors produced by the real reference.  Tolerances:
  Wx, dWx, Sx, dSx : ||d||/||ref|| <= 1e-5 (float32) / 1e-12 (float64)
  reassignment op on identical (Wx, dWx): bit-exact Tx (ordered accumulation)
  fused ssq: bin indices bit-exact w.r.t. the oracle applied to the SAME Wx, dWx;
             Tx <= 1e-6 norm-wise (atomic accumulation order)
"""
import ctypes as C
import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle import ssq_oracle as O

pytestmark = pytest.mark.gpu

TOL = {'float32': 1e-5, 'float64': 1e-12}


@pytest.fixture(scope='module')
def S():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import ssqueezepy_b200 as S_
    return S_


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)


def _wavs(tag, S):
    if 'morlet' in tag or 'lin' in tag:
        return S.Wavelet('morlet'), O.OracleWavelet('morlet', 'float32')
    if 'gmw_f64' in tag:
        return (S.Wavelet(('gmw', {'beta': 12, 'gamma': 3, 'dtype': 'float64'})),
                O.OracleWavelet('gmw', 'float64', beta=12, gamma=3))
    if 'piecewise' in tag:
        return S.Wavelet('gmw'), O.OracleWavelet('gmw', 'float32')
    return (S.Wavelet(('gmw', {'beta': 12, 'gamma': 3})),
            O.OracleWavelet('gmw', 'float32', beta=12, gamma=3))


CWT_CASES = ['cwt_morlet_f32', 'cwt_gmw_f64', 'cwt_gmw_f32_batch', 'cwt_lin_f32',
             'cwt_piecewise_f32']


# ---------------------------------------------------------------------------
@pytest.mark.parametrize('N,dtype', [(8, 'float32'), (100, 'float32'),
                                     (700, 'float64'), (1500, 'float32'),
                                     (3000, 'float32'), (10000, 'float32'),
                                     (10000, 'float64'), (40000, 'float32')])
def test_forward_fft_of_padded_signal(S, N, dtype):
    """_cwt.py:261-269: xh = fft(padsignal(x)) (here stored divided by n_up)."""
    rng = np.random.default_rng(N)
    x = rng.standard_normal((2, N)).astype(dtype)
    wav = S.Wavelet('morlet', dtype=dtype)
    n_up, n1, _ = S.utils.p2up(N)
    plan = S.CwtPlan.get(wav, np.array([4., 8.]), N, n_up, n1, 'reflect', 1.)
    xh = _np(plan.debug_xh(x))
    ref = np.fft.fft(O.padsignal(x.astype(np.float64))[0], axis=-1) / n_up
    assert relerr(xh, ref) < (3e-6 if dtype == 'float32' else 1e-14)


@pytest.mark.parametrize('tag', CWT_CASES)
def test_cwt_matches_reference_golden(S, tag):
    g = load_golden(tag)
    wav, _ = _wavs(tag, S)
    Wx, sc, dWx = S.cwt(g['x'], wav, scales=g['scales_in'], fs=float(g['fs']),
                        derivative=True)
    tol = TOL[wav.dtype]
    if 'piecewise' in tag:
        tol = 2e-5   # GMW beta=60 float32: reference-vs-numpy already differ by 3e-6
    assert tuple(Wx.shape) == g['Wx'].shape
    assert str(Wx.dtype).endswith('complex64' if wav.dtype == 'float32' else 'complex128')
    assert np.array_equal(_np(sc), g['scales_out'])
    assert relerr(_np(Wx), g['Wx']) < tol
    assert relerr(_np(dWx), g['dWx']) < tol


def test_cwt_l2_norm_rpadded_and_padtypes(S):
    g = load_golden('cwt_morlet_f32')
    wav, owav = _wavs('cwt_morlet_f32', S)
    Wl2, _ = S.cwt(g['x'], wav, scales=g['scales_in'], l1_norm=False)
    assert relerr(_np(Wl2), g['Wx_l2']) < 1e-5
    x = g['x'][:777]
    for padtype in ('reflect', 'zero', 'symmetric', 'replicate', 'wrap'):
        Wr, _, dWr = O.cwt(x, owav, g['scales_in'], padtype=padtype, rpadded=True)
        Wp, _, dWp = S.cwt(x, wav, scales=g['scales_in'], padtype=padtype,
                           rpadded=True, derivative=True)
        assert relerr(_np(Wp), Wr) < 1e-5, padtype
        assert relerr(_np(dWp), dWr) < 1e-5, padtype
    # padtype=None on a power-of-two length
    x2 = g['x'][:1024]
    Wr, _, _ = O.cwt(x2, owav, g['scales_in'], padtype=None)
    Wp, _ = S.cwt(x2, wav, scales=g['scales_in'], padtype=None)
    assert relerr(_np(Wp), Wr) < 1e-5
    # ... and on the fixture's own (non power-of-two) length: generic-length FFT
    Wr, _, _ = O.cwt(x, owav, g['scales_in'], padtype=None)
    Wp, _ = S.cwt(x, wav, scales=g['scales_in'], padtype=None)
    assert relerr(_np(Wp), Wr) < 1e-5


def test_cwt_argument_errors(S):
    wav = S.Wavelet('morlet')
    with pytest.raises(TypeError):
        S.cwt([1., 2., 3.], wav)
    with pytest.raises(ValueError):
        S.cwt(np.zeros((2, 3, 4)), wav)
    with pytest.raises(ValueError):
        S.cwt(np.zeros(64), S.Wavelet('gmw'), l1_norm=False)
    with pytest.raises(ValueError):
        S.cwt(np.zeros(64), wav, padtype='bogus')


def test_table_wavelet_equals_builtin(S):
    """A custom FunctionType wavelet goes through the psih-table path; with the
    Morlet formula it must agree with the in-kernel Morlet."""
    g = load_golden('cwt_morlet_f32')
    wav = S.Wavelet('morlet')
    fn = S.wavelets.morlet()
    custom = S.Wavelet(lambda w: fn(w))
    assert custom.device_spec() is None
    Wa, _, dWa = S.cwt(g['x'], wav, scales=g['scales_in'], derivative=True)
    Wb, _, dWb = S.cwt(g['x'], custom, scales=g['scales_in'], derivative=True)
    assert relerr(_np(Wb), _np(Wa)) < 2e-6 and relerr(_np(dWb), _np(dWa)) < 2e-6
    assert relerr(_np(Wb), g['Wx']) < 1e-5


# ---------------------------------------------------------------------------
@pytest.mark.parametrize('dtype', ['float32', 'float64'])
@pytest.mark.parametrize('flipud', [False, True])
def test_ssqueeze_operator_bit_exact(S, dtype, flipud):
    """ssqueeze_fast on the reference's own random inputs: Tx must be identical
    to the reference's CPU result (reference tests/fft_test.py:284-348)."""
    g = load_golden('reassign')
    Wx, dWx = g[f'{dtype}_Wx'], g[f'{dtype}_dWx']
    gamma = 10 * np.finfo(dtype).eps
    tag = f'{dtype}_flip{int(flipud)}'
    carr, Sfs = g[f'{dtype}_const_arr'], g[f'{dtype}_Sfs']
    cases = [('log', g[f'{dtype}_flog'], np.log(2) / 8, True, None),
             ('pw', g[f'{dtype}_fpw'], carr, True, None),
             ('lin', g[f'{dtype}_flin'], carr, False, None),
             ('stft', Sfs, float(Sfs[1] - Sfs[0]), False, Sfs)]
    for name, freqs, const, logscale, sfs in cases:
        Tx = S.ssqueeze_fast(Wx, dWx, freqs, const, logscale, flipud, gamma, Sfs=sfs)
        assert np.array_equal(_np(Tx), g[f'Tx_{name}_{tag}']), name
    w = S.phase_cwt_gpu(Wx, dWx, gamma)
    assert np.array_equal(_np(w), g[f'{dtype}_w_cwt'])
    ws = S.phase_stft_gpu(Wx, dWx, Sfs, gamma)
    assert np.array_equal(_np(ws), g[f'{dtype}_w_stft'])
    for name, freqs, const, logscale in [c[:4] for c in cases[:3]]:
        Ix = _np(S.indexed_sum_onfly(Wx, g[f'{dtype}_w_cwt'], freqs, const, logscale,
                                     flipud))
        ref = g[f'Ix_{name}_{tag}']
        # float32 log grids: the device evaluates glibc's log2f algorithm (the function
        # numba calls), so the bins -- and with ordered accumulation Tx -- are exact
        assert np.array_equal(Ix, ref), name


@pytest.mark.parametrize('tag', CWT_CASES)
def test_ssqueeze_on_reference_cwt_bit_exact(S, tag):
    """Public `ssqueeze` fed the reference's (Wx, dWx): host parameters + kernel
    must reproduce the reference's Tx and ssq_freqs exactly."""
    g = load_golden(tag)
    wav, _ = _wavs(tag, S)
    gamma = 10 * (O.EPS64 if wav.dtype == 'float64' else O.EPS32)
    st = S.utils.infer_scaletype(g['scales_in'])[0]
    Tx, freqs = S.ssqueeze(g['Wx'], None, ssq_freqs=st, scales=g['scales_out'],
                           fs=float(g['fs']), maprange='peak', wavelet=wav,
                           gamma=gamma, flipud=True, dWx=g['dWx'], transform='cwt')
    assert np.array_equal(np.asarray(freqs), g['ssq_freqs'])
    assert np.array_equal(_np(Tx), g['Tx'])


@pytest.mark.parametrize('tag', CWT_CASES)
def test_fused_ssq_cwt(S, tag):
    g = load_golden(tag)
    wav, owav = _wavs(tag, S)
    fs = float(g['fs'])
    Tx, Wx, freqs, sc, dWx = S.ssq_cwt(g['x'], wav, scales=g['scales_in'], fs=fs,
                                       get_dWx=True)
    Tx, Wx, dWx = _np(Tx), _np(Wx), _np(dWx)
    tol = TOL[wav.dtype] if 'piecewise' not in tag else 2e-5
    assert np.array_equal(np.asarray(freqs), g['ssq_freqs'])
    assert np.array_equal(_np(sc), g['scales_out'])
    assert relerr(Wx, g['Wx']) < tol and relerr(dWx, g['dWx']) < tol
    # (2) bin indices / reassignment exact given OUR Wx, dWx: rebuild with the oracle
    N = g['x'].shape[-1]
    st_in = O.infer_scaletype(g['scales_in'])[0]
    st, nv = O.infer_scaletype(g['scales_out'])
    ofreqs = O.ssq_freqs_cwt(g['scales_out'], N, owav, st_in, 'peak', 1 / fs, True)
    const = O.cwt_const(g['scales_out'], st, nv)
    gamma = 10 * (O.EPS64 if wav.dtype == 'float64' else O.EPS32)
    W3 = Wx if Wx.ndim == 3 else Wx[None]
    dW3 = dWx if dWx.ndim == 3 else dWx[None]
    T3 = Tx if Tx.ndim == 3 else Tx[None]
    G3 = g['Tx'] if g['Tx'].ndim == 3 else g['Tx'][None]
    for W, dW, T, Tg in zip(W3, dW3, T3, G3):
        Tref = O.ssqueeze_fused(W, dW, ofreqs, const, st_in.startswith('log'), True,
                                gamma)
        assert relerr(T, Tref) < (2e-6 if wav.dtype == 'float32' else 1e-14)
        # nonzero pattern identical <=> every point landed in the same bin
        assert np.array_equal(T != 0, Tref != 0)
        # (3) flip-invariant check against the reference's own Tx: column sums
        assert relerr(T.sum(0), Tg.sum(0)) < (5e-5 if wav.dtype == 'float32' else 1e-10)


def test_ssq_cwt_two_step_and_get_w(S):
    g = load_golden('cwt_morlet_f32')
    wav, _ = _wavs('cwt_morlet_f32', S)
    out = S.ssq_cwt(g['x'], wav, scales=g['scales_in'], get_w=True, get_dWx=True)
    Tx, Wx, freqs, sc, w, dWx = out
    assert tuple(w.shape) == g['Wx'].shape and str(w.dtype).endswith('float32')
    gamma = 10 * O.EPS32
    wref = O.phase_cwt(_np(Wx), _np(dWx), gamma)
    assert np.array_equal(_np(w), wref)
    with pytest.raises(NotImplementedError):
        S.ssq_cwt(np.zeros((2, 256), dtype='float32'), wav, get_w=True)
    with pytest.raises(ValueError):
        S.ssq_cwt(g['x'], wav, difftype='phase', get_w=True)
    # astensor=False returns numpy with the reference's dtypes (z_all_test.py:383-413)
    Tn, Wn, fn_, sn = S.ssq_cwt(g['x'], wav, scales=g['scales_in'], astensor=False)
    assert isinstance(Tn, np.ndarray) and Tn.dtype == np.complex64
    assert Wn.dtype == np.complex64 and sn.dtype == np.float32 and fn_.dtype == np.float64


def test_default_arguments_ssq_cwt(S):
    """`ssq_cwt(x)` with every default (GMW beta=60, 'log-piecewise' scales)."""
    g = load_golden('cwt_piecewise_f32')
    x = g['x']
    Tx, Wx, freqs, sc = S.ssq_cwt(x, nv=16)
    assert np.array_equal(_np(sc), g['scales_out'])
    assert np.array_equal(np.asarray(freqs), g['ssq_freqs'])
    assert relerr(_np(Wx), g['Wx']) < 2e-5


# ---------------------------------------------------------------------------
STFT_CASES = ['stft_f32', 'stft_f64_odd', 'stft_f32_batch', 'stft_f32_nomod']


@pytest.mark.parametrize('tag', STFT_CASES)
def test_stft_and_ssq_stft(S, tag):
    g = load_golden(tag)
    dtype = str(g['x'].dtype)
    n_fft, hop, fs = int(g['n_fft']), int(g['hop']), float(g['fs'])
    win_len, mod = int(g['win_len']), bool(g['modulated'])
    Sx, dSx = S.stft(g['x'], n_fft=n_fft, win_len=win_len, hop_len=hop, fs=fs,
                     modulated=mod, derivative=True, dtype=dtype)
    assert tuple(Sx.shape) == g['Sx'].shape
    assert relerr(_np(Sx), g['Sx']) < TOL[dtype]
    assert relerr(_np(dSx), g['dSx']) < TOL[dtype]
    Tx, Sx2, freqs, Sfs, dSx2 = S.ssq_stft(g['x'], n_fft=n_fft, win_len=win_len,
                                           hop_len=hop, fs=fs, modulated=mod,
                                           dtype=dtype, get_dWx=True)
    assert np.array_equal(_np(Sfs), g['Sfs'])
    assert np.array_equal(np.asarray(freqs), g['ssq_freqs'])
    assert np.array_equal(_np(Sx2), _np(Sx))
    gamma = 10 * (O.EPS64 if dtype == 'float64' else O.EPS32)
    S3 = _np(Sx2); dS3 = _np(dSx2); T3 = _np(Tx); G3 = g['Tx']
    if S3.ndim == 2:
        S3, dS3, T3, G3 = S3[None], dS3[None], T3[None], G3[None]
    sfs = g['Sfs']
    for S_, dS_, T, Tg in zip(S3, dS3, T3, G3):
        Tref = O.ssqueeze_fused(S_, dS_, sfs, sfs[1] - sfs[0], False, False, gamma,
                                Sfs=sfs)
        assert relerr(T, Tref) < (2e-6 if dtype == 'float32' else 1e-14)
        assert relerr(T.sum(0), Tg.sum(0)) < (5e-5 if dtype == 'float32' else 1e-10)
    # stand-alone operator on the reference's own Sx, dSx is bit exact
    Top = S.ssqueeze_fast(g['Sx'], g['dSx'], sfs, sfs[1] - sfs[0], False, False,
                          gamma, Sfs=sfs)
    assert np.array_equal(_np(Top), g['Tx'])


# ---------------------------------------------------------------------------
def test_host_buffer_c_abi(S):
    """The C entry points with HOST pointers (H2D + compute + D2H inside)."""
    import torch
    from ssqueezepy_b200 import _lib
    g = load_golden('cwt_morlet_f32')
    wav, _ = _wavs('cwt_morlet_f32', S)
    x = np.ascontiguousarray(g['x'])
    Tx_d, Wx_d, *_ = S.ssq_cwt(x, wav, scales=g['scales_in'])
    N = len(x)
    n_up, n1, _ = S.utils.p2up(N)
    plan = S.CwtPlan.get(wav, np.asarray(g['scales_in'], dtype='float32'), N, n_up,
                         n1, 'reflect', 1.)
    Wx = np.empty((1, plan.na, N), dtype=np.complex64)
    Tx = np.empty_like(Wx)
    lib = _lib.load()
    _lib.check(lib.ssqb_ssq_cwt_exec_host(plan.handle, x.ctypes.data, 1,
                                          Wx.ctypes.data, Tx.ctypes.data, None,
                                          torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(Wx[0], _np(Wx_d))
    assert relerr(Tx[0], _np(Tx_d)) < 2e-6
    assert _lib.launch_count() > 0


# ---------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties
# ---------------------------------------------------------------------------
def _torch_reference_rows(x, owav, scales, rows, fs=1.):
    """float64 cuFFT evaluation of a few CWT rows (checker only)."""
    import torch
    xp, n_up, n1, _ = O.padsignal(x.astype(np.float64))
    xh = torch.fft.fft(torch.as_tensor(xp, device='cuda'))
    psih = owav.psih(np.asarray(scales)[rows], n_up).astype(np.float64)
    P = torch.as_tensor(psih, device='cuda') * xh
    W = torch.fft.ifft(P, dim=-1)
    xi = torch.as_tensor(O.xi_grid(n_up, owav.dtype).astype(np.float64), device='cuda')
    dW = torch.fft.ifft(P * (1j * xi * fs), dim=-1)
    N = x.shape[-1]
    return W[:, n1:n1 + N].cpu().numpy(), dW[:, n1:n1 + N].cpu().numpy()


@pytest.mark.parametrize('cfg', ['C1', 'C2'])
def test_full_size_ssq_cwt_properties(S, cfg):
    import torch
    N = 10_000 if cfg == 'C1' else 160_000
    g = load_golden('host_params')
    wav, owav = _wavs('cwt_morlet_f32', S)
    scales = g[f'{cfg}_scales']
    x = O.chirp(N, 0)
    Tx, Wx, freqs, sc, dWx = S.ssq_cwt(x, wav, scales=scales, get_dWx=True)
    assert np.array_equal(np.asarray(freqs), g[f'{cfg}_ssq_freqs'][::-1])
    rows = [0, 7, 60, 150, 222, 299]
    Wr, dWr = _torch_reference_rows(x, owav, scales, rows)
    assert relerr(_np(Wx[rows]), Wr) < 1e-5
    assert relerr(_np(dWx[rows]), dWr) < 1e-5
    # column-sum identity: sum_k Tx[k, j] == sum_i [|Wx|>gamma] Wx[i, j] * const
    gamma = 10 * O.EPS32
    const = np.float32(np.log(2) / int(g[f'{cfg}_nv'][0]))
    act = Wx.abs() > gamma
    lhs = Tx.sum(0)
    rhs = (Wx * act).sum(0) * float(const)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 2e-5
    # deterministic Wx; linear in x
    Tx2, Wx2, *_ = S.ssq_cwt(2 * x, wav, scales=scales)
    assert relerr(_np(Wx2), 2 * _np(Wx)) < 1e-6
    # batched == per-sample (reference tests/fft_test.py:559-631)
    xb = np.stack([O.chirp(N, b) for b in range(3)])
    Tb, Wb, *_ = S.ssq_cwt(xb, wav, scales=scales)
    assert torch.equal(Wb[0], Wx)
    T1, W1, *_ = S.ssq_cwt(xb[2], wav, scales=scales)
    assert torch.equal(Wb[2], W1)
    assert float((Tb[2] - T1).abs().max()) < 1e-5 * float(T1.abs().max())


def test_full_size_ssq_stft_C3(S):
    N = 160_000
    x = O.chirp(N, 0)
    Tx, Sx, freqs, Sfs, dSx = S.ssq_stft(x, n_fft=512, hop_len=128, get_dWx=True)
    assert tuple(Sx.shape) == (257, 1250)
    Sr, dSr = O.stft(x, None, 512, None, 128, 1.)
    assert relerr(_np(Sx), Sr) < 1e-5 and relerr(_np(dSx), dSr) < 1e-5
    sfs = _np(Sfs)
    Tref = O.ssqueeze_fused(_np(Sx), _np(dSx), sfs, sfs[1] - sfs[0], False, False,
                            10 * O.EPS32, Sfs=sfs)
    assert relerr(_np(Tx), Tref) < 2e-6


# ---------------------------------------------------------------------------
# fast path (n_up >= 2^13: band tables, direct single-pass rows, two-pass rows)
# ---------------------------------------------------------------------------
def _pair(name, dtype, S):
    opts = {'dtype': dtype}
    okw = {}
    if name == 'gmw':
        opts.update(beta=12, gamma=3); okw = dict(beta=12, gamma=3)
    return S.Wavelet((name, opts)), O.OracleWavelet(name, dtype, **okw)


@pytest.mark.parametrize('N,dtype,name,B', [
    (6_000, 'float32', 'morlet', 1),       # n_up = 2^13, I2 = 16
    (10_000, 'float32', 'gmw', 2),         # C1 size, I2 = 32, batched
    (50_000, 'float32', 'morlet', 1),      # n_up = 2^17
    (50_000, 'float64', 'gmw', 2),         # float64 fast path
    (160_000, 'float64', 'morlet', 1),     # C2 size in float64
])
def test_fast_path_ssq_cwt(S, N, dtype, name, B):
    wav, owav = _pair(name, dtype, S)
    na = 96
    scales = O.bench_scales(owav, N, na)
    x = np.stack([O.chirp(N, b, dtype) for b in range(B)])
    xin = x if B > 1 else x[0]
    Tx, Wx, freqs, sc, dWx = S.ssq_cwt(xin, wav, scales=scales, get_dWx=True)
    Tx, Wx, dWx = [_np(t).reshape(B, na, N) for t in (Tx, Wx, dWx)]
    rows = [0, 1, 17, 40, 63, 80, 95]
    tol = TOL[dtype]
    st, nv = O.infer_scaletype(_np(sc))
    ofreqs = O.ssq_freqs_cwt(_np(sc), N, owav, 'log', 'peak', 1., True)
    assert np.array_equal(np.asarray(freqs), ofreqs[::-1])
    const = O.cwt_const(_np(sc), st, nv)
    gamma = 10 * (O.EPS64 if dtype == 'float64' else O.EPS32)
    sq = O.ssqueeze_fused_c if O.c_reassign_available() else O.ssqueeze_fused
    for b in range(B):
        Wr, dWr = _torch_reference_rows(x[b], owav, scales, rows)
        assert relerr(Wx[b][rows], Wr) < tol, (b, relerr(Wx[b][rows], Wr))
        assert relerr(dWx[b][rows], dWr) < tol
        Tref = sq(Wx[b], dWx[b], ofreqs, const, True, True, gamma)
        assert relerr(Tx[b], Tref) < (2e-6 if dtype == 'float32' else 1e-14)
        assert np.array_equal(Tx[b] != 0, Tref != 0)


@pytest.mark.parametrize('N,dtype', [(20_000, 'float32'), (50_000, 'float64')])
def test_fast_path_cwt_variants(S, N, dtype):
    """cwt-only epilogues of the fast path: no derivative, derivative, L2 norm,
    rpadded -- against the float64 cuFFT rows and against each other."""
    wav, owav = _pair('morlet', dtype, S)
    na = 80
    scales = O.bench_scales(owav, N, na)
    x = O.chirp(N, 3, dtype)
    rows = [0, 5, 33, 60, 79]
    Wr, dWr = _torch_reference_rows(x, owav, scales, rows, fs=2.)
    tol = TOL[dtype]
    W0, sc = S.cwt(x, wav, scales=scales, fs=2.)
    W1, _, dW1 = S.cwt(x, wav, scales=scales, fs=2., derivative=True)
    assert relerr(_np(W0)[rows], Wr) < tol and relerr(_np(dW1)[rows], dWr) < tol
    assert np.array_equal(_np(W0), _np(W1))
    W2, _ = S.cwt(x, wav, scales=scales, fs=2., l1_norm=False)
    ref2 = _np(W0) * np.sqrt(_np(sc).astype(dtype))[:, None]
    assert relerr(_np(W2), ref2) < 2e-7 if dtype == 'float32' else 1e-15
    Wp, _ = S.cwt(x, wav, scales=scales, fs=2., rpadded=True)
    n_up, n1, _ = S.utils.p2up(N)
    assert tuple(Wp.shape) == (na, n_up)
    # rpadded rows all take the whole-signal route; unpadded short-wavelet rows the
    # overlap-save block route: same values up to the time-aliased wavelet tails
    assert relerr(_np(Wp)[:, n1:n1 + N], _np(W0)) < (1e-6 if dtype == 'float32' else 1e-12)
    assert relerr(_np(Wp)[rows][:, n1:n1 + N], Wr) < tol


@pytest.mark.parametrize('N,dtype', [
    (1 << 20, 'float32'), (1 << 20, 'float64'),            # pass-1 length 4096
    ((1 << 18) + 77, 'float32'), (1 << 18, 'float64'),     # 1024
    (1 << 19, 'float32'), ((1 << 19) - 301, 'float64'),    # 2048
])
def test_very_long_signal_2pow20(S, N, dtype):
    """BASELINE configs[4] length (N = 2^20 -> n_up = 2^21) and the two sizes below it:
    pass-1 transforms of 1024 .. 4096 points (one array per CTA where two do not fit
    shared memory) + the row kernel: rows vs float64 cuFFT and the flip-invariant
    column-sum identity."""
    import torch
    na = 24
    wav, owav = _pair('gmw', dtype, S)
    scales = O.bench_scales(owav, N, 512)[::22][:na]
    scales = 2 ** np.linspace(np.log2(scales[0]), np.log2(scales[-1]), na)   # log grid
    x = O.chirp(N, 1, dtype)
    Tx, Wx, freqs, sc, dWx = S.ssq_cwt(x, wav, scales=scales, get_dWx=True)
    rows = [0, 3, 11, 23]
    Wr, dWr = _torch_reference_rows(x, owav, scales, rows)
    assert relerr(_np(Wx[rows]), Wr) < TOL[dtype]
    assert relerr(_np(dWx[rows]), dWr) < TOL[dtype]
    gamma = 10 * (O.EPS64 if dtype == 'float64' else O.EPS32)
    st, nv = O.infer_scaletype(_np(sc))
    const = O.cwt_const(_np(sc), st, nv)
    act = Wx.abs() > gamma
    lhs, rhs = Tx.sum(0), (Wx * act).sum(0) * float(np.dtype(dtype).type(const))
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < (2e-5 if dtype == 'float32' else 1e-12)
    del Tx, Wx, dWx
    torch.cuda.empty_cache()
