# -*- coding: utf-8 -*-
"""Pins oracle/ssq_oracle.py against the outputs of the REAL reference stored
in tests/golden/ (made by tests/golden/make_golden.py).  CPU only."""
import os
import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle import ssq_oracle as O


def _wav(tag):
    if 'morlet' in tag or 'lin' in tag:
        return O.OracleWavelet('morlet', 'float32')
    if 'gmw_f64' in tag:
        return O.OracleWavelet('gmw', 'float64', beta=12, gamma=3)
    if 'piecewise' in tag:
        return O.OracleWavelet('gmw', 'float32')            # beta=60 default
    return O.OracleWavelet('gmw', 'float32', beta=12, gamma=3)


CWT_CASES = ['cwt_morlet_f32', 'cwt_gmw_f64', 'cwt_gmw_f32_batch',
             'cwt_lin_f32', 'cwt_piecewise_f32']


@pytest.mark.parametrize('tag', CWT_CASES)
def test_cwt_matches_reference(tag):
    g = load_golden(tag)
    wav = _wav(tag)
    fs = float(g['fs'])
    Wx, sc, dWx = O.cwt(g['x'], wav, g['scales_in'], fs=fs)
    tol = 2e-6 if wav.dtype == np.float32 else 1e-13
    if 'piecewise' in tag:
        # GMW beta=60 in float32: beta*log(w) ~ 200, so libm powf/logf/expf
        # differences (numba vs numpy) show at the 3e-6 level
        tol = 6e-6
    assert Wx.dtype == g['Wx'].dtype
    assert np.array_equal(sc, g['scales_out'])
    assert relerr(Wx, g['Wx']) < tol
    assert relerr(dWx, g['dWx']) < tol
    if 'Wx_l2' in g.files:
        Wl2, _ = O.cwt(g['x'], wav, g['scales_in'], fs=fs, derivative=False,
                       l1_norm=False)
        assert relerr(Wl2, g['Wx_l2']) < tol


@pytest.mark.parametrize('tag', CWT_CASES)
def test_ssq_freqs_and_reassign_bit_exact_given_reference_cwt(tag):
    """Host parameters must equal the reference's doubles exactly; the
    reassignment fed with the REFERENCE's (Wx, dWx) must reproduce its Tx
    bit-for-bit (SURVEY section 8c contract (2))."""
    g = load_golden(tag)
    wav = _wav(tag)
    fs = float(g['fs'])
    N = g['x'].shape[-1]
    sc = g['scales_out']
    st, nv = O.infer_scaletype(sc)
    st_in, _ = O.infer_scaletype(g['scales_in'])
    freqs = O.ssq_freqs_cwt(sc, N, wav, st_in, 'peak', 1 / fs, True)
    assert np.array_equal(freqs[::-1], g['ssq_freqs'])
    const = O.cwt_const(sc, st, nv)
    gamma = 10 * (O.EPS64 if g['Wx'].dtype == np.complex128 else O.EPS32)
    Wx, dWx = g['Wx'], g['dWx']
    if Wx.ndim == 2:
        Wx, dWx, Txg = Wx[None], dWx[None], g['Tx'][None]
    else:
        Txg = g['Tx']
    for W, dW, T in zip(Wx, dWx, Txg):
        Tx = O.ssqueeze_fused(W, dW, freqs, const, st_in.startswith('log'),
                              True, gamma)
        assert np.array_equal(Tx, T)


@pytest.mark.parametrize('tag', ['cwt_morlet_f32', 'cwt_gmw_f64'])
def test_ssq_cwt_end_to_end(tag):
    g = load_golden(tag)
    wav = _wav(tag)
    Tx, Wx, freqs, sc = O.ssq_cwt(g['x'], wav, g['scales_in'], fs=float(g['fs']))
    assert np.array_equal(freqs, g['ssq_freqs'])
    # column sums are invariant to bin flips
    assert relerr(Tx.sum(0), g['Tx'].sum(0)) < (2e-5 if wav.dtype == np.float32 else 1e-11)


STFT_CASES = ['stft_f32', 'stft_f64_odd', 'stft_f32_batch', 'stft_f32_nomod']


@pytest.mark.parametrize('tag', STFT_CASES)
def test_stft_matches_reference(tag):
    g = load_golden(tag)
    dtype = str(g['x'].dtype)
    n_fft, hop, fs = int(g['n_fft']), int(g['hop']), float(g['fs'])
    win_len, mod = int(g['win_len']), bool(g['modulated'])
    w, dw = O.get_window(None, win_len, n_fft, dtype)
    tolw = 1e-6 if dtype == 'float32' else 1e-12
    assert relerr(w, g['window']) < tolw and relerr(dw, g['diff_window']) < tolw
    Sx, dSx = O.stft(g['x'], None, n_fft, win_len, hop, fs, modulated=mod,
                     dtype=dtype)
    tol = 2e-6 if dtype == 'float32' else 1e-12
    assert Sx.shape == g['Sx'].shape
    assert relerr(Sx, g['Sx']) < tol and relerr(dSx, g['dSx']) < tol
    Tx, Sx2, freqs, Sfs = O.ssq_stft(g['x'], None, n_fft, win_len, hop, fs,
                                     modulated=mod, dtype=dtype)
    assert np.array_equal(Sfs, g['Sfs']) and np.array_equal(freqs, g['ssq_freqs'])
    # reassignment on the reference's own Sx, dSx: bit exact
    gamma = 10 * (O.EPS64 if g['Sx'].dtype == np.complex128 else O.EPS32)
    Sxg, dSxg, Txg = g['Sx'], g['dSx'], g['Tx']
    if Sxg.ndim == 2:
        Sxg, dSxg, Txg = Sxg[None], dSxg[None], Txg[None]
    for S_, dS_, T in zip(Sxg, dSxg, Txg):
        Tx_ = O.ssqueeze_fused(S_, dS_, Sfs, Sfs[1] - Sfs[0], False, False, gamma,
                               Sfs=Sfs)
        assert np.array_equal(Tx_, T)


@pytest.mark.parametrize('dtype', ['float32', 'float64'])
@pytest.mark.parametrize('flipud', [False, True])
def test_reassign_kernels_bit_exact(dtype, flipud):
    g = load_golden('reassign')
    Wx, dWx = g[f'{dtype}_Wx'], g[f'{dtype}_dWx']
    gamma = 10 * np.finfo(dtype).eps
    tag = f'{dtype}_flip{int(flipud)}'
    carr = g[f'{dtype}_const_arr']
    Sfs = g[f'{dtype}_Sfs']
    cases = [('log', g[f'{dtype}_flog'], np.log(2) / 8, True, None),
             ('pw', g[f'{dtype}_fpw'], carr, True, None),
             ('lin', g[f'{dtype}_flin'], carr, False, None),
             ('stft', Sfs, float(Sfs[1] - Sfs[0]), False, Sfs)]
    for name, freqs, const, logscale, sfs in cases:
        Tx = O.ssqueeze_fused(Wx, dWx, freqs, const, logscale, flipud, gamma,
                              Sfs=sfs)
        assert np.array_equal(Tx, g[f'Tx_{name}_{tag}']), name
    w = O.phase_cwt(Wx, dWx, gamma)
    assert np.array_equal(w, g[f'{dtype}_w_cwt'])
    ws = O.phase_stft(Wx, dWx, Sfs, gamma)
    assert np.array_equal(ws, g[f'{dtype}_w_stft'])
    for name, freqs, const, logscale in [c[:4] for c in cases[:3]]:
        Ix = O.indexed_sum_onfly(Wx, w, freqs, const, logscale, flipud)
        assert np.array_equal(Ix, g[f'Ix_{name}_{tag}']), name


def test_host_params_baseline_configs():
    g = load_golden('host_params')
    cfgs = {'C1': ('morlet', {}, 10_000, 300, 'float32'),
            'C2': ('morlet', {}, 160_000, 300, 'float32'),
            'C4': ('gmw', dict(beta=12, gamma=3), 160_000, 300, 'float32'),
            'C5': ('gmw', dict(beta=12, gamma=3), 1 << 20, 512, 'float64')}
    for tag, (name, opts, N, na, dtype) in cfgs.items():
        wav = O.OracleWavelet(name, dtype, **opts)
        mn, mx = O.cwt_scalebounds_maximal(wav, N)
        assert np.array_equal(np.array([mn, mx]), g[f'{tag}_bounds'][:2]), tag
        scales = O.bench_scales(wav, N, na)
        assert np.array_equal(scales, g[f'{tag}_scales']), tag
        sc = scales.astype(dtype)
        st, nv = O.infer_scaletype(sc)
        assert st == str(g[f'{tag}_scaletype'][0]) and nv == int(g[f'{tag}_nv'][0])
        freqs = O.ssq_freqs_cwt(sc, N, wav, st, 'peak', 1., True)
        assert np.array_equal(freqs, g[f'{tag}_ssq_freqs']), tag
        p = O.reassign_params(freqs, True)
        assert np.array_equal(np.array([p['vlmin'], p['dvl']]), g[f'{tag}_vlmin_dvl'])


def test_buffer_exact():
    g = load_golden('buffer')
    x = g['x']
    for k in range(5):
        seg, ov, mod = [int(v) for v in g[f'p{k}']]
        assert np.array_equal(O.buffer(x, seg, ov, bool(mod)), g[f'b{k}'])


def test_pad_modes():
    x = np.arange(1, 8, dtype=np.float64)
    for mode in ('reflect', 'zero', 'symmetric', 'replicate', 'wrap'):
        xp, n_up, n1, n2 = O.padsignal(x, mode)
        assert len(xp) == n_up == 16 and n1 + n2 + 7 == 16
        assert np.array_equal(xp[n1:n1 + 7], x)
    assert np.array_equal(O.padsignal(np.array([1., 2, 3, 4]), 'symmetric',
                                      padlength=11)[0],
                          np.array([4, 3, 2, 1, 1, 2, 3, 4, 4, 3, 2.]))


@pytest.mark.parametrize('dtype', ['float32', 'float64'])
def test_c_reassign_oracle_equals_numpy_oracle_and_reference(dtype):
    """oracle/reassign_oracle.c (compiled by __graft_entry__.build / make -C oracle)
    is bit-identical to the reference on its own random test inputs."""
    if not O.c_reassign_available():
        pytest.skip("oracle/_build/libreassign_oracle.so not built")
    g = load_golden('reassign')
    Wx, dWx = g[f'{dtype}_Wx'], g[f'{dtype}_dWx']
    gamma = 10 * np.finfo(dtype).eps
    carr, Sfs = g[f'{dtype}_const_arr'], g[f'{dtype}_Sfs']
    for flipud in (False, True):
        tag = f'{dtype}_flip{int(flipud)}'
        for name, freqs, const, logscale, sfs in [
                ('log', g[f'{dtype}_flog'], np.log(2) / 8, True, None),
                ('pw', g[f'{dtype}_fpw'], carr, True, None),
                ('lin', g[f'{dtype}_flin'], carr, False, None),
                ('stft', Sfs, float(Sfs[1] - Sfs[0]), False, Sfs)]:
            Tx = O.ssqueeze_fused_c(Wx, dWx, freqs, const, logscale, flipud, gamma,
                                    Sfs=sfs)
            assert np.array_equal(Tx, g[f'Tx_{name}_{tag}']), (name, flipud)
    gp = load_golden('cwt_piecewise_f32')     # float64 `const` on float32 data
    wav = O.OracleWavelet('gmw', 'float32')
    sc = gp['scales_out']
    st, nv = O.infer_scaletype(sc)
    freqs = O.ssq_freqs_cwt(sc, gp['x'].shape[-1], wav, 'log-piecewise', 'peak', 1., True)
    Tx = O.ssqueeze_fused_c(gp['Wx'], gp['dWx'], freqs, O.cwt_const(sc, st, nv), True,
                            True, 10 * O.EPS32)
    assert np.array_equal(Tx, gp['Tx'])


def test_log2f_restatement_equals_libm_on_every_float32():
    """The two-step reassignment takes np.log2 of a float32 `w` (algos.py:175-216); numba
    calls the host libm's log2f for it.  oracle/log2f_glibc.c restates glibc's algorithm
    (the CUDA operator evaluates the same operations): compared here with the libm of this
    process on all 2 139 095 039 positive finite float32 inputs, both with and without
    FMA contraction of the polynomial (glibc ships both builds)."""
    import ctypes
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        'oracle', '_build', 'liblog2f_glibc.so')
    if not os.path.isfile(path):
        pytest.skip("oracle/_build/liblog2f_glibc.so not built (make -C oracle)")
    lib = ctypes.CDLL(path)
    lib.log2f_check.restype = ctypes.c_longlong
    for variant in (1, 0):
        bad = ctypes.c_uint32(0)
        assert lib.log2f_check(variant, ctypes.byref(bad)) == 0, hex(bad.value)
