# -*- coding: utf-8 -*-
"""CPU tests of the PRODUCT's host parameter layer against values produced by the
real reference (tests/golden/host_params.npz etc.): these float64 numbers decide
the reassignment bin edges, so equality is exact."""
import numpy as np
import pytest

from conftest import load_golden
import ssqueezepy_b200 as S
from ssqueezepy_b200._ssq_cwt import ssq_cwt_host_params
from ssqueezepy_b200.algos import reassign_params, make_reassign_desc
from ssqueezepy_b200._cwt import _band_limits
from ssqueezepy_b200._stft import _StftCall, get_window


CFGS = {'C1': ('morlet', {}, 10_000, 300, 'float32'),
        'C2': ('morlet', {}, 160_000, 300, 'float32'),
        'C4': ('gmw', dict(beta=12, gamma=3), 160_000, 300, 'float32'),
        'C5': ('gmw', dict(beta=12, gamma=3, dtype='float64'), 1 << 20, 512, 'float64')}


@pytest.mark.parametrize('tag', list(CFGS))
def test_baseline_config_parameters_equal_reference(tag):
    g = load_golden('host_params')
    name, opts, N, na, dtype = CFGS[tag]
    wav = S.Wavelet((name, dict(opts)))
    assert wav.dtype == dtype
    mn, mx = S.cwt_scalebounds(wav, N, preset='maximal')
    assert np.array_equal(np.array([mn, mx]), g[f'{tag}_bounds'][:2])
    nv = int(np.ceil(na / np.log2(mx / mn)))
    p0 = int(np.floor(nv * np.log2(mn)))
    scales = 2 ** (np.arange(p0, p0 + na) / nv)
    assert np.array_equal(scales, g[f'{tag}_scales'])
    hp = ssq_cwt_host_params(N, wav, scales, 'log', 'peak', True, 1.)
    assert hp['scales'].dtype == np.dtype(dtype)
    assert np.array_equal(hp['ssq_freqs'], g[f'{tag}_ssq_freqs'])
    p = reassign_params(hp['ssq_freqs'], True)
    assert np.array_equal(np.array([p['a0'], p['d0']]), g[f'{tag}_vlmin_dvl'])
    assert hp['const'] == np.log(2) / int(g[f'{tag}_nv'][0]) and hp['logscale']
    d = make_reassign_desc(hp['ssq_freqs'], hp['const'], na, True, True, 1e-6, dtype)
    assert d.kind == 0 and d.const_wide == 0
    assert d.cst_host[0] == np.dtype(dtype).type(hp['const'])


@pytest.mark.parametrize('tag', ['cwt_morlet_f32', 'cwt_gmw_f64', 'cwt_lin_f32',
                                 'cwt_piecewise_f32', 'cwt_gmw_f32_batch'])
def test_ssq_freqs_of_golden_cases(tag):
    g = load_golden(tag)
    if 'morlet' in tag or 'lin' in tag:
        wav = S.Wavelet('morlet')
    elif 'f64' in tag:
        wav = S.Wavelet(('gmw', {'beta': 12, 'gamma': 3, 'dtype': 'float64'}))
    elif 'piecewise' in tag:
        wav = S.Wavelet('gmw')
    else:
        wav = S.Wavelet(('gmw', {'beta': 12, 'gamma': 3}))
    N = g['x'].shape[-1]
    st = S.infer_scaletype(g['scales_in'])[0]
    hp = ssq_cwt_host_params(N, wav, g['scales_in'], st, 'peak', True, 1 / float(g['fs']))
    assert np.array_equal(hp['scales'].squeeze(), g['scales_out'])
    assert np.array_equal(hp['ssq_freqs'][::-1], g['ssq_freqs'])
    if 'piecewise' in tag:
        d = make_reassign_desc(hp['ssq_freqs'], hp['const'], len(g['scales_out']), True,
                               True, 1e-6, 'float32')
        assert d.kind == 1 and d.const_wide == 1      # float64 const on float32 data


def test_default_scales_and_bounds():
    g = load_golden('host_params')
    for N in (2000, 160_000):
        wav = S.Wavelet()
        assert wav.name == 'GMW L1' and wav.dtype == 'float32'
        for preset in ('maximal', 'minimal'):
            assert np.array_equal(np.array(S.cwt_scalebounds(wav, N, preset=preset)),
                                  g[f'default_bounds_{preset}_{N}'])
        sc = S.process_scales('log-piecewise', N, wav, nv=32).squeeze()
        assert np.array_equal(sc, g[f'default_scales_{N}'])
        st, nv = S.infer_scaletype(sc.astype('float32'))
        assert st == 'log-piecewise' and nv.shape == (len(sc), 1)


def test_band_limits_cover_the_wavelet():
    """Every frequency bin outside the band must hold a negligible wavelet value."""
    for spec, dtype in [('morlet', 'float32'), (('morlet', {'mu': 5}), 'float32'),
                        (('gmw', {'beta': 12, 'gamma': 3}), 'float32'),
                        (('gmw', {'beta': 12, 'gamma': 3, 'dtype': 'float64'}), 'float64')]:
        wav = S.Wavelet(spec)
        n_up = 4096
        scales = np.array([0.6, 2., 9., 40., 300., 1500.])
        lo, ln = _band_limits(wav, scales.astype(dtype), n_up)
        psih = np.abs(np.asarray(wav(scale=scales.astype(dtype), N=n_up), dtype=np.float64))
        peak = psih.max()                      # ~ the wavelet's global maximum
        for a in range(len(scales)):
            idx = (lo[a] + np.arange(ln[a])) % n_up
            mask = np.ones(n_up, bool); mask[idx] = False
            tol = (1e-9 if dtype == 'float32' else 1e-21) * peak
            assert ln[a] <= n_up and (psih[a][mask] <= tol).all(), (spec, a)


def test_stft_host_parameters():
    for tag in ['stft_f32', 'stft_f64_odd', 'stft_f32_batch', 'stft_f32_nomod']:
        g = load_golden(tag)
        dtype = str(g['x'].dtype)
        N = g['x'].shape[-1]
        call = _StftCall(N, None, int(g['n_fft']), int(g['win_len']), int(g['hop']),
                         float(g['fs']), 'reflect', bool(g['modulated']), dtype)
        tol = 1e-6 if dtype == 'float32' else 1e-12
        assert np.allclose(call.window, g['window'], rtol=tol, atol=tol * 1e-3)
        assert np.allclose(call.diff_window, g['diff_window'], rtol=tol, atol=tol * 1e-3)
        assert np.array_equal(call.Sfs, g['Sfs'])
        assert (call.n_rows, call.n_hops) == g['Sx'].shape[-2:]
    with pytest.raises(ValueError):
        get_window(None, 65, 64)


def test_wavelet_api_and_errors():
    w = S.Wavelet(('morlet', {'mu': 6}), N=128)
    assert w.xi.shape == (128,) and w.xi.dtype == np.float32
    assert w(scale=np.array([2., 4.]), N=64).shape == (2, 64)
    assert w(np.array([6.])).shape == (1,)
    assert w.Psih(np.array([2., 4.]), N=64, nohalf=False)[0, 32] == \
        w(scale=np.array([2., 4.]), N=64, nohalf=True)[0, 32] / 2
    with pytest.raises(ValueError):
        S.Wavelet('nope')
    with pytest.raises(TypeError):
        S.Wavelet(3)
    with pytest.raises(ValueError):
        S.Wavelet(('gmw', {'order': -1}))
    with pytest.raises(ValueError):
        S.Wavelet(('gmw', {'norm': 'l3'}))
    assert S.Wavelet(('gmw', {'norm': 'energy', 'dtype': 'float64'})).name == 'GMW L2'
    assert S.Wavelet('bump').device_spec() is None
    assert S.Wavelet('morlet').device_spec()[0] == 'morlet'
    assert abs(S.center_frequency(S.Wavelet('morlet'), kind='peak-ct') - 13.4) < 1e-2


def test_padsignal_and_buffer_semantics():
    from oracle import ssq_oracle as O
    x = np.arange(1., 12.)
    for mode in ('reflect', 'zero', 'symmetric', 'replicate', 'wrap'):
        assert np.array_equal(S.padsignal(x, mode), O.padsignal(x, mode)[0])
        assert np.array_equal(S.padsignal(x, mode, padlength=20), O.padsignal(x, mode, 20)[0])
    gb = load_golden('buffer')
    for k in range(5):
        seg, ov, mod = [int(v) for v in gb[f'p{k}']]
        assert np.array_equal(S.buffer(gb['x'], seg, ov, bool(mod)), gb[f'b{k}'])


# ---- GMW beyond L1 order 0 (host-evaluated, table path) ----------------------------
def test_gmw_l2_and_higher_order_match_reference():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    import ssqueezepy_b200 as S
    from ssqueezepy_b200.utils.cwt_utils import cwt_scalebounds
    g = load_golden('gmw_variants')
    variants = [dict(norm='energy', beta=12, gamma=3), dict(order=1),
                dict(order=2, beta=12, gamma=3), dict(norm='energy', order=1),
                dict(norm='energy', order=3, beta=5, gamma=2, dtype='float64'),
                dict(norm='energy', dtype='float64'),
                dict(order=2, dtype='float64', centered_scale=True)]
    for k, opts in enumerate(variants):
        wav = S.Wavelet(('gmw', dict(opts)))
        assert wav.name == str(g['name%d' % k][0])
        assert wav.device_spec() is None              # evaluated on the host, uploaded as a table
        v = np.asarray(wav.fn(g['w'].copy()))
        ref = g['v%d' % k]
        # (the reference's L2 order-0 function returns float64 even for a float32 wavelet:
        # its amplitude is a float64 expression; values are compared, not dtypes)
        assert str(v.dtype) == wav.dtype and np.all(np.isfinite(v))
        tol = 1e-13 if wav.dtype == 'float64' else 2e-5   # beta = 60 in float32: libm noise
        assert np.abs(v.astype(np.float64) - ref).max() <= tol * np.abs(ref).max()
        mine = np.array(cwt_scalebounds(wav, 4096, preset='maximal'))
        assert np.allclose(mine, g['bounds%d' % k], rtol=1e-3 if wav.dtype == 'float32' else 1e-9)
    # default beta = 60 in float32: Gamma(r) overflows in the reference (all-zero wavelet);
    # the log-form evaluation here stays finite and agrees with the float64 wavelet
    w = np.linspace(0.5, 6, 500)
    v32 = S.Wavelet(('gmw', dict(norm='energy'))).fn(w)
    v64 = S.Wavelet(('gmw', dict(norm='energy', dtype='float64'))).fn(w)
    assert np.all(np.isfinite(v32)) and np.abs(v32 - v64).max() < 1e-4 * v64.max()


def test_gmw_variant_tables_reproduce_reference_cwt():
    """The `psih` table the plan uploads for the host-evaluated GMW variants, pushed
    through a plain NumPy FFT convolution, gives the reference's `cwt` output: pins the
    host side of the table path (sampling, Nyquist halving, sqrt(scale) of `l1_norm=False`)
    without a GPU; tests/test_zz_gmw_variants_gpu.py runs the same through the kernels."""
    import scipy.fft as sfft
    import ssqueezepy_b200 as S
    from ssqueezepy_b200._cwt import _process_gmw_wavelet
    from ssqueezepy_b200.utils.common import p2up
    g = load_golden('gmw_variants')
    x, sc = g['x'], g['scales']
    N = len(x)
    n_up, n1, n2 = p2up(N)
    xh = sfft.fft(np.pad(x, [n1, n2], mode='reflect')).astype(np.complex64)
    for key, spec, l1 in (('Wx_l2', ('gmw', {'beta': 12, 'gamma': 3}), False),
                          ('Wx_k2', ('gmw', {'beta': 12, 'gamma': 3, 'order': 2}), True)):
        wav = S.Wavelet._init_if_not_isinstance(_process_gmw_wavelet(spec, l1))
        sc_t = np.asarray(sc, dtype=wav.dtype).reshape(-1, 1)
        tab = np.asarray(wav(scale=sc_t, N=n_up, nohalf=False))
        W = sfft.ifft(tab * xh, axis=-1)[:, n1:n1 + N]
        if not l1:
            W = W * np.sqrt(sc_t)
        assert np.linalg.norm(W - g[key]) / np.linalg.norm(g[key]) < 5e-6


def test_cwt_higher_order_composition(monkeypatch):
    """`cwt(order=...)` / `cwt_higher_order`: which wavelets are built, how scales are
    shared and how orders are averaged -- checked against the reference's outputs with the
    per-order transform replaced by a NumPy evaluation of the same `psih` table (the
    device transform itself is covered by the GPU tests)."""
    import scipy.fft as sfft
    import torch
    import ssqueezepy_b200 as S
    from ssqueezepy_b200 import _cwt as M
    from ssqueezepy_b200.utils.common import p2up
    g = load_golden('gmw_variants')
    x, sc = g['x'], g['scales']

    def fake_cwt(x, wavelet, scales=None, derivative=False, order=0, **kw):
        assert order == 0
        wav = S.Wavelet._init_if_not_isinstance(wavelet)
        N = len(x)
        n_up, n1, n2 = p2up(N)
        xh = sfft.fft(np.pad(x, [n1, n2], mode='reflect')).astype(np.complex64)
        sc_t = np.asarray(scales, dtype=wav.dtype).reshape(-1, 1)
        tab = np.asarray(wav(scale=sc_t, N=n_up, nohalf=False))
        W = sfft.ifft(tab * xh, axis=-1)
        xi = S.wavelets.xi_grid(n_up, 1., wav.dtype)
        dW = sfft.ifft(tab * xh * (1j * xi).astype(np.complex64), axis=-1)
        out = (torch.as_tensor(W[:, n1:n1 + N]), torch.as_tensor(sc_t.squeeze()))
        return out + (torch.as_tensor(dW[:, n1:n1 + N]),) if derivative else out

    monkeypatch.setattr(M, 'cwt', fake_cwt)
    rel = lambda a, b: np.linalg.norm(np.asarray(a) - b) / np.linalg.norm(b)
    W2, s2 = M.cwt_higher_order(x, ('gmw', {'beta': 12, 'gamma': 3}), order=2, scales=sc)
    assert rel(W2, g['Wx_order2']) < 5e-6 and rel(W2, g['Wx_k2']) < 5e-6
    W, s, dW = M.cwt_higher_order(x, ('gmw', {'beta': 12, 'gamma': 3}), order=(0, 1, 2),
                                  scales=sc, derivative=True)
    assert tuple(W.shape) == g['Wx_order012'].shape
    assert rel(W, g['Wx_order012']) < 5e-6 and rel(dW, g['dWx_order012']) < 5e-6
    Wl, _ = M.cwt_higher_order(x, ('gmw', {'beta': 12, 'gamma': 3}), order=(0, 1),
                               scales=sc, average=False)
    assert isinstance(Wl, list) and len(Wl) == 2
    with pytest.raises(ValueError):
        M.cwt_higher_order(x, 'morlet', order=1, scales=sc)
