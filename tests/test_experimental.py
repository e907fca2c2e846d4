# -*- coding: utf-8 -*-
"""`experimental.phase_ssqueeze` / `phase_transform` against
outputs of the real reference on its own `Wx, dWx` / `Sx, dSx` (tests/golden/experimental.npz).
The operators underneath are the deterministic column-owner kernels, so the comparison
is bit for bit."""
import numpy as np
import pytest

from conftest import load_golden


def test_argument_errors_need_no_device():
    import ssqueezepy_b200 as S
    Wx = np.zeros((4, 16), dtype=np.complex64)
    with pytest.raises(NotImplementedError):
        S.phase_ssqueeze(Wx, None, scales=np.arange(1., 5.))
    with pytest.raises(ValueError):
        S.phase_transform(Wx, Wx, rpadded=True)
    with pytest.raises(ValueError):
        S.phase_transform(Wx, Wx, difftype='numeric')
    with pytest.raises(NotImplementedError):
        S.phase_transform(np.zeros((2, 4, 16), dtype=np.complex64),
                          np.zeros((2, 4, 16), dtype=np.complex64), get_w=True)


def test_scale_frequency_conversions_match_reference():
    from ssqueezepy_b200.experimental import freq_to_scale, scale_to_freq
    ref = load_golden('experimental')
    fr, sc = ref['conv_freqs'], ref['conv_scales_in']
    assert np.array_equal(freq_to_scale(fr, 'morlet', 2048, fs=500), ref['f2s_morlet'])
    assert np.array_equal(scale_to_freq(sc, ('gmw', {'beta': 12, 'gamma': 3}), 1500, fs=2.),
                          ref['s2f_gmw'])
    assert np.array_equal(scale_to_freq(sc, 'morlet', 1500, padtype=None),
                          ref['s2f_morlet_nopad'])


@pytest.fixture(scope='module')
def S():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import ssqueezepy_b200 as S_
    return S_


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)


@pytest.mark.gpu
@pytest.mark.parametrize('tag,kw', [('fused', {}), ('twostep', dict(get_w=True)),
                                    ('flip', dict(flipud=True))])
def test_gpu_phase_ssqueeze_cwt(S, tag, kw):
    ref = load_golden('experimental')
    g = load_golden('cwt_morlet_f32')
    Tx, Wx, freqs, scales, Sfs, w, dWx = S.phase_ssqueeze(
        g['Wx'], g['dWx'], scales=g['scales_in'], wavelet='morlet', **kw)
    assert np.array_equal(_np(Tx), ref['cwt_Tx_' + tag])
    assert np.array_equal(np.asarray(freqs), ref['cwt_freqs_' + tag])
    assert Sfs is None
    if tag == 'twostep':
        assert np.array_equal(_np(w), ref['cwt_w']) and dWx is None
    else:
        assert w is None and dWx is not None


@pytest.mark.gpu
@pytest.mark.parametrize('tag,kw', [('fused', {}), ('twostep', dict(get_w=True))])
def test_gpu_phase_ssqueeze_stft(S, tag, kw):
    ref = load_golden('experimental')
    g = load_golden('stft_f32')
    Tx, Sx, freqs, _, Sfs, w, dSx = S.phase_ssqueeze(
        g['Sx'], g['dSx'], ssq_freqs=g['Sfs'], transform='stft', **kw)
    assert np.array_equal(_np(Tx), ref['stft_Tx_' + tag])
    assert np.array_equal(_np(freqs), ref['stft_freqs_' + tag])
    assert np.array_equal(_np(Sfs), ref['stft_Sfs'])


@pytest.mark.gpu
def test_gpu_phase_ssqueeze_rpadded(S):
    """Padded planes in, signal part out: same Tx as on the unpadded planes."""
    g = load_golden('cwt_morlet_f32')
    N = g['Wx'].shape[1]
    n_up, n1, n2 = S.utils.p2up(N)
    pad = lambda a: np.pad(a, [(0, 0), (n1, n2)])
    T0 = S.phase_ssqueeze(g['Wx'], g['dWx'], scales=g['scales_in'], wavelet='morlet')[0]
    T1, W1, *_ = S.phase_ssqueeze(pad(g['Wx']), pad(g['dWx']), scales=g['scales_in'],
                                  wavelet='morlet', rpadded=True, N=N)
    assert tuple(W1.shape) == g['Wx'].shape
    assert np.array_equal(_np(T0), _np(T1))
