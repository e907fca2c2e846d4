# -*- coding: utf-8 -*-
"""`phase_transform` / `phase_ssqueeze`: the phase transform followed by `ssqueeze` on an
arbitrary CWT- / STFT-like plane -- same arguments and return tuples as the reference's
`ssqueezepy/experimental.py:145-259`, running on the CUDA operators of `algos.py`.

`dWx` (the time derivative of the transform) must be given: the reference's
`trigdiff` fallback (row-wise frequency-domain differentiation of `Wx`) and the
`'phase'` / `'numeric'` difference schemes have no device implementation and raise.
"""
import warnings

from . import backend as Bk
from ._ssq_cwt import phase_cwt
from ._ssq_stft import phase_stft
from .ssqueezing import ssqueeze
from .utils.common import EPS32, EPS64, p2up
from .utils.cwt_utils import cwt_scalebounds
from .wavelets import Wavelet, center_frequency

import numpy as np

__all__ = ['phase_ssqueeze', 'phase_transform', 'freq_to_scale', 'scale_to_freq']


def _stft_freqs(Sx, fs):
    """Bin centre frequencies of an STFT plane (reference `_ssq_stft.py:249-257`)."""
    return np.linspace(0, .5 * fs, Sx.shape[-2], dtype=Bk.dtype_of_complex(Sx))


def phase_transform(Wx, dWx=None, difftype='trig', difforder=4, gamma=None,
                    fs=1., Sfs=None, rpadded=False, padtype='reflect', N=None,
                    n1=None, get_w=False, transform='cwt'):
    """Returns `(w, Wx, dWx, Sfs, gamma)`; `w` is None unless `get_w`."""
    if transform not in ('cwt', 'stft'):
        raise ValueError("`transform` must be one of: cwt, stft (got %s)" % transform)
    if dWx is None:
        raise NotImplementedError(
            "`phase_transform` without `dWx` is not supported on the device "
            "(pass `dWx` from `cwt(..., derivative=True)` / `stft(..., derivative=True)`)")
    if rpadded and N is None:
        raise ValueError("`rpadded=True` requires `N`")
    if Wx.ndim > 2 and get_w:
        raise NotImplementedError("`get_w=True` unsupported with batched input.")
    if difftype not in (None, 'trig'):
        raise ValueError("GPU computation only supports `difftype = 'trig'`")
    if gamma is None:
        gamma = 10 * (EPS64 if Bk.dtype_of_complex(Wx) == 'float64' else EPS32)

    if transform == 'cwt':
        if rpadded:                       # planes came in padded: keep the signal part
            if n1 is None:
                _, n1, _ = p2up(N)
            Wx, dWx = Wx[..., n1:n1 + N], dWx[..., n1:n1 + N]
        w = phase_cwt(Wx, dWx, 'trig', gamma) if get_w else None
        return w, Wx, dWx, None, gamma
    if Sfs is None:
        Sfs = _stft_freqs(Wx, fs)
    w = phase_stft(Wx, dWx, Sfs, gamma) if get_w else None
    return w, Wx, dWx, Sfs, gamma


def phase_ssqueeze(Wx, dWx=None, ssq_freqs=None, scales=None, Sfs=None, fs=1.,
                   t=None, squeezing='sum', maprange=None, wavelet=None,
                   gamma=None, was_padded=True, flipud=False,
                   rpadded=False, padtype=None, N=None, n1=None,
                   difftype=None, difforder=None,
                   get_w=False, get_dWx=False, transform='cwt'):
    """Phase transform, then synchrosqueezing, of a given transform and its derivative.
    Returns `(Tx, Wx, ssq_freqs, scales, Sfs, w, dWx)` like the reference."""
    w, Wx, dWx, Sfs, gamma = phase_transform(
        Wx, dWx, difftype, difforder=difforder, gamma=gamma, rpadded=rpadded,
        padtype=padtype, N=N, n1=n1, get_w=get_w, fs=fs, Sfs=Sfs, transform=transform)
    dWx_out = dWx if (w is None or get_dWx) else None
    if maprange is None:
        maprange = 'peak' if transform == 'cwt' else 'maximal'
    Tx, ssq_freqs = ssqueeze(Wx, w, ssq_freqs, scales, Sfs, fs=fs, t=t,
                             squeezing=squeezing, maprange=maprange,
                             wavelet=wavelet, gamma=gamma, was_padded=was_padded,
                             flipud=flipud, dWx=None if w is not None else dWx,
                             transform=transform)
    return Tx, Wx, ssq_freqs, scales, Sfs, w, dWx_out


# ---- host-side conversions between scales and frequencies --------------------------
def freq_to_scale(freqs, wavelet, N, fs=1, n_search_scales=None, kind='peak', base=2):
    """Scales (exponentially spaced, len(freqs) of them) whose centre frequencies span
    `freqs[0]`..`freqs[-1]` Hz: the scale range is located on a search grid of
    `n_search_scales` scales through `center_frequency` (reference
    `experimental.py:15-82`).  `freqs` ascending, within [0, fs/2]."""
    f = np.asarray(freqs, dtype=np.float64) / fs
    if not np.all(f >= 0):
        raise AssertionError("frequencies must be positive")
    if f.max() > 0.5:
        raise AssertionError("max frequency must be 0.5")
    if f.max() != f[-1] or f.min() != f[0]:
        raise AssertionError("min / max frequency must be the first / last sample")
    M = len(f)
    n_search = 10 * M if n_search_scales is None else n_search_scales
    lg = lambda v: np.log(v) / np.log(base)
    smin, smax = cwt_scalebounds(wavelet, N, preset='maximal', use_padded_N=False)
    grid = np.logspace(lg(smin), lg(smax), n_search, base=base)
    fc = np.array([min(max(center_frequency(wavelet, s, N, kind=kind), 0), np.pi)
                   for s in grid]) / (2 * np.pi)
    s_hi = grid[np.argmin(np.abs(fc - f[0]))]
    s_lo = grid[np.argmin(np.abs(fc - f[-1]))]
    return np.logspace(lg(s_hi), lg(s_lo), M, base=base)


def scale_to_freq(scales, wavelet, N, fs=1, padtype='reflect'):
    """Peak frequency (Hz) of the sampled wavelet at each scale (reference
    `experimental.py:85-141`); the sampling length is the padded one unless
    `padtype is None`."""
    if isinstance(scales, float):
        scales = np.array([scales])
    wavelet = Wavelet._init_if_not_isinstance(wavelet)
    Npad = p2up(N)[0] if padtype is not None else N
    psis = np.asarray(wavelet(scale=scales, N=Npad))
    psis = psis.reshape(-1, Npad)
    idxs = np.argmax(psis, axis=-1)
    bad = (idxs > Npad // 2) | (idxs == 0)
    if bad.any():
        warnings.warn("found potentially ill-behaved wavelets (peak indices at "
                      "negative freqs or at dc); will round idxs to 1 or N/2")
        low = np.arange(len(idxs)) > len(idxs) // 2       # later rows = larger scales
        idxs = np.where(bad, np.where(low, 1, Npad // 2), idxs)
    return idxs / Npad * fs
