# -*- coding: utf-8 -*-
"""`phase_transform` / `phase_ssqueeze`: the phase transform followed by `ssqueeze` on an
arbitrary CWT- / STFT-like plane -- same arguments and return tuples as the reference's
`ssqueezepy/experimental.py:145-259`, running on the CUDA operators of `algos.py`.

`dWx` (the time derivative of the transform) must be given: the reference's
`trigdiff` fallback (row-wise frequency-domain differentiation of `Wx`) and the
`'phase'` / `'numeric'` difference schemes have no device implementation and raise.
"""
from . import backend as Bk
from ._ssq_cwt import phase_cwt
from ._ssq_stft import phase_stft
from .ssqueezing import ssqueeze
from .utils.common import EPS32, EPS64, p2up

import numpy as np

__all__ = ['phase_ssqueeze', 'phase_transform']


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
