# -*- coding: utf-8 -*-
"""Synchrosqueezed STFT on B200 -- same signature / returns as the reference's
`ssqueezepy/_ssq_stft.py:13-136` (`ssq_stft`) and `:201-246` (`phase_stft`).
Default path: one fused kernel (framing + windows + FFT + phase transform +
linear-bin reassignment)."""
import ctypes as C
import numpy as np
import torch

from . import _lib, backend as Bk
from ._stft import stft, _StftCall, _get_call, get_window, _check_NOLA
from .algos import phase_stft_gpu, make_reassign_desc
from .ssqueezing import ssqueeze, _check_ssqueezing_args
from .utils.common import EPS32, EPS64, WARN
from .utils.cwt_utils import infer_scaletype, _process_fs_and_t

__all__ = ['ssq_stft', 'issq_stft', 'phase_stft']


def ssq_stft(x, window=None, n_fft=None, win_len=None, hop_len=1, fs=None, t=None,
             modulated=True, ssq_freqs=None, padtype='reflect', squeezing='sum',
             gamma=None, preserve_transform=None, dtype=None, astensor=True,
             flipud=False, get_w=False, get_dWx=False):
    """Returns `(Tx, Sx, ssq_freqs, Sfs[, w][, dSx])` like the reference."""
    if x.ndim == 2 and get_w:
        raise NotImplementedError("`get_w=True` unsupported with batched input.")
    N = x.shape[-1]
    _, fs, _ = _process_fs_and_t(fs, t, N)
    _check_ssqueezing_args(squeezing)
    if (isinstance(ssq_freqs, np.ndarray) and
            infer_scaletype(ssq_freqs)[0] != 'linear'):
        raise ValueError("`ssq_freqs` must be linearly distributed "
                         "for `ssq_stft`")
    fused = (squeezing == 'sum') and not get_w and ssq_freqs is None
    if fused:
        lib = Bk.require_cuda()
        call = _get_call(N, window, n_fft, win_len, hop_len, fs, padtype, modulated, dtype)
        if gamma is None:
            gamma = 10 * (EPS64 if call.dtype == 'float64' else EPS32)
        Sfs = call.Sfs.copy()
        desc = call.reassign_desc(flipud, gamma, make_reassign_desc)
        xd = Bk.to_device(x, call.dtype)
        x2 = xd if xd.ndim == 2 else xd.unsqueeze(0)
        B = x2.shape[0]
        outs = call.outputs(B, 3 if get_dWx else 2)
        Sx, Tx = outs[0], outs[1]
        dSx = outs[2] if get_dWx else None
        _lib.check(lib.ssqb_ssq_stft_exec(C.byref(call.desc), C.byref(desc),
                                          x2.data_ptr(), B, Sx.data_ptr(),
                                          Tx.data_ptr(), Bk.ptr(dSx),
                                          Bk.stream_ptr()))
        if x.ndim == 1:
            Sx, Tx = Sx[0], Tx[0]
            dSx = dSx[0] if get_dWx else None
        w = None
        ssq_freqs = Sfs[::-1].copy() if flipud else Sfs.copy()
        Sfs_out = call.Sfs_tensor() if astensor else Sfs
    else:
        Sx, dSx = stft(x, window, n_fft=n_fft, win_len=win_len, hop_len=hop_len,
                       fs=fs, padtype=padtype, modulated=modulated, derivative=True,
                       dtype=dtype)
        rdt = Bk.dtype_of_complex(Sx)
        n_rows = Sx.shape[-2]
        Sfs = np.linspace(0, .5 * fs, n_rows, dtype=rdt)
        if gamma is None:
            gamma = 10 * (EPS64 if rdt == 'float64' else EPS32)
        w = phase_stft(Sx, dSx, Sfs, gamma) if get_w else None
        if ssq_freqs is None:
            ssq_freqs = Sfs
        Tx, ssq_freqs = ssqueeze(Sx, w, squeezing=squeezing, ssq_freqs=ssq_freqs,
                                 Sfs=Sfs, flipud=flipud, gamma=gamma,
                                 dWx=None if get_w else dSx, maprange='maximal',
                                 transform='stft')
        if not get_dWx:
            dSx = None
        Sfs_out = torch.as_tensor(Sfs, device='cuda') if astensor else Sfs

    if not astensor:
        Tx, Sx, w, dSx = [Bk.finish(g, False) for g in (Tx, Sx, w, dSx)]
    if get_w and get_dWx:
        return Tx, Sx, ssq_freqs, Sfs_out, w, dSx
    elif get_w:
        return Tx, Sx, ssq_freqs, Sfs_out, w
    elif get_dWx:
        return Tx, Sx, ssq_freqs, Sfs_out, dSx
    return Tx, Sx, ssq_freqs, Sfs_out


def phase_stft(Sx, dSx, Sfs, gamma=None, parallel=None):
    """STFT phase transform `w[u, k] = |Sfs[u] - Im(dSx / Sx) / (2 pi)|`."""
    if gamma is None:
        gamma = 10 * (EPS64 if Bk.dtype_of_complex(Sx) == 'float64' else EPS32)
    return phase_stft_gpu(Sx, dSx, Sfs, gamma)


def issq_stft(Tx, window=None, cc=None, cw=None, n_fft=None, win_len=None,
              hop_len=1, modulated=True):
    """Inverse synchrosqueezed STFT (reference `_ssq_stft.py:139-198`): sum of `Tx.real`
    over frequency rows (or over the component bands `cc +- cw`) times
    2 / window[n_fft // 2].  Only `hop_len=1`, `modulated=True`, as in the reference."""
    from ._ssq_cwt import _invert_plane
    if not modulated:
        raise ValueError("inversion with `modulated == False` "
                         "is unsupported.")
    if hop_len != 1:
        raise ValueError("inversion with `hop_len != 1` is unsupported.")
    n_fft = n_fft or (Tx.shape[0] - 1) * 2
    win_len = win_len or n_fft
    window = get_window(window, win_len, n_fft=n_fft)
    _check_NOLA(window, hop_len)
    if abs(np.argmax(window) - len(window) // 2) > 1:
        WARN("`window` maximum not centered; results may be inaccurate.")
    return _invert_plane(Tx, cc, cw, 2 / window[len(window) // 2])
