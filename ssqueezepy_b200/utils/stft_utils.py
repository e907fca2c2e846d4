# -*- coding: utf-8 -*-
"""STFT framing helper (reference: ssqueezepy/utils/stft_utils.py:20-98).

The STFT kernels frame the signal themselves (no [n_fft, n_hops] copy is ever
made); `buffer` is kept as a convenience with the reference's exact semantics
and works on numpy arrays and torch tensors by index-gather."""
import numpy as np

__all__ = ['buffer']


def _frame_rows(seg_len, modulated):
    if not modulated:
        return np.arange(seg_len)
    s20 = int(np.ceil(seg_len / 2))
    s21 = s20 - 1 if (seg_len % 2 == 1) else s20
    return np.concatenate([np.arange(s21, s21 + s20), np.arange(0, s21)])


def buffer(x, seg_len, n_overlap, modulated=False, parallel=None):
    """Columns = successive length-`seg_len` slices of `x` hopping by
    `seg_len - n_overlap`; `modulated` stores each frame ifftshift-ed.
    x: [N] -> [seg_len, n_segs];  [B, N] -> [B, seg_len, n_segs]."""
    assert x.ndim in (1, 2)
    hop = seg_len - n_overlap
    n_segs = (x.shape[-1] - seg_len) // hop + 1
    idx = _frame_rows(seg_len, modulated)[:, None] + hop * np.arange(n_segs)[None]
    if isinstance(x, np.ndarray):
        return x[..., idx]
    import torch
    return x[..., torch.as_tensor(idx, device=x.device)]
