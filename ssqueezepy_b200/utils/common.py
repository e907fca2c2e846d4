# -*- coding: utf-8 -*-
"""Small host utilities (reference: ssqueezepy/utils/common.py)."""
import logging
import numpy as np

logging.basicConfig(format='')
WARN = lambda msg: logging.warning("WARNING: %s" % msg)
NOTE = lambda msg: logging.warning("NOTE: %s" % msg)
pi = np.pi
EPS32 = np.finfo(np.float32).eps
EPS64 = np.finfo(np.float64).eps

__all__ = ['WARN', 'NOTE', 'pi', 'EPS32', 'EPS64', 'p2up', 'pad_geometry',
           'padsignal', 'assert_is_one_of']

PADTYPES = ('reflect', 'symmetric', 'replicate', 'wrap', 'zero')


def assert_is_one_of(x, name, supported, e=ValueError):
    if x not in supported:
        raise e("`{}` must be one of: {} (got {})".format(
            name, ', '.join(supported), x))


def p2up(n):
    """(n_up, n1, n2): power of two `2**(1 + round(log2 n))` and the left / right
    pad lengths that centre the signal (reference common.py:32-51)."""
    n_up = int(2 ** (1 + np.round(np.log2(n))))
    n2 = int((n_up - n) // 2)
    return n_up, int(n_up - n - n2), n2


def pad_geometry(N, padlength=None):
    """(n_up, n1, n2) for the default power-of-two or an explicit `padlength`
    (odd totals put the extra sample on the left; reference common.py:108-120)."""
    if padlength is None:
        return p2up(N)
    n_up = int(padlength)
    n2 = (n_up - N) // 2
    n1 = n2 if (n_up - N) % 2 == 0 else n2 + 1
    return n_up, int(n1), int(n2)


def padsignal(x, padtype='reflect', padlength=None, get_params=False):
    """Pad `x` ([N] or [B, N], numpy or torch) along the last axis.  Host/torch
    convenience with the reference's semantics (common.py:54-158); the transforms
    themselves pad inside their kernels."""
    assert_is_one_of(padtype, 'padtype', PADTYPES)
    if not hasattr(x, 'ndim'):
        raise TypeError("`x` must be a numpy array or torch Tensor "
                        "(got %s)" % type(x))
    if x.ndim not in (1, 2):
        raise ValueError("`x` must be 1D or 2D (got x.ndim == %s)" % x.ndim)
    is_np = isinstance(x, np.ndarray)
    xn = x if is_np else x.detach().cpu().numpy()
    N = xn.shape[-1]
    n_up, n1, n2 = pad_geometry(N, padlength)
    t = np.arange(n_up) - n1
    if padtype == 'zero':
        idx = np.clip(t, 0, N - 1)
    elif padtype == 'reflect':
        P = max(2 * (N - 1), 1)
        m = np.mod(t, P)
        idx = np.where(m < N, m, P - m)
    elif padtype == 'symmetric':
        m = np.mod(t, 2 * N)
        idx = np.where(m < N, m, 2 * N - 1 - m)
    elif padtype == 'replicate':
        idx = np.clip(t, 0, N - 1)
    else:
        idx = np.mod(t, N)
    xp = xn[..., idx]
    if padtype == 'zero':
        xp = xp * ((t >= 0) & (t < N))
    xp = np.ascontiguousarray(xp.astype(xn.dtype, copy=False))
    if not is_np:
        import torch
        xp = torch.as_tensor(xp, device=x.device)
    return (xp, n_up, n1, n2) if get_params else xp
