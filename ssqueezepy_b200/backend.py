# -*- coding: utf-8 -*-
"""Device plumbing: torch provides device memory and streams, nothing else.

Conventions (same as the reference's GPU mode, ssqueezepy/utils/backend.py and
_cwt.py:255-258, _ssq_cwt.py:297-300): inputs may be numpy arrays or torch
tensors; compute happens on the current CUDA device and stream; results are
CUDA tensors (`astensor=True`) or numpy arrays (`astensor=False`)."""
import numpy as np
import torch

from . import _lib

_TORCH_REAL = {'float32': torch.float32, 'float64': torch.float64}
_TORCH_CPLX = {'float32': torch.complex64, 'float64': torch.complex128}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("ssqueezepy_b200 needs a CUDA device (sm_100a); there "
                           "is no CPU fallback")
    return _lib.load(require_device=True)


def real_dtype(dtype):
    return _TORCH_REAL[str(dtype)]


def cplx_dtype(dtype):
    return _TORCH_CPLX[str(dtype)]


def dtype_code(dtype):
    return _lib.F32 if str(dtype) == 'float32' else _lib.F64


def dtype_of_complex(t):
    """'float32' / 'float64' from a complex (or real) tensor / array."""
    s = str(t.dtype)
    return 'float64' if ('128' in s or s.endswith('float64')) else 'float32'


def to_device(x, dtype=None, complex_=False):
    """Contiguous CUDA tensor of `x` (numpy / torch, any device)."""
    td = None
    if dtype is not None:
        td = cplx_dtype(dtype) if complex_ else real_dtype(dtype)
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    elif not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    x = x.to(device='cuda', dtype=td, non_blocking=False)
    return x.contiguous()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def finish(t, astensor):
    if t is None or astensor or not isinstance(t, torch.Tensor):
        return t
    return t.cpu().numpy()


def is_tensor(x):
    return isinstance(x, torch.Tensor)
