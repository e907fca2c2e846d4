# -*- coding: utf-8 -*-
"""Batched ssq_cwt in groups of signals with zero-ahead (csrc/cwt_impl.cuh `group_size`): the row
kernels of group g store the zeros of group g+1's Tx next to their own Wx stores, so only the first
group is zeroed by a kernel of its own.  Whatever the group size, the result must be the one-group
result: Wx / dWx bit-identical (same kernels, same inputs), Tx with the identical non-zero
pattern -- every contribution in the same bin, nothing lost to a late zero and nothing left from
an earlier call -- and equal up to the order of the atomic sums.  Reference behaviour:
ssqueezepy/_ssq_cwt.py:208-233 (each signal of a batch is transformed independently)."""
import os
import numpy as np
import pytest

from conftest import relerr
from oracle import ssq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def S():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ssqueezepy_b200 as S_
    return S_


def _np(t):
    return t.detach().cpu().numpy()


def _run(S, x, wav, scales, group, poison=True):
    import torch
    old = os.environ.get('SSQB_GROUP')
    os.environ['SSQB_GROUP'] = str(group)
    try:
        if poison:
            # the caching allocator hands the next call the same blocks: leave garbage behind so
            # that a Tx element nobody zeroed shows up
            junk = [torch.full((x.shape[0], len(scales), x.shape[1]), 7.0 + 3.0j,
                               dtype=torch.complex64 if x.dtype == np.float32 else torch.complex128,
                               device='cuda') for _ in range(3)]
            del junk
        Tx, Wx, freqs, sc, dWx = S.ssq_cwt(x, wav, scales=scales, get_dWx=True)
        torch.cuda.synchronize()
        return _np(Tx), _np(Wx), _np(dWx), np.asarray(freqs), _np(sc)
    finally:
        if old is None:
            os.environ.pop('SSQB_GROUP', None)
        else:
            os.environ['SSQB_GROUP'] = old


@pytest.mark.parametrize('dtype,N,B,name', [('float32', 40000, 6, 'gmw'), ('float32', 20000, 4, 'morlet'),
                                            ('float64', 20000, 4, 'gmw'), ('float32', 3000, 5, 'gmw')])
def test_groups_equal_one_group(S, dtype, N, B, name):
    opts = {'dtype': dtype}
    if name == 'gmw':
        opts.update(beta=12, gamma=3)
    wav = S.Wavelet((name, opts))
    # 0.42 .. 370: Nyquist-cut, short-block, direct and gridded rows in one plan (N >= 2^13 padded);
    # N = 3000 takes the generic two-pass kernels
    scales = 2 ** (np.arange(-10, 70) / 8.) * (4.2 if name == 'morlet' else 1.)
    x = np.stack([O.chirp(N, b, dtype) for b in range(B)])
    T0, W0, dW0, f0, sc0 = _run(S, x, wav, scales, 0)
    tol = 2e-6 if dtype == 'float32' else 1e-13
    for group in (1, 2, 3):
        T1, W1, dW1, f1, sc1 = _run(S, x, wav, scales, group)
        assert np.array_equal(W0, W1) and np.array_equal(dW0, dW1), group
        assert np.array_equal(f0, f1) and np.array_equal(sc0, sc1)
        assert np.array_equal(T0 != 0, T1 != 0), group
        assert relerr(T1, T0) < tol, group


def test_grouped_batch_matches_oracle_bins(S):
    """every signal of a grouped batch: Tx = the oracle's ordered reassignment of the CUDA Wx, dWx"""
    dtype = 'float32'
    wav = S.Wavelet(('gmw', {'beta': 12, 'gamma': 3, 'dtype': dtype}))
    N, B = 20000, 4
    scales = 2 ** (np.arange(-10, 54) / 8.)
    x = np.stack([O.chirp(N, b, dtype) for b in range(B)])
    Tx, Wx, dWx, freqs, sc = _run(S, x, wav, scales, 1)
    st, nv = O.infer_scaletype(sc)
    const = O.cwt_const(sc, st, nv)
    for b in range(B):
        Tref = O.ssqueeze_fused(Wx[b], dWx[b], freqs[::-1], const, True, True, 10 * O.EPS32)
        assert np.array_equal(Tx[b] != 0, Tref != 0), b
        assert relerr(Tx[b], Tref) < 2e-6, b
