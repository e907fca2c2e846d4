# -*- coding: utf-8 -*-
"""CWT through the host-evaluated generalized Morse variants (L2 norm via
`l1_norm=False`, order 2): the `psih` table path of the plan, against the reference's
output stored in tests/golden/gmw_variants.npz."""
import numpy as np
import pytest

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu


def test_cwt_gmw_l2_and_order2():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import ssqueezepy_b200 as S
    g = load_golden('gmw_variants')
    W, sc = S.cwt(g['x'], ('gmw', {'beta': 12, 'gamma': 3}), scales=g['scales'], l1_norm=False)
    assert relerr(W.cpu().numpy(), g['Wx_l2']) < 1e-5
    W2, _ = S.cwt(g['x'], ('gmw', {'beta': 12, 'gamma': 3, 'order': 2}), scales=g['scales'])
    assert relerr(W2.cpu().numpy(), g['Wx_k2']) < 1e-5


def test_cwt_order_argument():
    """`cwt(order=k)` and the average over a tuple of orders (reference
    `_cwt.py:517-610`), end to end on the device."""
    import ssqueezepy_b200 as S
    g = load_golden('gmw_variants')
    wav = ('gmw', {'beta': 12, 'gamma': 3})
    W2, sc = S.cwt(g['x'], wav, scales=g['scales'], order=2)
    assert relerr(W2.cpu().numpy(), g['Wx_order2']) < 1e-5
    W, sc, dW = S.cwt(g['x'], wav, scales=g['scales'], order=(0, 1, 2), derivative=True)
    assert relerr(W.cpu().numpy(), g['Wx_order012']) < 1e-5
    assert relerr(dW.cpu().numpy(), g['dWx_order012']) < 1e-5
    Wn, scn = S.cwt(g['x'], wav, scales=g['scales'], order=2, astensor=False)
    assert isinstance(Wn, np.ndarray) and np.array_equal(Wn, W2.cpu().numpy())


def test_ssq_cwt_order_argument():
    """`ssq_cwt(order=(0, 1))`: averaged higher-order transform, then the reassignment
    operator.  `Wx` to transform accuracy; `Tx` through the flip-invariant column sums
    (float32 bin flips at rounding level are expected)."""
    import ssqueezepy_b200 as S
    g = load_golden('gmw_variants')
    Tx, Wx, freqs, sc = S.ssq_cwt(g['x'], ('gmw', {'beta': 12, 'gamma': 3}),
                                  scales=g['scales'], order=(0, 1))
    assert relerr(Wx.cpu().numpy(), g['ssq_Wx_order01']) < 1e-5
    assert np.array_equal(np.asarray(freqs), g['ssq_freqs_order01'])
    assert relerr(Tx.cpu().numpy().sum(0), g['ssq_Tx_order01'].sum(0)) < 5e-4
