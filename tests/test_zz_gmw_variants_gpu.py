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
