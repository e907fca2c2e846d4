# -*- coding: utf-8 -*-
"""Backward pass of `cwt` (torch.autograd): the adjoint kernels against autograd through a
float64 torch restatement of the same linear map (reflect pad -> fft -> psih -> ifft -> unpad,
ssqueezepy/_cwt.py:167-177), and the reference's own use (examples/reconstruction.py:38-70)."""
import numpy as np
import pytest

from oracle import ssq_oracle as O

pytestmark = pytest.mark.gpu


def _torch_cwt(x, psih, xi, n1, N, pad_idx, derivative):
    import torch
    xp = x[..., pad_idx]
    P = psih * torch.fft.fft(xp.to(torch.complex128), dim=-1)[..., None, :]
    W = torch.fft.ifft(P, dim=-1)[..., n1:n1 + N]
    if not derivative:
        return W, None
    dW = torch.fft.ifft(P * (1j * xi), dim=-1)[..., n1:n1 + N]
    return W, dW


@pytest.mark.parametrize('N,dtype,padtype,B', [(700, 'float32', 'reflect', 1), (1000, 'float64', 'reflect', 2),
                                              (512, 'float64', None, 1), (600, 'float64', None, 1)])
def test_cwt_backward_matches_torch_autograd(N, dtype, padtype, B):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ssqueezepy_b200 as S
    wav = S.Wavelet(('morlet', {'dtype': dtype}))
    owav = O.OracleWavelet('morlet', 'float64')
    na = 20
    scales = 2 ** np.linspace(2.5, 6.5, na)
    rng = np.random.default_rng(N)
    x0 = rng.standard_normal((B, N))
    if padtype is None:
        n_up, n1 = N, 0
        pad_idx = np.arange(N)
    else:
        n_up, n1, _ = S.utils.p2up(N)
        pad_idx = np.pad(np.arange(N), (n1, n_up - N - n1), mode='reflect')
    psih = torch.as_tensor(owav.psih(np.asarray(scales, dtype=dtype).astype(np.float64), n_up), device='cuda')
    xi = torch.as_tensor(O.xi_grid(n_up, 'float64'), device='cuda')
    idx = torch.as_tensor(pad_idx, device='cuda')
    wts = torch.as_tensor(rng.standard_normal((B, na, N)), device='cuda')
    for derivative in (False, True):
        xr = torch.tensor(x0, device='cuda', dtype=torch.float64, requires_grad=True)
        Wr, dWr = _torch_cwt(xr, psih, xi, n1, N, idx, derivative)
        Lr = (Wr.abs() ** 2 * wts).sum() + ((dWr.real * wts).sum() if derivative else 0.)
        Lr.backward()
        xt = torch.tensor(x0, device='cuda', dtype=getattr(torch, dtype), requires_grad=True)
        out = S.cwt(xt if B > 1 else xt[0], wav, scales=scales, padtype=padtype, derivative=derivative)
        W = out[0].reshape(B, na, N)
        L = (W.abs() ** 2 * wts.to(W.real.dtype)).sum()
        if derivative:
            L = L + (out[2].reshape(B, na, N).real * wts.to(W.real.dtype)).sum()
        L.backward()
        g, gr = xt.grad.double(), xr.grad
        err = float((g - gr).norm() / gr.norm())
        assert err < (2e-5 if dtype == 'float32' else 1e-11), (derivative, err)


def test_signal_recovery_from_scalogram_decreases_loss():
    """examples/reconstruction.py:38-70 in miniature: optimise x so that |cwt(x)| matches a target."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ssqueezepy_b200 as S
    N = 512
    wav = S.Wavelet('morlet')
    scales = 2 ** np.linspace(2.5, 6., 24)
    y = torch.as_tensor(O.chirp(N, 3, 'float32'), device='cuda')
    Sy = S.cwt(y, wav, scales=scales)[0].abs()
    torch.manual_seed(1)
    x = torch.randn(N, device='cuda')
    x = (x / x.abs().max()).requires_grad_(True)
    opt = torch.optim.Adam([x], lr=.05)
    losses = []
    for _ in range(60):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(S.cwt(x, wav, scales=scales)[0].abs(), Sy)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.2 * losses[0]
