# -*- coding: utf-8 -*-
"""Parity AT THE BASELINE SHAPES (BASELINE.json configs C1, C2, C4, C5).

Two independent checkers:
 (1) the REAL reference (ssqueezepy 0.6.6 run by tests/golden/make_golden_shapes.py):
     every row of its `Wx`, `dWx` is pinned through 8 fixed +-1 projections and its
     2-norm, a few rows are stored decimated, `Tx` through column sums / non-zero counts;
 (2) the oracle on the box's host cores: whole planes (all rows), and the
     reassignment pattern of the fused epilogue -- `Tx != 0` must EQUAL the pattern of
     the oracle's column-ordered `ssqueeze` applied to the CUDA `Wx, dWx` (bit-exact
     bin indices; SURVEY 8c contract 2/3, algos.py:912-924).
Tolerances: 1e-5 (float32) / 1e-12 (float64) per row (+ 10 ulp of the strongest row, see `_bound`).
"""
import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle import ssq_oracle as O

pytestmark = pytest.mark.gpu
TOL = {'float32': 1e-5, 'float64': 1e-12}
NPROJ = 8


@pytest.fixture(scope='module')
def S():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ssqueezepy_b200 as S_
    return S_


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)


def _signs(N, seed=99):
    return np.random.default_rng(seed).integers(0, 2, size=(N, NPROJ)).astype(np.float64) * 2 - 1


def _proj(plane, R):
    """[na, N] device plane -> [na, NPROJ] complex128 projections, row chunks on the GPU."""
    import torch
    Rc = torch.as_tensor(R, device=plane.device).to(torch.complex128)
    out = []
    for r0 in range(0, plane.shape[0], 32):
        out.append((plane[r0:r0 + 32].to(torch.complex128) @ Rc).cpu())
    return torch.cat(out).numpy()


EPS = {'float32': float(np.finfo(np.float32).eps), 'float64': float(np.finfo(np.float64).eps)}


def _bound(nrm, dtype):
    """Per-row error bound: TOL x the row's own 2-norm + 10 ulp of the strongest row.
    The second term is the float rounding floor of ANY implementation in that dtype: the
    REFERENCE's own float32 rows are 2e-5 .. 9e-5 (relative, per row) away from a float64
    evaluation wherever a row carries < 1 % of the strongest row's energy (rounding noise
    of the forward FFT in bins the signal barely reaches; measured on C1 with the fixtures
    of make_golden_shapes.py: |d| ~ 1e-7 of the strongest row's norm), so a purely
    relative per-row bound is only meaningful for rows above that floor."""
    return TOL[dtype] * nrm + 10 * EPS[dtype] * nrm.max()


def _row_err(P, Pref, nrm, dtype='float32'):
    """rms over the projections of |dP| (an estimate of ||d row||_2), in units of _bound"""
    d = np.sqrt((np.abs(P - Pref) ** 2).mean(1))
    return d / _bound(nrm, dtype)


def _pair(cfg, S):
    if cfg in ('C1', 'C2'):
        return S.Wavelet('morlet'), O.OracleWavelet('morlet', 'float32')
    if cfg == 'C4':
        return (S.Wavelet(('gmw', {'beta': 12, 'gamma': 3})),
                O.OracleWavelet('gmw', 'float32', beta=12, gamma=3))
    return (S.Wavelet(('gmw', {'beta': 12, 'gamma': 3, 'dtype': 'float64'})),
            O.OracleWavelet('gmw', 'float64', beta=12, gamma=3))


def _check_vs_reference(g, i, Tx, Wx, dWx, dtype, N):
    """CUDA planes of one signal against the real reference's reductions (index i)."""
    tol = TOL[dtype]
    R = _signs(N, int(g['sign_seed']))
    eW = _row_err(_proj(Wx, R), g[f'Wx_proj{i}'], g[f'Wx_norm{i}'], dtype)
    eD = _row_err(_proj(dWx, R), g[f'dWx_proj{i}'], g[f'dWx_norm{i}'], dtype)
    assert eW.max() < 1, ("Wx row %d: %.2f of the bound" % (eW.argmax(), eW.max()))
    assert eD.max() < 1, ("dWx row %d: %.2f of the bound" % (eD.argmax(), eD.max()))
    rows, dec = g['rows_kept'], int(g['dec'])
    for nm, P in (('Wx', Wx), ('dWx', dWx)):
        got = _np(P[rows.tolist()][:, ::dec])
        ref = g[f'{nm}_rows{i}']
        floor = 10 * EPS[dtype] * g[f'{nm}_norm{i}'].max() / np.sqrt(dec)   # see _bound
        for k in range(len(rows)):
            err = np.linalg.norm(got[k] - ref[k])
            assert err < tol * np.linalg.norm(ref[k]) + floor, (nm, int(rows[k]), err)
    # Tx of the reference itself: flip-invariant quantities (its own float32 vs float64
    # runs differ 5.6e-4 norm-wise in Tx)
    cs = _np(Tx.sum(0))
    assert relerr(cs, g[f'Tx_colsum{i}']) < 5e-5
    nnz = _np((Tx != 0).sum(1))
    ref_nnz = g[f'Tx_row_nnz{i}']
    # number of occupied (bin, time) cells: sensitive to the points with |Wx| ~ gamma = 10 eps
    # (rounding noise of either implementation), hence only a loose bound
    assert abs(int(nnz.sum()) - int(ref_nnz.sum())) <= 0.03 * ref_nnz.sum()
    return float(eW.max()), float(eD.max())


def _check_bins_vs_oracle(Tx, Wx, dWx, freqs_flipped, sc, dtype, chunk=None):
    """`Tx != 0` pattern and values against the oracle's ordered reassignment of the SAME
    (CUDA) Wx, dWx; column chunks bound the host memory."""
    sq = O.ssqueeze_fused_c if O.c_reassign_available() else O.ssqueeze_fused
    st, nv = O.infer_scaletype(sc)
    const = O.cwt_const(sc, st, nv)
    gamma = 10 * (O.EPS64 if dtype == 'float64' else O.EPS32)
    ssq_freqs = np.asarray(freqs_flipped)[::-1]
    N = Wx.shape[-1]
    chunk = chunk or N
    worst = 0.
    for c0 in range(0, N, chunk):
        W = _np(Wx[:, c0:c0 + chunk].contiguous())
        dW = _np(dWx[:, c0:c0 + chunk].contiguous())
        T = _np(Tx[:, c0:c0 + chunk].contiguous())
        Tref = sq(W, dW, ssq_freqs, const, True, True, gamma)
        assert np.array_equal(T != 0, Tref != 0), "bin pattern differs in columns %d.." % c0
        worst = max(worst, relerr(T, Tref))
    assert worst < (2e-6 if dtype == 'float32' else 1e-14), worst
    return worst


# ---------------------------------------------------------------------------
@pytest.mark.parametrize('cfg', ['C1', 'C2'])
def test_shape_single_signal_vs_reference_and_oracle(S, cfg):
    """C1 / C2 (Morlet, 300 scales, float32): all 300 rows against the real reference's
    projections AND against the oracle's full planes; reassignment pattern exact."""
    g = load_golden('shape_' + cfg)
    N = 10_000 if cfg == 'C1' else 160_000
    wav, owav = _pair(cfg, S)
    x = O.chirp(N, 0, 'float32')
    Tx, Wx, freqs, sc, dWx = S.ssq_cwt(x, wav, scales=g['scales_in'], get_dWx=True)
    assert np.array_equal(_np(sc), g['scales_out'])
    assert np.array_equal(np.asarray(freqs), g['ssq_freqs'])
    _check_vs_reference(g, 0, Tx, Wx, dWx, 'float32', N)
    # whole planes vs the oracle (pocketfft float32, like the reference): every row
    Wo, sco, dWo = O.cwt(x, owav, g['scales_in'])
    Wc, dWc = _np(Wx), _np(dWx)
    for nm, got, ref in (('Wx', Wc, Wo), ('dWx', dWc, dWo)):
        num = np.linalg.norm(got - ref, axis=1)
        den = np.linalg.norm(ref, axis=1)
        e = num / _bound(den, 'float32')
        assert e.max() < 1, (nm, int(e.argmax()), float(e.max()))
    _check_bins_vs_oracle(Tx, Wx, dWx, freqs, _np(sc), 'float32')
    # cwt() alone (configs[0] is the plain transform) returns the same Wx
    W2, sc2 = S.cwt(x, wav, scales=g['scales_in'])
    assert np.array_equal(_np(W2), Wc)


def test_shape_C4_batched_gmw(S):
    """C4 per-GPU share: x[8, 160 000], GMW(12,3), 300 scales, float32."""
    import torch
    g = load_golden('shape_C4')
    N, B = 160_000, 8
    wav, owav = _pair('C4', S)
    x = np.stack([O.chirp(N, b, 'float32') for b in range(B)])
    Tx, Wx, freqs, sc, dWx = S.ssq_cwt(x, wav, scales=g['scales_in'], get_dWx=True)
    assert tuple(Wx.shape) == (B, 300, N)
    assert np.array_equal(_np(sc), g['scales_out'])
    assert np.array_equal(np.asarray(freqs), g['ssq_freqs'])
    for i, b in enumerate(g['bs'].tolist()):           # signals the reference ran
        _check_vs_reference(g, i, Tx[b], Wx[b], dWx[b], 'float32', N)
    for b in range(B):                                 # every signal: exact bins
        _check_bins_vs_oracle(Tx[b], Wx[b], dWx[b], freqs, _np(sc), 'float32')
    # batched == per-sample (reference tests/fft_test.py:559-631)
    T1, W1, *_ = S.ssq_cwt(x[3], wav, scales=g['scales_in'])
    assert torch.equal(W1, Wx[3])
    assert relerr(_np(T1), _np(Tx[3])) < 2e-6
    del Tx, Wx, dWx
    torch.cuda.empty_cache()


def test_shape_C5_float64(S):
    """C5 per-GPU share: one signal of N = 2^20, 512 scales, GMW(12,3), float64, 1e-12."""
    import torch
    g = load_golden('shape_C5')
    N = 1 << 20
    wav, owav = _pair('C5', S)
    x = O.chirp(N, 0, 'float64')
    scales = g['scales_in']
    Tx, Wx, freqs, sc, dWx = S.ssq_cwt(x, wav, scales=scales, get_dWx=True)
    assert tuple(Wx.shape) == (512, N) and str(Wx.dtype).endswith('complex128')
    assert np.array_equal(np.asarray(freqs), load_golden('host_params')['C5_ssq_freqs'][::-1])
    rows = g['rows'].tolist()
    R = _signs(N, int(g['sign_seed']))
    eW = _row_err(_proj(Wx[rows], R), g['Wx_proj'], g['Wx_norm'], 'float64')
    eD = _row_err(_proj(dWx[rows], R), g['dWx_proj'], g['dWx_norm'], 'float64')
    assert eW.max() < 1, (int(eW.argmax()), float(eW.max()))
    assert eD.max() < 1, (int(eD.argmax()), float(eD.max()))
    dec = int(g['dec'])
    gotW, gotD = _np(Wx[rows][:, ::dec]), _np(dWx[rows][:, ::dec])
    for nm, got in (('Wx', gotW), ('dWx', gotD)):
        floor = 10 * EPS['float64'] * g[nm + '_norm'].max() / np.sqrt(dec)       # see _bound
        for k in range(len(rows)):
            err = np.linalg.norm(got[k] - g[nm + '_rows'][k])
            assert err < 1e-12 * np.linalg.norm(g[nm + '_rows'][k]) + floor, (nm, rows[k], err)
    # all 512 rows against a float64 cuFFT evaluation of the oracle's filter bank, in row chunks
    xp, n_up, n1, _ = O.padsignal(x)
    xh = torch.fft.fft(torch.as_tensor(xp, device='cuda'))
    xi = torch.as_tensor(O.xi_grid(n_up, 'float64'), device='cuda')
    worst = 0.
    for r0 in range(0, 512, 16):
        psih = torch.as_tensor(owav.psih(np.asarray(scales)[r0:r0 + 16], n_up), device='cuda')
        P = psih * xh
        Wr = torch.fft.ifft(P, dim=-1)[:, n1:n1 + N]
        dWr = torch.fft.ifft(P * (1j * xi), dim=-1)[:, n1:n1 + N]
        for nm, got, ref in (('Wx', Wx[r0:r0 + 16], Wr), ('dWx', dWx[r0:r0 + 16], dWr)):
            den = torch.linalg.vector_norm(ref, dim=1)
            bound = 1e-12 * den + 10 * EPS['float64'] * float(g[nm + '_norm'].max())
            e = torch.linalg.vector_norm(got - ref, dim=1) / bound
            worst = max(worst, float(e.max()))
            assert float(e.max()) < 1, (nm, r0 + int(e.argmax()), float(e.max()))
        del psih, P, Wr, dWr
    del xh, xi
    _check_bins_vs_oracle(Tx, Wx, dWx, freqs, _np(sc), 'float64', chunk=1 << 16)
    del Tx, Wx, dWx
    torch.cuda.empty_cache()
