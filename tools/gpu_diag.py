# -*- coding: utf-8 -*-
"""GPU bring-up diagnostics (not a test): prints per-stage errors and a first
timing so that one gpurun call tells where a failure comes from."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import ssqueezepy_b200 as S
from oracle import ssq_oracle as O


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def np_(t):
    return t.detach().cpu().numpy()


def stage(name, fn):
    try:
        t0 = time.time(); out = fn(); torch.cuda.synchronize()
        print("[%-34s] %s  (%.2fs)" % (name, out, time.time() - t0), flush=True)
    except Exception:
        print("[%-34s] EXCEPTION\n%s" % (name, traceback.format_exc()), flush=True)


def fwd(N, dtype):
    rng = np.random.default_rng(N)
    x = rng.standard_normal((2, N)).astype(dtype)
    wav = S.Wavelet('morlet', dtype=dtype)
    n_up, n1, _ = S.utils.p2up(N)
    plan = S.CwtPlan.get(wav, np.array([4., 8.]), N, n_up, n1, 'reflect', 1.)
    xh = np_(plan.debug_xh(x))
    ref = np.fft.fft(O.padsignal(x.astype(np.float64))[0], axis=-1) / n_up
    return "n_up=%d relerr=%.3e" % (n_up, rel(xh, ref))


def identity(N, dtype):
    """table wavelet == 1 -> Wx rows must equal x (ifft(fft(xp)) unpadded)."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal(N).astype(dtype)
    wav = S.Wavelet(lambda w: np.ones_like(np.asarray(w, dtype=dtype)), dtype=dtype)
    Wx, sc = S.cwt(x, wav, scales=np.array([2., 4., 8.]))
    W = np_(Wx)
    # nyquist halving perturbs slightly: compare against oracle-like numpy
    xp, n_up, n1, _ = O.padsignal(x.astype(np.float64))
    ps = np.ones(n_up); ps[n_up // 2] = .5
    ref = np.fft.ifft(np.fft.fft(xp) * ps)[n1:n1 + N]
    return "relerr=%.3e imag=%.2e" % (rel(W[1], ref), np.abs(W[1].imag).max())


def cwt_case(N, na, dtype, name):
    x = O.chirp(N, 0, dtype)
    ow = O.OracleWavelet(name, dtype, **({'beta': 12, 'gamma': 3} if name == 'gmw' else {}))
    w = S.Wavelet((name, {'dtype': dtype, **({'beta': 12, 'gamma': 3} if name == 'gmw' else {})}))
    scales = O.bench_scales(ow, N, na)
    Wr, _, dWr = O.cwt(x, ow, scales)
    Wx, sc, dWx = S.cwt(x, w, scales=scales, derivative=True)
    return "Wx %.3e dWx %.3e" % (rel(np_(Wx), Wr), rel(np_(dWx), dWr))


def ssq_case(N, na, dtype):
    x = O.chirp(N, 0, dtype)
    ow = O.OracleWavelet('morlet', dtype)
    w = S.Wavelet('morlet', dtype=dtype)
    scales = O.bench_scales(ow, N, na)
    Tr, Wr, fr, sr = O.ssq_cwt(x, ow, scales)
    Tx, Wx, f, sc, dWx = S.ssq_cwt(x, w, scales=scales, get_dWx=True)
    gamma = 10 * (O.EPS64 if dtype == 'float64' else O.EPS32)
    st, nv = O.infer_scaletype(np_(sc))
    T2 = O.ssqueeze_fused(np_(Wx), np_(dWx), fr[::-1], O.cwt_const(np_(sc), st, nv), True, True, gamma)
    return ("Wx %.3e Tx(vs oracle e2e) %.3e Tx(vs oracle on my Wx) %.3e freqs_eq %s colsum %.3e" %
            (rel(np_(Wx), Wr), rel(np_(Tx), Tr), rel(np_(Tx), T2), np.array_equal(f, fr),
             rel(np_(Tx).sum(0), Tr.sum(0))))


def timing(N, na, B, dtype, name='morlet', iters=5):
    x = np.stack([O.chirp(N, b, dtype) for b in range(B)])
    opts = {'dtype': dtype}
    if name == 'gmw':
        opts.update(beta=12, gamma=3)
    ow = O.OracleWavelet(name, dtype, **{k: v for k, v in opts.items() if k != 'dtype'})
    w = S.Wavelet((name, opts))
    scales = O.bench_scales(ow, N, na)
    xd = torch.as_tensor(x, device='cuda')
    for _ in range(2):
        out = S.ssq_cwt(xd, w, scales=scales)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = S.ssq_cwt(xd, w, scales=scales)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    by = B * N * (4 if dtype == 'float32' else 8) * (1 + 4 * na)
    return "%.3f ms/call  %.1f Msamples/s  alg %.2f TB/s" % (ms, B * N / ms / 1e3, by / ms / 1e9)


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0), flush=True)
    for N, dt in [(8, 'float32'), (100, 'float32'), (1500, 'float32'), (1500, 'float64'),
                  (10000, 'float32'), (160000, 'float32'), (1 << 20, 'float64')]:
        stage("fwd fft N=%d %s" % (N, dt), lambda: fwd(N, dt))
    for N, dt in [(100, 'float32'), (1500, 'float32'), (1500, 'float64'), (10000, 'float32')]:
        stage("identity N=%d %s" % (N, dt), lambda: identity(N, dt))
    for N, na, dt, nm in [(1500, 48, 'float32', 'morlet'), (1500, 40, 'float64', 'gmw'),
                          (10000, 300, 'float32', 'morlet'), (10000, 300, 'float32', 'gmw')]:
        stage("cwt N=%d na=%d %s %s" % (N, na, dt, nm), lambda: cwt_case(N, na, dt, nm))
    for N, na, dt in [(1500, 48, 'float32'), (1500, 48, 'float64'), (10000, 300, 'float32')]:
        stage("ssq N=%d na=%d %s" % (N, na, dt), lambda: ssq_case(N, na, dt))
    stage("time C1-like ssq 10k", lambda: timing(10000, 300, 1, 'float32'))
    stage("time C2 ssq 160k", lambda: timing(160000, 300, 1, 'float32'))
    stage("time C4/8 B=8 gmw", lambda: timing(160000, 300, 8, 'float32', 'gmw', iters=3))
    stage("time C5/8 f64 2^20 x512", lambda: timing(1 << 20, 512, 1, 'float64', 'gmw', iters=2))
    print("launches:", S.launch_count())
