#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Benchmark of the ssq_cwt hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (B200)
    python bench.py --impl reference --steps K --warmup W    # CPU reference arm

Workload (`config.workload`): BASELINE configs[1] -- ssq_cwt, Morlet (mu=13.4),
300 log scales, float32, N=160 000 synthetic linear chirps, `--batch`
signals per GPU per step (default 1).  One "step" = one full pass of the hot path
(pad -> FFT -> 300 x (wavelet multiply, inverse FFT, derivative) -> phase transform
-> reassignment) over the batch.  With N GPUs every rank processes its own batch
(the path shards over the signal axis; no data-path collective): weak scaling.

Numbers on the JSON line:
  value      Msamples/s, inputs resident in HBM, CUDA events around the K steps,
             barrier + synchronize on both sides, max over ranks
  e2e        same metric through the C ABI with HOST buffers (pinned): H2D of x and
             D2H of Tx, Wx inside the timed region
  roofline   dominant kernel (inverse pass 2 + fused epilogue): algorithmic bytes
             per launch / mean launch duration (CUDA events on the launch stream,
             separate profiling pass) against MEASURED_PEAKS.json
  cpu_baseline  the oracle port (NumPy/SciPy pocketfft + compiled C reassignment)
             on this host's cores, same workload, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SIG = 160_000
NA = 300
BYTES_PER_SAMPLE = 4 * (1 + 4 * NA)          # read x, write Tx and Wx
WORKLOAD = "ssq_cwt morlet(mu=13.4) 300 log scales float32 N=160000 (BASELINE configs[1])"


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.lines, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [c.strip() for c in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def chirp(N, b=0, dtype='float32'):
    """Unit-amplitude linear chirp, fs=1, seeded per signal index."""
    u, v = np.random.default_rng(1234 + b).random(2)
    f0, f1 = 0.02 + 0.03 * u, 0.20 + 0.20 * v
    t = np.arange(N) / N
    return np.cos(2 * np.pi * (f0 * N * t + 0.5 * (f1 - f0) * N * t**2)).astype(dtype)


def make_batch(B, rank):
    return np.stack([chirp(N_SIG, rank * B + b, 'float32') for b in range(B)])


def bench_scales_product(wavelet):
    """SURVEY 8d scale recipe with the PRODUCT's own host logic."""
    from ssqueezepy_b200.utils import cwt_scalebounds
    mn, mx = cwt_scalebounds(wavelet, N_SIG, preset='maximal')
    nv = int(np.ceil(NA / np.log2(mx / mn)))
    p0 = int(np.floor(nv * np.log2(mn)))
    return 2 ** (np.arange(p0, p0 + NA) / nv)


# ---------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_note=None):
    """Time the oracle port of the reference's SSQ_PARALLEL path on host cores."""
    import multiprocessing
    from oracle import ssq_oracle as O
    cores = multiprocessing.cpu_count()
    os.environ['OMP_NUM_THREADS'] = str(cores)     # torchrun exports 1; the C loop is the
                                                   # reference's numba prange over all cores
    wav = O.OracleWavelet('morlet', 'float32')
    scales = O.bench_scales(wav, N_SIG, NA)
    x = O.chirp(N_SIG, 0, 'float32')
    use_c = O.c_reassign_available()
    for _ in range(max(warmup, 1)):
        O.ssq_cwt(x, wav, scales, workers=cores, use_c=use_c)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.ssq_cwt(x, wav, scales, workers=cores, use_c=use_c)
        ts.append(time.perf_counter() - t0)
    mean = float(np.mean(ts))
    return {"value": N_SIG / mean / 1e6, "unit": "Msamples/s", "cores": cores,
            "kind": "port",
            "sample": sample_note or ("%d calls of ssq_cwt on one 160k-sample chirp, 300 scales, "
                                      "float32; scipy.fft workers=%d, %s reassignment; wavelet "
                                      "filter bank cached (reference Psih cache)"
                                      % (steps, cores, "OpenMP C" if use_c else "NumPy")),
            "ms_per_step": mean * 1e3, "min_ms": float(np.min(ts)) * 1e3}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = min(args.steps, 5)
    base = cpu_reference_run(steps, min(args.warmup, 2))
    line = {"metric": "ssq_cwt Msamples/s (300 scales, N=160k, fp32)", "value": base["value"], "unit": "Msamples/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 2), "ms_per_step": base["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_step": 1,
                       "note": "reference CPU path (oracle port) on host cores"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "Msamples/s",
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import ctypes as C
    import ssqueezepy_b200 as S
    from ssqueezepy_b200 import _lib
    from ssqueezepy_b200._ssq_cwt import ssq_cwt_host_params
    from ssqueezepy_b200.algos import make_reassign_desc
    from ssqueezepy_b200.utils.common import EPS32, p2up

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    B = args.batch
    lib = _lib.load(require_device=True)

    wav = S.Wavelet('morlet')
    scales = bench_scales_product(wav)
    n_up, n1, _ = p2up(N_SIG)
    hp = ssq_cwt_host_params(N_SIG, wav, scales, 'log', 'peak', True, 1.)
    plan = S.CwtPlan.get(wav, hp['scales'], N_SIG, n_up, n1, 'reflect', 1.)
    desc = make_reassign_desc(hp['ssq_freqs'], hp['const'], plan.na, hp['logscale'], True,
                              10 * EPS32, 'float32')
    plan.set_reassign(desc, 'bench')

    x_host = torch.from_numpy(make_batch(B, rank)).pin_memory()
    x_dev = x_host.cuda()
    Wx = torch.empty((B, NA, N_SIG), dtype=torch.complex64, device='cuda')
    Tx = torch.empty_like(Wx)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        _lib.check(lib.ssqb_ssq_cwt_exec(plan.handle, x_dev.data_ptr(), B, Wx.data_ptr(),
                                         Tx.data_ptr(), None, stream))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: device-resident inputs ----------------------------------------
    for _ in range(args.warmup):
        step_device()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    launches = _lib.launch_count() - launches0            # kernels of the timed region only
    if rank == 0:
        # the timed region is ~10 ms, shorter than nvidia-smi's sampling period: keep the
        # same step loop running for another 0.5 s so the clocks line has enough samples
        t_probe = time.perf_counter() + 0.5
        while time.perf_counter() < t_probe:
            for _ in range(20):
                step_device()
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region + 0.5 s of the same step loop"
    sync_all()
    ms_per_step = total_ms / args.steps
    value = world * B * N_SIG / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host buffers through the C ABI (H2D + compute + D2H) -----------------
    Wx_h = torch.empty((B, NA, N_SIG), dtype=torch.complex64).pin_memory()
    Tx_h = torch.empty((B, NA, N_SIG), dtype=torch.complex64).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))

    def step_host():
        _lib.check(lib.ssqb_ssq_cwt_exec_host(plan.handle, x_host.data_ptr(), B,
                                              Wx_h.data_ptr(), Tx_h.data_ptr(), None, stream))
    step_host()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device='cuda',
                         dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = world * B * N_SIG / float(t_e2e.item()) / 1e6
    del Wx_h, Tx_h

    # ---- extra (not the headline): 8 signals per GPU per step, same plan -----------------
    batch8 = None
    if B == 1 and not args.no_batch8:
        B8 = 8
        x8 = torch.from_numpy(make_batch(B8, rank)).cuda()
        Wx8 = torch.empty((B8, NA, N_SIG), dtype=torch.complex64, device='cuda')
        Tx8 = torch.empty_like(Wx8)

        def step8():
            _lib.check(lib.ssqb_ssq_cwt_exec(plan.handle, x8.data_ptr(), B8, Wx8.data_ptr(),
                                             Tx8.data_ptr(), None, stream))
        for _ in range(3):
            step8()
        sync_all()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(5):
            step8()
        f1.record()
        sync_all()
        ms8 = torch.tensor([f0.elapsed_time(f1) / 5], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms8, op=dist.ReduceOp.MAX)
        ms8 = float(ms8.item())
        batch8 = {"value": world * B8 * N_SIG / (ms8 * 1e-3) / 1e6, "unit": "Msamples/s",
                  "ms_per_step": ms8, "batch_per_gpu_per_step": B8,
                  "hbm_frac": BYTES_PER_SAMPLE * N_SIG * B8 / (ms8 * 1e-3) / 1e9 / _peaks()[0]}
        del x8, Wx8, Tx8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (separate profiling pass, rank 0) ------------
    peak, peak_src = _peaks()
    _lib.check(lib.ssqb_cwt_plan_set_profiling(plan.handle, 1))
    for _ in range(3):
        step_device()
    torch.cuda.synchronize()
    pms = (C.c_double * 3)(); pl = (C.c_longlong * 3)(); pr = (C.c_longlong * 3)()
    _lib.check(lib.ssqb_cwt_plan_get_profile(plan.handle, pms, pl, pr))
    _lib.check(lib.ssqb_cwt_plan_set_profiling(plan.handle, 0))
    kinds = ['fwd_fft_passes', 'inverse_pass1', 'inverse_pass2_epilogue']
    prof = {k: {"ms_total": pms[i], "launches": int(pl[i]), "rows": int(pr[i])}
            for i, k in enumerate(kinds)}
    dom = 2 if pms[2] >= pms[1] else 1
    rows_per_launch = pr[dom] / max(pl[dom], 1)
    alg_bytes_launch = BYTES_PER_SAMPLE * N_SIG * rows_per_launch / NA
    dur = pms[dom] / max(pl[dom], 1) * 1e-3
    achieved = alg_bytes_launch / dur / 1e9
    # whole-step figure too (all kernels + memset), the number the target is quoted on
    step_gbs = BYTES_PER_SAMPLE * N_SIG * B / (ms_per_step * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    if os.path.isfile(tpath):      # DRAM bytes / launch of the row kernels (ncu --set full)
        with open(tpath) as f:
            tr = json.load(f)
        rk = [v for k, v in tr.items() if 'cwt_rows_kernel' in k]
        if rk and dom == 2:
            traffic = 1e6 * sum(v['dram_read_MB'] + v['dram_write_MB'] for v in rk) / len(rk)
    roofline = {"bound": "hbm", "kernel": kinds[dom], "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "ncu per-kernel replay: outputs smaller than the 126 MB L2 are "
                                "still cached when a replay ends, so DRAM writes are under-counted "
                                "(profiles/README.md); reads <= the bytes a launch must read",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_launch,
                "mean_launch_ms": dur * 1e3,
                "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak},
                "kernel_share_of_step": {k: prof[k]["ms_total"] / max(sum(pms), 1e-12)
                                         for k in kinds},
                "profile": prof}

    cpu = cpu_reference_run(3, 1)
    line = {"metric": "ssq_cwt Msamples/s (300 scales, N=160k, fp32)", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu_per_step": B,
                       "outputs": "Tx, Wx complex64 [B,300,160000] left on the producing GPU",
                       "l2": "each step streams %.0f MB of outputs (> 126 MB L2) between "
                             "re-uses of any line" % (BYTES_PER_SAMPLE * N_SIG * B / 1e6),
                       "parallelism": "batch-sharded x%d, no collective" % world},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": "Msamples/s",
                    "h2d_bytes_per_step": int(B * N_SIG * 4),
                    "d2h_bytes_per_step": int(2 * B * NA * N_SIG * 8),
                    "note": "ssqb_ssq_cwt_exec_host: pinned host x in, Tx and Wx copied back"},
            "roofline": roofline, "batch8": batch8,
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=1, help='signals per GPU per step')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-batch8', action='store_true', help='skip the extra 8-signal measurement')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
