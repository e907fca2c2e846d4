"""Per-kernel-kind CUDA-event timing of one ssq_cwt configuration through the plan's
profiling hooks (ssqb_cwt_plan_set_profiling / get_profile).
Usage: python tools/profile_kinds.py N na dtype wavelet [B]   (e.g. 1048576 512 float64 gmw)"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ssqueezepy_b200 as S
from ssqueezepy_b200 import _lib
from ssqueezepy_b200._ssq_cwt import ssq_cwt_host_params
from ssqueezepy_b200.algos import make_reassign_desc
from ssqueezepy_b200.utils.common import p2up, EPS32, EPS64
from oracle import ssq_oracle as O

N, na, dtype, name = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
B = int(sys.argv[5]) if len(sys.argv) > 5 else 1
opts = {'dtype': dtype}
if name == 'gmw':
    opts.update(beta=12, gamma=3)
wav = S.Wavelet((name, opts))
ow = O.OracleWavelet(name, dtype, **{k: v for k, v in opts.items() if k != 'dtype'})
scales = O.bench_scales(ow, N, na)
lib = _lib.load(require_device=True)
n_up, n1, _ = p2up(N)
hp = ssq_cwt_host_params(N, wav, scales, 'log', 'peak', True, 1.)
plan = S.CwtPlan.get(wav, hp['scales'], N, n_up, n1, 'reflect', 1.)
desc = make_reassign_desc(hp['ssq_freqs'], hp['const'], plan.na, hp['logscale'], True,
                          10 * (EPS64 if dtype == 'float64' else EPS32), dtype)
plan.set_reassign(desc, 'prof')
x = torch.as_tensor(np.stack([O.chirp(N, b, dtype) for b in range(B)]), device='cuda')
cdt = torch.complex128 if dtype == 'float64' else torch.complex64
Wx = torch.empty((B, na, N), dtype=cdt, device='cuda'); Tx = torch.empty_like(Wx)
st = torch.cuda.current_stream().cuda_stream
run = lambda: _lib.check(lib.ssqb_ssq_cwt_exec(plan.handle, x.data_ptr(), B, Wx.data_ptr(), Tx.data_ptr(), None, st))
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
print("step %.3f ms" % (e0.elapsed_time(e1) / 2))
_lib.check(lib.ssqb_cwt_plan_set_profiling(plan.handle, 1))
run(); torch.cuda.synchronize()
pms = (C.c_double * 6)(); pl = (C.c_longlong * 6)(); pr = (C.c_longlong * 6)()
_lib.check(lib.ssqb_cwt_plan_get_profile(plan.handle, pms, pl, pr))
bpr = (4 if dtype == 'float32' else 8) * 4 * N          # algorithmic bytes per row (Wx + Tx)
for i, k in enumerate(_lib.PROFILE_KINDS):
    if pl[i]:
        print("%-44s %9.3f ms  %4d launches  %6d rows  %7.3f us/row  alg %.0f GB/s"
              % (k, pms[i], pl[i], pr[i], 1e3 * pms[i] / max(pr[i], 1), pr[i] * bpr / pms[i] / 1e6))
