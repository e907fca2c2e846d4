# -*- coding: utf-8 -*-
"""Two ranks over NCCL (needs 2 GPUs on the box, skipped otherwise): batch-sharded ssq_cwt with
the outputs left on the producing GPU, and the optional NCCL all_gather of the shards; each
rank's shard must equal what a single GPU computes for the same signals."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import ssqueezepy_b200 as S
    from ssqueezepy_b200.distributed import shard_bounds, ssq_cwt_sharded
    from oracle import ssq_oracle as O
    N, B, na = 20_000, 5, 64
    wav = S.Wavelet('morlet')
    scales = O.bench_scales(O.OracleWavelet('morlet', 'float32'), N, na)
    x = np.stack([O.chirp(N, b, 'float32') for b in range(B)])
    Tl, Wl, f, sc = ssq_cwt_sharded(x, wav, scales=scales)
    lo, hi = shard_bounds(B, rank, world)
    Tf, Wf, *_ = S.ssq_cwt(x, wav, scales=scales)              # the whole batch on this GPU
    ok = tuple(Wl.shape) == (hi - lo, na, N) and torch.equal(Wl, Wf[lo:hi])
    ok = ok and float((Tl - Tf[lo:hi]).abs().max()) <= 1e-5 * float(Tf.abs().max())
    Tg, Wg, *_ = ssq_cwt_sharded(x, wav, scales=scales, gather=True)
    ok = ok and torch.equal(Wg, Wf) and tuple(Tg.shape) == (B, na, N)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_nccl_shard_and_gather():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29711, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(120)
    assert sorted(res) == [(0, True), (1, True)]
