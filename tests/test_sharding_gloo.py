# -*- coding: utf-8 -*-
"""World-size-2 gloo test (CPU) of the N>1 path: batch sharding + optional gather.
The transform itself is replaced by a CPU stand-in (the kernels need a GPU); what
is tested is the host logic every rank runs: slice bounds, ragged gather, the
single-signal refusal."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from ssqueezepy_b200.distributed import shard_bounds, ssq_cwt_sharded


def test_shard_bounds_partition():
    for B in (1, 2, 7, 8, 64):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(B, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _stub(x, *a, **k):
    """stand-in transform: 'Tx' = x outer [1,2,3], 'Wx' = 2*that"""
    t = torch.as_tensor(x, dtype=torch.float32)
    T = t[:, None, :] * torch.tensor([1., 2., 3.])[None, :, None]
    return T, 2 * T, np.array([.1, .2, .3]), np.array([1., 2., 3.])


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    x = np.arange(B * 5, dtype=np.float32).reshape(B, 5)
    Tl, Wl, f, s = ssq_cwt_sharded(x, _compute=_stub)
    lo, hi = shard_bounds(B, rank, world)
    ok = Tl.shape[0] == hi - lo and torch.equal(Tl, _stub(x[lo:hi])[0]) if hi > lo else Tl.shape[0] == 0
    Tg, Wg, _, _ = ssq_cwt_sharded(x, gather=True, _compute=_stub)
    ok = ok and torch.equal(Tg, _stub(x)[0]) and torch.equal(Wg, _stub(x)[1])
    try:
        ssq_cwt_sharded(x[0], _compute=_stub)
        ok = False
    except ValueError:
        pass
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('B', [1, 5, 8])
def test_world_size_2_gloo(B):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + B
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
