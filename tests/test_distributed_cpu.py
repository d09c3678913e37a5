"""world_size-2 gloo test of the multi-rank logic (no GPU): sequence sharding, the packed
statistics all-reduce and the replicated M-step.  The oracle's E-step stands in for the
device E-step, behind the same get/set-stats interface the GPU context exposes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dnbc_scala_b200.distributed import StatsAllReduce, shard_range
from tests.util import perturb, random_model, sample


class OracleBackedContext:
    """Duck-typed stand-in for _native.Context: only the members StatsAllReduce / em_iteration use."""
    device = 0

    def __init__(self, orc, model, data):
        self.orc, self.model, self.data, self.stats = orc, model, data, None

    def stats_size(self):
        return self.orc.stats_size(self.model)

    def em_estep(self):
        self.stats = self.orc.estep(self.model, self.data)

    def em_get_stats(self):
        return self.stats

    def em_set_stats(self, st):
        self.stats = np.array(st, np.float64)

    def em_mstep(self, total):
        self.orc.mstep(self.model, self.stats, total)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from dnbc_scala_b200.distributed import em_iteration
    m = random_model(orc, 6, [5, 3], [2, 1], 1)
    data = sample(orc, m, [40, 7, 1, 23, 40, 15, 9], 2)
    q = perturb(orc, m, 3)
    lo, hi = shard_range(data.S, rank, world)
    off = data.off[lo:hi + 1] - data.off[lo]
    p0, p1 = data.off[lo], data.off[hi]
    shard = orc.Data(off, data.disc[:, p0:p1], data.cont[:, p0:p1], data.labels[p0:p1])
    ctx = OracleBackedContext(orc, q, shard)
    allreduce = StatsAllReduce(ctx)
    lls = []
    for _ in range(3):
        em_iteration(ctx, allreduce, data.S)
        lls.append(float(ctx.stats[0]))
    want = perturb(orc, m, 3)
    ll_want = orc.em_iterate(want, data, 3)
    ok = (np.allclose(lls, ll_want, rtol=1e-12) and np.allclose(q.A, want.A, rtol=1e-10, atol=1e-14)
          and np.allclose(q.mu, want.mu, rtol=1e-10) and np.allclose(q.var, want.var, rtol=1e-9)
          and np.allclose(q.B, want.B, rtol=1e-10, atol=1e-14) and np.allclose(q.pi, want.pi, rtol=1e-10))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    for S in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            parts = [shard_range(S, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == S
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(hi - lo for lo, hi in parts) <= -(-S // world)


def test_two_rank_em_equals_single_process(oracle):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        assert out.get(0) is True and out.get(1) is True
