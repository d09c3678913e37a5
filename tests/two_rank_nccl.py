"""One rank of the two-process NCCL test (tests/test_gpu_baseline_shapes.py): each process owns one
GPU and one shard; rank 0 writes the library's 128-byte unique id to a file (the JVM shim would ship
it in a Spark broadcast variable), both join with dnbc_comm_init_rank and run dnbc_em_iterate_sharded.
No torch.distributed anywhere on this path.  usage: two_rank_nccl.py RANK WORLD DIR"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dnbc_scala_b200 import _native as nat  # noqa: E402
from dnbc_scala_b200.distributed import shard_range  # noqa: E402
from oracle import oracle  # noqa: E402
from tests.util import ctx_for, perturb, random_model, sample, upload  # noqa: E402


def main():
    rank, world, d = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    m = random_model(oracle, 10, [6, 4], [3, 2], 351)
    data = sample(oracle, m, [40] * 21 + [9], 352)
    q = perturb(oracle, m, 353)
    want = q.copy()
    ll_want = oracle.em_iterate(want, data, 3)
    lo, hi = shard_range(data.S, rank, world)
    p0, p1 = int(data.off[lo]), int(data.off[hi])
    shard = oracle.Data(data.off[lo:hi + 1] - p0, data.disc[:, p0:p1], data.cont[:, p0:p1])
    id_path = os.path.join(d, "nccl_id.bin")
    if rank == 0:
        uid = nat.comm_unique_id()
        with open(id_path + ".tmp", "wb") as fh:
            fh.write(uid)
        os.replace(id_path + ".tmp", id_path)
    else:
        t0 = time.time()
        while not os.path.exists(id_path):
            if time.time() - t0 > 120:
                raise SystemExit("timed out waiting for the unique id")
            time.sleep(0.05)
        uid = open(id_path, "rb").read()
    with ctx_for(nat, q, device=rank) as ctx:
        upload(ctx, shard, labels=False, dtype=np.float32)
        ctx.comm_init_rank(world, rank, uid)
        assert ctx.comm_info() == (world, rank)
        ll = ctx.em_iterate_sharded(3, data.S)
        p = ctx.get_params()
        ctx.comm_destroy()
    np.testing.assert_allclose(ll, ll_want, rtol=1e-5)
    np.testing.assert_allclose(p["A"], want.A, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(p["B"], want.B, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(p["mu"], want.mu, rtol=2e-4, atol=2e-5)
    print(f"rank ok {rank}/{world}: loglik {ll.tolist()}")


if __name__ == "__main__":
    main()
