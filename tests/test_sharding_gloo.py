"""world_size-2 gloo test of the multi-GPU host logic (block partition + final gather); the per-shard compute is
stood in for by the oracle right-hand side, since this tier has no GPU."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from chemtools_b200.sharding import gather_rows, shard_range
    from oracle import oracle as O
    om = O.Mechanism("gri12")
    rng = np.random.default_rng(5)
    T = rng.uniform(900, 2000, n); p = np.full(n, 101325.0)
    Y = rng.random((n, om.K)); Y /= Y.sum(1, keepdims=True)
    lo, hi = shard_range(n, rank, world)
    st = np.concatenate([T[:, None], Y], 1)
    ydot = om.rhs_batch("Const_Pres", T[lo:hi], p[lo:hi], Y[lo:hi], st[lo:hi])[0]
    full = gather_rows(ydot, n, rank, world, dist)
    idx = gather_rows(np.arange(lo, hi, dtype=np.int64)[:, None], n, rank, world, dist)
    if rank == 0:
        np.savez(out_path, full=full, idx=idx)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 64])
def test_block_partition_and_gather(tmp_path, oracle, n):
    import torch.multiprocessing as mp
    out = str(tmp_path / "g.npz")
    mp.spawn(_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    r = np.load(out)
    assert np.array_equal(r["idx"][:, 0], np.arange(n))
    om = oracle.Mechanism("gri12")
    rng = np.random.default_rng(5)
    T = rng.uniform(900, 2000, n); p = np.full(n, 101325.0)
    Y = rng.random((n, om.K)); Y /= Y.sum(1, keepdims=True)
    ref = om.rhs_batch("Const_Pres", T, p, Y, np.concatenate([T[:, None], Y], 1))[0]
    assert np.array_equal(r["full"], ref)


def test_shard_range_properties():
    from chemtools_b200.workloads import shard_range
    for n in (0, 1, 5, 4096, 16777216):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
            for r, (lo, hi) in enumerate(parts):
                if hi > lo:  # reactor i -> rank floor(i * world / n)
                    assert (lo * w) // n == r and ((hi - 1) * w) // n == r
