"""Multi-GPU host logic: independent ensembles, no collective on the solve path.

Reactor i of n goes to rank floor(i * world / n) (contiguous blocks); every rank integrates its block on its own
GPU; the only communication is the final gather of per-reactor results (<= 1 KB each) to rank 0 for the results
writer.  `torch.distributed` is plumbing only (gloo on CPU hosts, nccl on GPU boxes).
"""
import numpy as np

from .workloads import shard_range  # noqa: F401  (re-exported)


def gather_rows(local, n_total, rank, world, dist=None, device=None):
    """Gather per-rank row blocks (shard_range order) on rank 0. `local`: numpy [n_local, ...]; returns the full
    array on rank 0 and None elsewhere.  Works with any initialised torch.distributed backend."""
    local = np.ascontiguousarray(local)
    if world == 1 or dist is None:
        return local
    import torch
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = np.zeros((nmax,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, bufs, dst=0)
    if rank != 0:
        return None
    return np.concatenate([bufs[r][: hi - lo].cpu().numpy() for r, (lo, hi) in enumerate(sizes)], 0)
