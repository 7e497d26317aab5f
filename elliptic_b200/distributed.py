"""Multi-GPU sharding of a batch: one process per GPU (torch.distributed), no data-path
collective -- every rank verifies its own contiguous shard; the only exchange is the gather
of 1 status byte per item (BASELINE.json: "NCCL only for the final result gather").
"""
import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous block partition: rank g gets [g*n/G, (g+1)*n/G)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_status(local_status, n, world, rank, device=None, group=None):
    """all_gather of variable-size status shards (padded to the largest shard).
    local_status: uint8 numpy array of this rank's shard.  Returns the full (n,) array on every rank.
    Backend follows the process group: NCCL with `device` set (GPU box), gloo on CPU (tests)."""
    import torch
    import torch.distributed as dist
    sizes = [shard_bounds(n, world, g)[1] - shard_bounds(n, world, g)[0] for g in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=torch.uint8, device=device)
    buf[:local_status.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_status)).to(buf.device)
    out = torch.empty(pad * world, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().reshape(world, pad)
    return np.concatenate([out[g, :sizes[g]] for g in range(world)])


def verify_sharded(verify_fn, e, r, s, pub, world, rank, device=None, group=None):
    """Each rank runs `verify_fn` (e.g. EC(...).verify_batch_packed bound to its GPU) on its shard of the
    same global batch and the statuses are gathered everywhere."""
    n = e.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    local = verify_fn(e[lo:hi], r[lo:hi], s[lo:hi], pub[lo:hi]) if hi > lo else np.zeros(0, np.uint8)
    return gather_status(local, n, world, rank, device=device, group=group)
