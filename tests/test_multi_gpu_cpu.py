"""World-size-2 test of the sharding / gather logic over gloo on CPU (the N>1 bench path uses the
same code over NCCL).  The per-shard compute is injected: here the C oracle stands in for the GPU."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import benchdata
    from elliptic_b200.distributed import verify_sharded, shard_bounds
    from oracle import c_oracle
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ds = benchdata.gen_secp256k1_verify(n, n_keys=8)
    full = verify_sharded(lambda e, r, s, p: c_oracle.verify_batch(e, r, s, p, 1), ds["e"], ds["r"], ds["s"], ds["pub"], world, rank)
    lo, hi = shard_bounds(n, world, rank)
    q.put((rank, bool(np.array_equal(full, ds["expected"])), lo, hi))
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_over_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 301                                  # odd on purpose: ragged shards
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 150, 150, 301)


def test_shard_bounds_cover_everything():
    from elliptic_b200.distributed import shard_bounds
    for n in (0, 1, 7, 1 << 20, (1 << 23) + 5):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, w, g) for g in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
