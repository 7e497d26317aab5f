"""One-shot GPU script: end-to-end (pinned host buffers -> statuses) timing of eb200_ecdsa_verify_batch
for several host-pipeline chunk counts (EB200_CHUNKS).  Development aid, run under gpurun."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import benchdata
from elliptic_b200 import _native as nat
from elliptic_b200.ec import EC

n = 1 << 20
ds = benchdata.gen_ecdsa_verify("secp256k1", n, seed=0xE1110002, cache_dir="/tmp/eb200_cache")
h = {k: torch.from_numpy(ds[k]).pin_memory().numpy() for k in ("e", "r", "s", "pub")}
ec = EC("secp256k1")
for chunks in [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "4,8,16").split(",")]:
    os.environ["EB200_CHUNKS"] = str(chunks)
    walls, ks = [], []
    for it in range(12):
        t = time.perf_counter()
        st = ec.verify_batch_packed(h["e"], h["r"], h["s"], h["pub"])
        walls.append((time.perf_counter() - t) * 1e3)
        ks.append(nat.last_timing()["kernel_ms"])
    assert np.array_equal(st, ds["expected"])
    walls, ks = sorted(walls[2:]), sorted(ks[2:])
    print(json.dumps({"chunks": chunks, "wall_ms_med": walls[len(walls) // 2], "gpu_timeline_ms_med": ks[len(ks) // 2],
                      "verifies_per_s": n / (walls[len(walls) // 2] * 1e-3)}), flush=True)
