"""One-shot GPU script: quick kernel timing at 2^20 (used during development under gpurun)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import benchdata
from elliptic_b200 import _native as nat
from elliptic_b200.ec import EC

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
curve = sys.argv[2] if len(sys.argv) > 2 else "secp256k1"
seed = {"secp256k1": 0xE1110002, "p256": 0xE1110256, "p384": 0xE1110384}[curve]
t = time.time()
ds = benchdata.gen_ecdsa_verify(curve, n, seed=seed, cache_dir="/tmp/eb200_cache")
print("gen %.1fs" % (time.time() - t), flush=True)
ec = EC(curve)
res = []
for it in range(4):
    t = time.time()
    st = ec.verify_batch_packed(ds["e"], ds["r"], ds["s"], ds["pub"])
    wall = time.time() - t
    tm = nat.last_timing()
    tm["wall_ms"] = wall * 1e3
    tm["ok"] = bool(np.array_equal(st, ds["expected"]))
    tm["verifies_per_s_kernel"] = n / (tm["kernel_ms"] * 1e-3)
    res.append(tm)
    print(json.dumps(tm), flush=True)
