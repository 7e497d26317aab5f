"""One-shot GPU script: kernel / end-to-end rates of every accelerated path at batch N
(BASELINE.json configs 2-5).  Used under gpurun during development and for profiles/."""
import ctypes
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import benchdata
from elliptic_b200 import _native as nat
from elliptic_b200.ec import EC
from elliptic_b200.eddsa import EDDSA

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["secp256k1", "p256", "p384", "ed25519", "curve25519"]
CACHE = "/tmp/eb200_cache"
res = {}
for name in which:
    t = time.time()
    if name in ("secp256k1", "p256", "p384"):
        seed = {"secp256k1": 0xE1110002, "p256": 0xE1110256, "p384": 0xE1110384}[name]
        ds = benchdata.gen_ecdsa_verify(name, n, seed=seed, cache_dir=CACHE)
        ec = EC(name)
        run = lambda: ec.verify_batch_packed(ds["e"], ds["r"], ds["s"], ds["pub"])
    elif name == "ed25519":
        ds = benchdata.gen_ed25519_verify(n, cache_dir=CACHE)
        ed = EDDSA()
        run = lambda: ed.verify_batch_packed(ds["R"], ds["S"], ds["A"], ds["h"])
    else:
        ds = benchdata.gen_x25519_derive(n, cache_dir=CACHE)
        lib = nat.init(0)
        out = np.zeros((n, 32), np.uint8)
        st = np.zeros(n, np.uint8)

        def run():
            nat.check(lib.eb200_x25519_derive_batch(n, ds["priv"].ctypes.data, ds["pubx"].ctypes.data, out.ctypes.data, st.ctypes.data))
            return st
    gen_s = time.time() - t
    best = None
    for it in range(4):
        t = time.time()
        st_ = run()
        wall = time.time() - t
        tm = nat.last_timing()
        ok = bool(np.array_equal(st_, ds["expected"]))
        if best is None or tm["main_kernel_ms"] < best["main_kernel_ms"]:
            best = dict(tm, wall_ms=wall * 1e3, ok=ok)
    best["items_per_s_kernel"] = n / (best["main_kernel_ms"] * 1e-3)
    best["items_per_s_wall_pageable"] = n / (best["wall_ms"] * 1e-3)
    best["gen_s"] = gen_s
    res[name] = best
    print(name, json.dumps(best), flush=True)
json.dump({"n": n, "results": res}, open("gpurun_out/sweep_%d.json" % n, "w"), indent=1)
