"""The C restatement (oracle/c/k256_ref.c, also the CPU baseline) against the Python oracle."""
import random

import numpy as np

from oracle import c_oracle
from oracle.ref_py.ec import EC

P = 2**256 - 2**32 - 977


def _pack(items, k):
    return np.frombuffer(b"".join(it[k].to_bytes(32, "big") for it in items), np.uint8).reshape(-1, 32)


def _run_c(items, threads=1):
    return c_oracle.verify_batch(_pack(items, 0), _pack(items, 1), _pack(items, 2),
                                 np.concatenate([_pack(items, 3), _pack(items, 4)], axis=1), threads)


def _py(ec, it):
    e, r, s, x, y = it
    if not (1 <= r < ec.n and 1 <= s < ec.n):
        return 0
    return int(ec.verify((e if e < ec.n else e - ec.n).to_bytes(32, "big"), {"r": r, "s": s}, {"x": x, "y": y}))


def test_c_restatement_matches_python_oracle_on_edge_cases():
    from test_gpu_k256 import _edge_items
    ec = EC("secp256k1")
    items = _edge_items(ec, random.Random(2024))
    st = _run_c(items, threads=2)
    exp = [_py(ec, it) for it in items]
    assert [int(v) for v in st] == exp


def test_c_restatement_replays_off_curve_keys_exactly():
    """The reference does not validate {x,y} keys (ec/key.js:95); its answer for an off-curve key is
    whatever its GLV/JSF schedule produces.  Both restatements follow that schedule."""
    ec = EC("secp256k1")
    rnd = random.Random(9)
    items = [(rnd.randrange(2**256), rnd.randrange(1, ec.n), rnd.randrange(1, ec.n), rnd.randrange(P), rnd.randrange(P))
             for _ in range(120)]
    assert [int(v) for v in _run_c(items)] == [_py(ec, it) for it in items]


def test_c_restatement_field_mult_count():
    """~2216 field multiplications per verify, the figure BASELINE.md derives its MAC32 count from."""
    import benchdata
    ds = benchdata.gen_secp256k1_verify(512, n_keys=16)
    lib = c_oracle.load()
    lib.k256_ref_fm_reset()
    st = c_oracle.verify_batch(ds["e"], ds["r"], ds["s"], ds["pub"], 1)
    assert np.array_equal(st, ds["expected"])
    per = lib.k256_ref_fm_count() / 512
    assert 2150 < per < 2300, per


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path = the C restatement, Node.js being absent) must run
    without a GPU and print one JSON line with the keys the driver reads."""
    import json
    import os
    import subprocess
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "verifies/s" and line["higher_is_better"] is True
    assert line["metric"] == "secp256k1 ECDSA verifies/sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "verifies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["n_gpus"] == 1 and line["steps"] == 1
