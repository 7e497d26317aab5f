"""Every BASELINE.json configuration at its STATED size through the host-buffer C ABI: all statuses (and, for
ECDH, all outputs' oracle spot checks) against the generator's expectation, >= 512 items against the Python
oracle, plus the size-independent properties the domain offers (ECDH agreement, verify(sign(.)) round trip at
2^20).  Generators run on all host cores (benchdata._pmap) and cache under /tmp."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CACHE = os.environ.get("EB200_CACHE", "/tmp/eb200_cache")


def _bench():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import bench          # the spot-check workers must be importable by name in the pool's processes
    return bench


@pytest.mark.parametrize("name,log2n,seed,keys", [("secp256k1", 20, 0xE1110002, 4096), ("p256", 20, 0xE1110256, 1024),
                                                   ("p384", 20, 0xE1110384, 1024), ("p521", 18, 0xE1110521, 1024)])
def test_ecdsa_verify_at_config_size(native, name, log2n, seed, keys):
    import benchdata
    from elliptic_b200.ec import EC
    b = _bench()
    n = 1 << log2n
    ds = benchdata.gen_ecdsa_verify(name, n, seed=seed, n_keys=keys, cache_dir=CACHE)
    st = EC(name).verify_batch_packed(ds["e"], ds["r"], ds["s"], ds["pub"])
    assert np.array_equal(st, ds["expected"])
    assert int(st.sum()) == n - n // 64
    idx = b.spot_indices(n)
    got = b.pmap(b._spot_ecdsa, [(ds["e"][i].tobytes(), ds["r"][i].tobytes(), ds["s"][i].tobytes(), ds["pub"][i].tobytes()) for i in idx],
                 wrap=lambda c: (name, c))
    assert len(idx) >= 512 and [int(st[i]) for i in idx] == got


def test_ed25519_verify_at_config_size(native):
    import benchdata
    from elliptic_b200.eddsa import EDDSA
    b = _bench()
    n = 1 << 20
    ds = benchdata.gen_ed25519_verify(n, cache_dir=CACHE, with_msgs=True)
    ed = EDDSA()
    st = ed.verify_batch_packed(ds["R"], ds["S"], ds["A"], ds["h"])                      # h supplied by the host
    assert np.array_equal(st, ds["expected"])
    off = np.arange(n + 1, dtype=np.uint64) * 32
    st2 = ed.verify_batch_msgs_packed(ds["R"], ds["S"], ds["A"], ds["msgs"].reshape(-1), off)   # SHA-512 on the GPU
    assert np.array_equal(st2, ds["expected"])
    idx = b.spot_indices(n)
    got = b.pmap(b._spot_ed, [(ds["R"][i].tobytes(), ds["S"][i].tobytes(), ds["A"][i].tobytes(), ds["msgs"][i].tobytes()) for i in idx])
    assert len(idx) >= 512 and [int(st[i]) for i in idx] == got


def test_curve25519_derive_at_config_size(native):
    import benchdata
    from elliptic_b200.ec import EC
    b = _bench()
    n = 1 << 20
    ds = benchdata.gen_x25519_derive(n, cache_dir=CACHE)
    ec = EC("curve25519")
    out, st = ec.derive_batch_packed(ds["priv"], ds["pubx"])
    assert np.array_equal(st, ds["expected"])
    idx = b.spot_indices(n, corrupt_every=256)
    got = b.pmap(b._spot_x, [(ds["priv"][i].tobytes(), ds["pubx"][i].tobytes()) for i in idx])
    assert len(idx) >= 512
    for i, (stv, x) in zip(idx, got):
        assert int(st[i]) == stv
        if stv == 1:
            assert out[i].tobytes() == x.to_bytes(32, "big")
    # ECDH agreement over the whole batch: a * (b * 9) == b * (a * 9) with a = priv[i], b = priv[i ^ 1]
    nine = np.zeros((n, 32), np.uint8); nine[:, 31] = 9
    pa, sa = ec.derive_batch_packed(ds["priv"], nine)                 # a_i * 9
    assert (sa == 1).all()
    swapped = ds["priv"].reshape(-1, 2, 32)[:, ::-1].reshape(n, 32).copy()
    left, s1 = ec.derive_batch_packed(swapped, pa)                    # b_i * (a_i * 9)
    right = left.reshape(-1, 2, 32)[:, ::-1].reshape(n, 32)           # item i^1 computed a_i * (b_i * 9)
    assert (s1 == 1).all() and np.array_equal(left, right)


def test_secp256k1_sign_verify_round_trip_2e20(native):
    from elliptic_b200 import _native as nat
    from elliptic_b200.ec import EC
    lib = nat.init(0)
    n = 1 << 20
    rng = np.random.default_rng(20)
    e = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); e[:, 0] &= 0x7F
    priv = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); priv[:, 0] &= 0x7F; priv[:, 31] |= 1
    r = np.zeros((n, 32), np.uint8); s = np.zeros((n, 32), np.uint8); rec = np.zeros(n, np.uint8); st = np.zeros(n, np.uint8)
    pub = np.zeros((n, 64), np.uint8)
    nat.check(lib.eb200_scalar_mul_batch(1, n, priv.ctypes.data, None, pub.ctypes.data, st.ctypes.data))
    assert (st == 1).all()
    nat.check(lib.eb200_ecdsa_sign_batch(1, n, e.ctypes.data, priv.ctypes.data, 0, r.ctypes.data, s.ctypes.data, rec.ctypes.data, st.ctypes.data))
    assert (st == 1).all()
    assert (EC("secp256k1").verify_batch_packed(e, r, s, pub) == 1).all()
    e[:, 31] ^= 1
    assert (EC("secp256k1").verify_batch_packed(e, r, s, pub) == 0).all()
