"""GPU parity tests: ed25519 EdDSA verify and curve25519 ECDH derive vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_ed25519_verify_parity(native):
    from elliptic_b200.eddsa import EDDSA as GpuEd
    from oracle.ref_py.eddsa import EDDSA
    from ed_items import ed_items, ed_expected
    ed, ged = EDDSA(), GpuEd()
    items = ed_items()
    arr = lambda k: np.frombuffer(b"".join(it[k] for it in items), np.uint8).reshape(-1, 32)
    h = np.frombuffer(b"".join(ed.hash_int(it[0], it[2], it[3]).to_bytes(32, "little") for it in items), np.uint8).reshape(-1, 32)
    st = ged.verify_batch_packed(arr(0), arr(1), arr(2), h)
    exp = [ed_expected(ed, it) for it in items]
    bad = [(i, int(st[i]), exp[i]) for i in range(len(items)) if int(st[i]) != exp[i]]
    assert not bad, bad[:10]
    assert {0, 1, 2, 5} <= set(exp)


def test_ed25519_reference_argument_forms(native):
    """test/ed25519-test.js:44-85 through the single-item API: hex strings and byte arrays."""
    import gzip, json, os
    from elliptic_b200.eddsa import EDDSA as GpuEd
    from elliptic_b200.ec import EllipticError
    data = json.load(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "ed25519_sign_input.json.gz"), "rt"))
    ged = GpuEd()
    for v in data["vectors"][:24]:
        assert ged.verify(v["msg"], v["sig"], v["pk"]) is True
        assert ged.verify(list(bytes.fromhex(v["msg"])), list(bytes.fromhex(v["sig"])), list(bytes.fromhex(v["pk"]))) is True
        assert ged.verify(v["msg"] + "00", v["sig"], v["pk"]) is False
    v = data["vectors"][5]
    with pytest.raises(EllipticError, match="Signature has invalid size"):
        ged.verify(v["msg"], v["sig"][:-2], v["pk"])
    with pytest.raises(EllipticError, match="Assertion failed|invalid point"):
        ged.verify(v["msg"], v["sig"], "02" + "00" * 31)


def test_eddsa_sign_all_1024_sign_input_kats(native):
    """EDDSA.sign on the GPU (eddsa/index.js:34-44): byte-identical signatures and public keys for all 1024 lines
    of the reference's test/fixtures/sign.input (test/ed25519-test.js:44-85), then every one of them verifies and
    every forged message does not."""
    import gzip, json, os
    from elliptic_b200.eddsa import EDDSA as GpuEd
    data = json.load(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "ed25519_sign_input.json.gz"), "rt"))
    vecs = data["vectors"]
    assert len(vecs) == 1024
    ged = GpuEd()
    sigs = ged.sign_batch([v["msg"] for v in vecs], [v["secret"] for v in vecs])
    assert [s.hex() for s in sigs] == [v["sig"] for v in vecs]
    sec = np.frombuffer(b"".join(bytes.fromhex(v["secret"]) for v in vecs), np.uint8).reshape(-1, 32)
    pubs = ged.public_from_secret_batch(sec)
    assert [pubs[i].tobytes().hex() for i in range(1024)] == [v["pk"] for v in vecs]
    msgs = [bytes.fromhex(v["msg"]) for v in vecs]
    assert (ged.verify_batch(msgs, sigs, [bytes.fromhex(v["pk"]) for v in vecs]) == 1).all()
    forged = [(m[:-1] + bytes([(m[-1] + 1) & 255])) if m else b"x" for m in msgs]
    assert (ged.verify_batch(forged, sigs, [bytes.fromhex(v["pk"]) for v in vecs]) == 0).all()
    # single-item reference forms
    assert ged.sign(vecs[7]["msg"], vecs[7]["secret"]).hex() == vecs[7]["sig"]
    assert ged.sign(list(bytes.fromhex(vecs[9]["msg"])), list(bytes.fromhex(vecs[9]["secret"]))).hex() == vecs[9]["sig"]


def test_eddsa_sign_verify_round_trip_2e18(native):
    """2^18 random secrets and 32-byte messages: sign on the GPU, verify on the GPU, libsodium agrees on a sample."""
    import nacl.signing
    from elliptic_b200.eddsa import EDDSA as GpuEd
    n = 1 << 18
    rng = np.random.default_rng(33)
    sec = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64) * 32
    ged = GpuEd()
    sig, pub = ged.sign_batch_packed(sec, msgs.reshape(-1), off, want_pub=True)
    st = ged.verify_batch_msgs_packed(sig[:, :32].copy(), sig[:, 32:].copy(), pub, msgs.reshape(-1), off)
    assert (st == 1).all()
    msgs[:, 5] ^= 1
    st = ged.verify_batch_msgs_packed(sig[:, :32].copy(), sig[:, 32:].copy(), pub, msgs.reshape(-1), off)
    assert (st == 0).all()
    msgs[:, 5] ^= 1
    for i in range(0, n, n // 64):
        k = nacl.signing.SigningKey(sec[i].tobytes())
        assert bytes(k.verify_key) == pub[i].tobytes()
        assert k.sign(msgs[i].tobytes()).signature == sig[i].tobytes()


def test_curve25519_derive_parity(native):
    from elliptic_b200.ec import EC as GpuEC
    from elliptic_b200 import _native as nat
    from oracle.ref_py.ec import EC
    from oracle.ref_py import curves
    from ed_items import x_items, x_expected
    ec, c = EC("curve25519"), curves.get("curve25519").curve
    its = x_items(ec.n, count=300)
    vals, st = GpuEC("curve25519").derive_batch([k for k, _ in its], [x for _, x in its])
    for (k, x), v, s in zip(its, vals, st):
        es, ev = x_expected(ec, c, k, x)
        assert int(s) == es and (v or 0) == ev, (k, x)
    # reference KAT (test/curve-test.js:348-356) and ECDH agreement (test/ecdh-test.js:8-29)
    g = GpuEC("curve25519")
    assert "%x" % g.derive(6, 9) == "26954ccdc99ebf34f8f1dde5e6bb080685fec73640494c28f9fe0bfa8c794531"
    a, b = 0x1111111111111111111111, 0x2222222222222222222222222
    pa, pb = g.derive(a, 9), g.derive(b, 9)
    assert g.derive(a, pb) == g.derive(b, pa)


def test_ed25519_gpu_hashing_equals_host_hashing(native):
    """hashInt on the GPU (SHA-512 + reduction mod n) vs hashlib, over the reference's sign.input messages
    (lengths 0..1023 bytes) -- statuses must be identical and match the oracle."""
    from elliptic_b200.eddsa import EDDSA as GpuEd
    from oracle.ref_py.eddsa import EDDSA
    from ed_items import ed_items, ed_expected
    ed, ged = EDDSA(), GpuEd()
    items = ed_items()
    msgs = [it[3] for it in items]
    sigs = [it[0] + it[1] for it in items]
    pubs = [it[2] for it in items]
    a = ged.verify_batch(msgs, sigs, pubs, gpu_hash=True)
    b = ged.verify_batch(msgs, sigs, pubs, gpu_hash=False)
    assert np.array_equal(a, b)
    assert [int(v) for v in a] == [ed_expected(ed, it) for it in items]


def test_f25519_field_bit_exact(native):
    """PTX field ops mod 2^255 - 19 (weak representatives allowed, value mod p exact), edge values included."""
    import random
    from elliptic_b200 import _native as nat
    P = 2**255 - 19
    rnd = random.Random(21)
    edge = [0, 1, 2, P - 1, P, P + 1, 2**256 - 1, 2**256 - 2, 2**255, 2**255 - 1, 19, 38, 2**256 - 38, 2**256 - 19,
            2**256 - 2**32, (1 << 256) - 39, 2 * P, 2 * P + 37]
    a = edge + [rnd.randrange(2**256) for _ in range(3000)]
    b = [a[(7 * i + 3) % len(a)] for i in range(len(a))]
    a += [x for x in edge for _ in edge]
    b += [y for _ in edge for y in edge]

    def limbs(vals):
        m = np.zeros((len(vals), 8), np.uint32)
        for i, v in enumerate(vals):
            for k in range(8):
                m[i, k] = (v >> (32 * k)) & 0xFFFFFFFF
        return m

    def run(op, x, y):
        X, Y = limbs(x), limbs(y)
        out = np.zeros_like(X)
        nat.check(native.eb200_selftest_fe(nat.CURVE_ED25519, op, len(x), X.ctypes.data, Y.ctypes.data, out.ctypes.data))
        return [sum(int(out[i, k]) << (32 * k) for k in range(8)) for i in range(len(x))]
    for op, fn in ((0, lambda x, y: x * y), (1, lambda x, y: x * x), (2, lambda x, y: x + y), (3, lambda x, y: x - y), (4, lambda x, y: -x)):
        for x, y, g in zip(a, b, run(op, a, b)):
            assert g < 2**256 and g % P == fn(x, y) % P, (op, hex(x), hex(y))
    ks = [121666 if i % 2 else 486662 for i in range(len(a))]
    for x, k, g in zip(a, ks, run(5, a, ks)):
        assert g % P == x * k % P
    for x, g in zip(a, run(6, a, b)):
        assert g == x % P
    for x, g in zip(a[:48], run(7, a[:48], b[:48])):
        assert g % P == pow(x % P, P - 2, P)
    for x, g in zip(a[:48], run(8, a[:48], b[:48])):
        assert g % P == pow(x % P, (P - 5) // 8, P)
