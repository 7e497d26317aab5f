"""GPU parity tests: p256 / p384 ECDSA verify (generic Montgomery path) and SEC1 key decoding."""
import ctypes
import hashlib
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
KATS = json.load(open(os.path.join(G, "ecdsa_kats.json")))
CURVES = {"p256": (2, 32), "p384": (3, 48), "p521": (6, 66), "p192": (7, 24), "p224": (8, 28)}
NL = {"p256": 8, "p384": 12, "p521": 18, "p192": 6, "p224": 8}          # 32-bit limbs per field element (p521: 17 + one spare)


def limbs(vals, n):
    a = np.zeros((len(vals), n), np.uint32)
    for i, v in enumerate(vals):
        for k in range(n):
            a[i, k] = (v >> (32 * k)) & 0xFFFFFFFF
    return a


def ints(a):
    return [sum(int(a[i, k]) << (32 * k) for k in range(a.shape[1])) for i in range(a.shape[0])]


@pytest.mark.parametrize("name", ["p256", "p384", "p521", "p192", "p224"])
def test_montgomery_field_bit_exact(native, name):
    from elliptic_b200 import _native as nat
    from oracle.ref_py import curves
    cid, ln = CURVES[name]
    p = curves.get(name).curve.p
    nl = NL[name]
    rnd = random.Random(3)
    edge = [0, 1, 2, p - 1, p - 2, (1 << (8 * ln)) - 1, p, p + 1, 1 << (8 * ln - 1), (1 << 32) - 1, 1 << 32]
    a = edge + [rnd.randrange(1 << (8 * ln)) for _ in range(1500)]
    b = [a[(7 * i + 3) % len(a)] for i in range(len(a))]
    a += [x for x in edge for _ in edge]
    b += [y for _ in edge for y in edge]
    A, B = limbs(a, nl), limbs(b, nl)
    for op, fn in ((0, lambda x, y: x * y), (1, lambda x, y: x * x), (2, lambda x, y: x + y), (3, lambda x, y: x - y),
                   (4, lambda x, y: -x)):
        out = np.zeros_like(A)
        nat.check(native.eb200_selftest_fe(cid, op, len(a), A.ctypes.data, B.ctypes.data, out.ctypes.data))
        for x, y, g in zip(a, b, ints(out)):
            assert g == fn(x, y) % p, (name, op, hex(x), hex(y))
    out = np.zeros_like(A[:48])
    nat.check(native.eb200_selftest_fe(cid, 7, 48, A.ctypes.data, B.ctypes.data, out.ctypes.data))
    for x, g in zip(a[:48], ints(out)):
        assert g == pow(x % p, p - 2, p)


@pytest.mark.parametrize("name", ["p256", "p384", "p521", "p192", "p224"])
def test_fixed_base_table(native, name):
    from elliptic_b200 import _native as nat
    from oracle.ref_py.ec import EC
    cid, ln = CURVES[name]
    ec = EC(name)
    p, n, nl = ec.curve.p, ec.n, NL[name]
    W, E, B = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    nat.check(native.eb200_selftest_gtab_dims(cid, ctypes.byref(W), ctypes.byref(E), ctypes.byref(B)))
    W, E, B = W.value, E.value, B.value
    tab = np.zeros(W * E * 2 * nl, np.uint32)
    nat.check(native.eb200_selftest_gtab(cid, tab.ctypes.data, tab.size))
    tab = tab.reshape(W, E, 2, nl)
    # p256 / p384 / p521 hold plain residues (fp_special.cuh); p192 / p224 are in Montgomery form (fp_mont.cuh)
    R = 1 if name in ("p256", "p384", "p521") else 1 << (32 * nl)
    rnd = random.Random(6)
    for j, i in [(0, 0), (0, E - 1), (W - 1, 0), (W - 1, E - 1)] + [(rnd.randrange(W), rnd.randrange(E)) for _ in range(30)]:
        pt = ec.g.mul(((2 * i + 1) << (B * j)) % n)
        x, y = ints(tab[j, i])
        assert (x, y) == (pt.x * R % p, pt.y * R % p), (name, j, i)


@pytest.mark.parametrize("name", ["p256", "p384", "p521", "p192", "p224"])
def test_verify_parity(native, name):
    from elliptic_b200.ec import EC as GpuEC
    from oracle.ref_py.ec import EC
    from sw_items import sw_edge_items, sw_expected
    cid, ln = CURVES[name]
    ec = EC(name)
    items = sw_edge_items(ec, ln, seed=11, count=200 if ln < 66 else 40, ebits=520 if ln == 66 else None)
    pack = lambda k: np.frombuffer(b"".join(it[k].to_bytes(ln, "big") for it in items), np.uint8).reshape(-1, ln)
    st = GpuEC(name).verify_batch_packed(pack(0), pack(1), pack(2), np.concatenate([pack(3), pack(4)], axis=1))
    exp = [sw_expected(ec, ln, it) for it in items]
    bad = [(i, int(st[i]), exp[i]) for i in range(len(items)) if int(st[i]) != exp[i]]
    assert not bad, bad[:10]
    assert {0, 1} <= set(exp)


@pytest.mark.parametrize("name", ["p256", "p384", "p521", "p192", "p224"])
def test_off_curve_keys_get_the_reference_answer(native, name):
    """Un-validated off-curve keys (ec/key.js:95) are re-run by the exact-replay kernel
    (ecdsa_sw_replay.cuh); no item may come back as NEEDS_HOST."""
    from elliptic_b200.ec import EC as GpuEC
    from oracle.ref_py.ec import EC
    from sw_items import sw_off_curve_items
    cid, ln = CURVES[name]
    ec = EC(name)
    items = sw_off_curve_items(ec, ln, seed=8, count=40 if ln < 66 else 12, ebits=520 if ln == 66 else None)
    pack = lambda k: np.frombuffer(b"".join(it[k].to_bytes(ln, "big") for it in items), np.uint8).reshape(-1, ln)
    st = GpuEC(name).verify_batch_packed(pack(0), pack(1), pack(2), np.concatenate([pack(3), pack(4)], axis=1))
    exp = [int(ec.verify(it[0], {"r": it[1], "s": it[2]}, {"x": it[3], "y": it[4]})) for it in items]
    assert [int(v) for v in st] == exp
    assert exp.count(1) == len(items) // 2


def test_reference_maxwell_vectors_on_gpu(native):
    """test/ecdsa-test.js:352-451 through the single-item API (DER sig, SEC1 key, hex)."""
    from elliptic_b200.ec import EC as GpuEC
    for v in KATS["maxwell"]:
        assert GpuEC(v["curve"]).verify(v["msg"], v["sig"], v["pub"], "hex") is v["result"], v


def test_reference_rfc6979_vectors_verify_on_gpu(native):
    """test/ecdsa-test.js:135-350 (p256, p384, p521): the published (r, s) must verify."""
    from elliptic_b200.ec import EC as GpuEC
    hs = {"sha1": hashlib.sha1, "sha224": hashlib.sha224, "sha256": hashlib.sha256, "sha384": hashlib.sha384, "sha512": hashlib.sha512}
    seen = 0
    for blk in KATS["rfc6979"]:
        if blk["curve"] not in ("p256", "p384", "p521", "p192", "p224"):
            continue
        gec = GpuEC(blk["curve"])
        for c in blk["cases"]:
            dg = hs[c["hash"]](c["message"].encode()).digest()
            assert gec.verify(dg, {"r": c["r"], "s": c["s"]}, {"x": blk["x"], "y": blk["y"]}) is True
            bad = bytearray(dg); bad[0] ^= 1
            assert gec.verify(bytes(bad), {"r": c["r"], "s": c["s"]}, {"x": blk["x"], "y": blk["y"]}) is False
            seen += 1
    assert seen >= 17


@pytest.mark.parametrize("name", ["secp256k1", "p256", "p384", "p192", "p224"])
def test_sec1_key_formats_on_gpu(native, name):
    """Compressed / hybrid / uncompressed keys decoded on the GPU (base.js:270-292, short.js:187-204),
    including x with no square root (-> the reference throws 'invalid point')."""
    from elliptic_b200 import _native as nat
    from elliptic_b200.ec import EC as GpuEC
    from oracle.ref_py.ec import EC
    from oracle.ref_py.bn import RefError
    ec, gec = EC(name), GpuEC(name)
    ln = (ec.curve.p.bit_length() + 7) // 8
    rnd = random.Random(8)
    e, r, s, comp, unc, exp_c, exp_u = [], [], [], [], [], [], []
    for t in range(96):
        d = rnd.randrange(1, ec.n)
        Q = ec.g.mul(d)
        m = rnd.randrange(1 << (8 * ln - 1))
        sig = ec.sign(m.to_bytes(ln, "big"), d)
        c = bytearray(Q.encode(True))
        u = bytearray(Q.encode())
        kind = t % 8
        if kind == 1: c[0] ^= 1                     # wrong parity -> valid point, other key -> false
        if kind == 2: c[0] = 5; u[0] = 5            # unknown prefix
        if kind == 3: u[0] = 6 + (Q.y & 1)          # correct hybrid
        if kind == 4: u[0] = 7 - (Q.y & 1)          # hybrid parity assertion fails
        if kind == 5:                                # x without a square root
            x = Q.x
            while True:
                x = (x + 1) % ec.curve.p
                try:
                    ec.curve.point_from_x(x, False)
                except RefError:
                    break
            c[1:] = x.to_bytes(ln, "big")
        for buf, exp in ((c, exp_c), (u, exp_u)):
            try:
                exp.append(int(ec.verify(m.to_bytes(ln, "big"), sig, bytes(buf))))
            except RefError as ex:
                exp.append({"invalid point": 2, "Assertion failed": 5, "Unknown point format": 6}[ex.args[0]])
        e.append(m.to_bytes(ln, "big")); r.append(sig.r.to_bytes(ln, "big")); s.append(sig.s.to_bytes(ln, "big"))
        comp.append(bytes(c)); unc.append(bytes(u))
    arr = lambda lst: np.frombuffer(b"".join(lst), np.uint8).reshape(len(lst), -1)
    st_c = gec.verify_batch_packed(arr(e), arr(r), arr(s), arr(comp), nat.PUB_SEC1_33)
    st_u = gec.verify_batch_packed(arr(e), arr(r), arr(s), arr(unc), nat.PUB_SEC1_65)
    assert [int(v) for v in st_c] == exp_c
    assert [int(v) for v in st_u] == exp_u
    assert {1, 0, 6} <= set(exp_c) and (2 in exp_c or 5 in exp_c) and {1, 5, 6} <= set(exp_u)    # p224: Tonelli-Shanks asserts


@pytest.mark.parametrize("name", ["p256", "p384", "p521", "p192", "p224"])
def test_mul_and_mul_add_batches(native, name):
    """curve.point(x, y).mul(k) (_wnafMul), G.mul(k), G.mulAdd(k1, P, k2) (short.js:422-441) on the non-GLV
    curves: hostemu edge cases (oversize / zero scalars, P = +-G, off-curve points) plus random items."""
    import os, sys
    sys.path.insert(0, os.path.dirname(__file__))
    from test_hostemu_k256 import mul_cases
    from elliptic_b200.ec import EC as GpuEC
    from oracle.ref_py.ec import EC
    cid, ln = CURVES[name]
    ec = EC(name)
    rnd = random.Random(13)
    cases = mul_cases(ec, seed=7, bits=8 * ln)
    base = [ec.g.mul(rnd.randrange(1, ec.n)) for _ in range(4)]
    for t in range(60):
        P = base[t % 4]
        cases.append((rnd.randrange(2 ** (8 * ln)), rnd.randrange(2 ** (8 * ln)), P.x, P.y))
    ref = lambda pt: None if pt.is_infinity() else (pt.get_x(), pt.get_y())
    g = GpuEC(name)
    pts = [(c[2], c[3]) for c in cases]
    assert g.mul_add_batch([c[0] for c in cases], pts, [c[1] for c in cases]) == \
        [ref(ec.g.mul_add(c[0], ec.curve.point(c[2], c[3]), c[1])) for c in cases]
    assert g.mul_batch(pts, [c[1] for c in cases]) == [ref(ec.curve.point(c[2], c[3]).mul(c[1])) for c in cases]
    assert g.g_mul_batch([c[1] for c in cases]) == [ref(ec.g.mul(c[1])) for c in cases]


@pytest.mark.parametrize("name", ["secp256k1", "p256", "p384", "p521", "p192", "p224"])
def test_ecdh_derive_on_short_curves(native, name):
    """KeyPair.derive (ec/key.js:102-107; test/ecdh-test.js:8-43): shared x equals the oracle's, both
    sides agree, and the twist-attack point {x: 14, y: 16} is refused with the reference's message."""
    from elliptic_b200.ec import EC as GpuEC, EllipticError
    from oracle.ref_py.ec import EC
    ec, g = EC(name), GpuEC(name)
    rnd = random.Random(17)
    privs = [rnd.randrange(1, ec.n) for _ in range(40)] + [1, ec.n - 1, ec.n + 5]
    peers = [ec.g.mul(rnd.randrange(1, ec.n)) for _ in range(len(privs))]
    pubs = [{"x": q.x, "y": q.y} for q in peers]
    pubs[3] = {"x": 14, "y": 16}
    pubs[5] = {"x": peers[5].x, "y": (peers[5].y + 1) % ec.curve.p}
    vals, st = g.derive_batch(privs, pubs)
    for i, (d, pb) in enumerate(zip(privs, pubs)):
        pt = ec.curve.point(pb["x"], pb["y"])
        if not ec.curve.validate(pt):
            assert (int(st[i]), vals[i]) == (3, None)
        else:
            assert (int(st[i]), vals[i]) == (1, ec.key_from_private(d % ec.n).derive(pt)), i
    a, b = privs[0], privs[1]
    A, B = g.g_mul_batch([a, b])
    assert g.derive(a, {"x": B[0], "y": B[1]}) == g.derive(b, {"x": A[0], "y": A[1]})
    with pytest.raises(EllipticError, match="public point not validated"):
        g.derive(a, {"x": 14, "y": 16})


@pytest.mark.parametrize("name", ["p256", "p384"])
def test_sign_batch_matches_reference_rfc6979(native, name):
    """EC.sign on the NIST curves (ec/index.js:110-186): r, s, recoveryParam equal the oracle's (HMAC-DRBG over
    the curve's default hash generated on the GPU), the RFC 6979 A.2.5 / A.2.6 vectors whose hash is the curve's
    default come out exactly (test/ecdsa-test.js:135-350), and what was signed verifies."""
    from elliptic_b200.ec import EC as GpuEC
    from oracle.ref_py.ec import EC
    cid, ln = CURVES[name]
    ec, gec = EC(name), GpuEC(name)
    rnd = random.Random(31)
    msgs = [rnd.randbytes(ln) for _ in range(64)] + [b"\x00" * ln, b"\xff" * ln, rnd.randbytes(20), rnd.randbytes(ln + 9)]
    privs = [rnd.randrange(1, ec.n) for _ in msgs]
    privs[1], privs[2] = 1, ec.n - 1
    for canonical in (False, True):
        r, s, rec = gec.sign_batch(msgs, privs, canonical=canonical)
        for i, (m, d) in enumerate(zip(msgs, privs)):
            sig = ec.sign(m, d, canonical=canonical)
            assert (r[i], s[i], int(rec[i])) == (sig.r, sig.s, sig.recovery_param), (i, canonical)
    pubs = [ec.g.mul(d) for d in privs]
    st = gec.verify_batch(msgs, [{"r": a, "s": b} for a, b in zip(r, s)], [{"x": q.x, "y": q.y} for q in pubs])
    assert bool((st == 1).all())
    hs = {"sha256": hashlib.sha256, "sha384": hashlib.sha384}
    default = {"p256": "sha256", "p384": "sha384"}[name]
    seen = 0
    for blk in KATS["rfc6979"]:
        if blk["curve"] != name:
            continue
        for c in blk["cases"]:
            if c["hash"] != default:
                continue
            dg = hs[c["hash"]](c["message"].encode()).digest()
            got = gec.sign(dg, int(blk["key"], 16))
            assert (got["r"], got["s"]) == (int(c["r"], 16), int(c["s"], 16)), c
            seen += 1
    assert seen >= 1


@pytest.mark.parametrize("name", ["p256", "p384", "p521", "p192", "p224"])
def test_recover_pub_key_parity(native, name):
    """EC.recoverPubKey on the NIST curves (ec/index.js:231-259): all four recovery params of real signatures,
    the second-key throw, x without a square root, r = 0 (point at infinity); the signer's key comes back for
    the signature's own recovery param."""
    from elliptic_b200.ec import EC as GpuEC
    from oracle.ref_py.ec import EC
    from rec_items import rec_items, rec_expected
    ec = EC(name)
    items, truth = rec_items(ec, count=40 if name != "p521" else 10)
    gec = GpuEC(name)
    pts, st = gec.recover_pub_key_batch([it[0] for it in items], [{"r": it[1] or "00", "s": it[2] or "00"} for it in items],
                                        [it[3] for it in items])
    for i, it in enumerate(items):
        assert (int(st[i]), pts[i]) == rec_expected(ec, it), (i, it[3])
        if i in truth:
            assert pts[i] == truth[i]


@pytest.mark.parametrize("name", ["p192", "p224", "p521"])
def test_sign_batch_on_the_remaining_presets(native, name):
    """EC.sign where key / message / digest lengths differ (p192, p224: SHA-256 with 24 / 28-byte keys; p521: SHA-512
    with 66-byte keys and the value-dependent _truncateToN(k, true)): r, s, recoveryParam equal the oracle's, the
    RFC 6979 vectors for the curve's default hash come out exactly, and what was signed verifies."""
    from elliptic_b200.ec import EC as GpuEC
    from oracle.ref_py.ec import EC
    cid, ln = CURVES[name]
    ec, gec = EC(name), GpuEC(name)
    rnd = random.Random(41)
    msgs = [rnd.randbytes(ln) for _ in range(24)] + [b"\x00" * ln, b"\xff" * ln, rnd.randbytes(20), rnd.randbytes(ln + 9)]
    privs = [rnd.randrange(1, ec.n) for _ in msgs]
    privs[1], privs[2] = 1, ec.n - 1
    for canonical in (False, True):
        r, s, rec = gec.sign_batch(msgs, privs, canonical=canonical)
        for i, (m, d) in enumerate(zip(msgs, privs)):
            sig = ec.sign(m, d, canonical=canonical)
            assert (r[i], s[i], int(rec[i])) == (sig.r, sig.s, sig.recovery_param), (i, canonical)
    pubs = [ec.g.mul(d) for d in privs]
    st = gec.verify_batch(msgs, [{"r": a, "s": b} for a, b in zip(r, s)], [{"x": q.x, "y": q.y} for q in pubs])
    assert bool((st == 1).all())
    hs = {"sha256": hashlib.sha256, "sha512": hashlib.sha512}
    default = {"p192": "sha256", "p224": "sha256", "p521": "sha512"}[name]
    seen = 0
    for blk in KATS["rfc6979"]:
        if blk["curve"] != name:
            continue
        for c in blk["cases"]:
            if c["hash"] != default:
                continue
            got = gec.sign(hs[default](c["message"].encode()).digest(), int(blk["key"], 16))
            assert (got["r"], got["s"]) == (int(c["r"], 16), int(c["s"], 16)), c
            seen += 1
    assert seen >= 1
