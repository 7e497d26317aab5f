"""EC.sign with the `k` / `pers` options and EC.genKeyPair({entropy, pers}) on the GPU, through the reference-shaped
host API, against the oracle (lib/elliptic/ec/index.js:55-79, 143-157; the reference exercises them at
test/ecdsa-test.js:72-87, 453-465)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CURVES = ["secp256k1", "p256", "p384", "p521", "p192", "p224"]


@pytest.mark.parametrize("name", CURVES)
def test_sign_with_pers_matches_the_oracle(native, name):
    from elliptic_b200.ec import EC as GpuEC
    from oracle.ref_py.ec import EC
    ec, gec = EC(name), GpuEC(name)
    rnd = random.Random(41)
    ln = (ec.curve.p.bit_length() + 7) // 8
    cnt = 96 if ln < 66 else 24
    privs = [rnd.randrange(1, ec.n) for _ in range(cnt)]
    msgs = [rnd.randrange(min(ec.n, 1 << (8 * ln - 8))) for _ in range(cnt)]
    for pers, enc, raw in (("my.pers", None, b"my.pers"), ("0a0b0c", "hex", bytes([10, 11, 12])), (list(range(70)), None, bytes(range(70)))):
        for canon in (False, True):
            r, s, rec = gec.sign_batch(msgs, privs, canonical=canon, pers=pers, pers_enc=enc)
            for i in range(cnt):
                sig = ec.sign(msgs[i], privs[i], canonical=canon, pers=raw)
                assert (r[i], s[i], int(rec[i])) == (sig.r, sig.s, sig.recovery_param), (name, i)
    # and the signatures verify
    pubs = [ec.g.mul(d) for d in privs]
    st = gec.verify_batch(msgs, [{"r": a, "s": b} for a, b in zip(r, s)], [{"x": q.x, "y": q.y} for q in pubs])
    assert (st == 1).all()


@pytest.mark.parametrize("name", CURVES)
def test_sign_with_caller_nonces(native, name):
    """options.k(iter): the first nonce of some items is rejected by the reference's loop (k <= 1, k >= n - 1), so the
    host mirror asks k for the next one exactly as the reference does."""
    from elliptic_b200.ec import EC as GpuEC
    from oracle.ref_py.ec import EC
    ec, gec = EC(name), GpuEC(name)
    rnd = random.Random(43)
    ln = (ec.curve.p.bit_length() + 7) // 8
    cnt = 64 if ln < 66 else 16
    privs = [rnd.randrange(1, ec.n) for _ in range(cnt)]
    msgs = [rnd.randrange(min(ec.n, 1 << (8 * ln - 8))) for _ in range(cnt)]
    first = [rnd.randrange(2, ec.n - 1) for _ in range(cnt)]
    first[3], first[7], first[11] = 0, 1, (ec.n - 1 if ln != 66 else 1)
    second = [rnd.randrange(2, ec.n - 1) for _ in range(cnt)]
    asked = []

    def k(i, it):
        asked.append((i, it))
        return first[i] if it == 0 else second[i]
    r, s, rec = gec.sign_batch(msgs, privs, k=k)
    assert sorted(x for x in asked if x[1] == 1) == [(3, 1), (7, 1), (11, 1)]
    for i in range(cnt):
        sig = ec.sign(msgs[i], privs[i], k_fn=lambda it, i=i: first[i] if it == 0 else second[i])
        assert (r[i], s[i], int(rec[i])) == (sig.r, sig.s, sig.recovery_param), (name, i)
    # single-item form, hex nonce
    one = gec.sign(msgs[0], privs[0], k=lambda it: "%x" % second[0])
    sig = ec.sign(msgs[0], privs[0], k_fn=lambda it: second[0])
    assert (one["r"], one["s"], one["recoveryParam"]) == (sig.r, sig.s, sig.recovery_param)


@pytest.mark.parametrize("name", CURVES)
def test_gen_key_pair_from_entropy(native, name):
    from elliptic_b200.ec import EC as GpuEC, EllipticError
    from oracle.ref_py.ec import EC
    ec, gec = EC(name), GpuEC(name)
    rnd = random.Random(47)
    cnt = 64 if name != "p521" else 12
    for ne, pers in ((24, None), (32, None), (40, "kp")):
        ents = [bytes(rnd.randrange(256) for _ in range(ne)) for _ in range(cnt)]
        privs, pubs = gec.gen_key_pair_batch(ents, pers=pers)
        for i in range(cnt):
            kp = ec.gen_key_pair(ents[i], (pers or "").encode())
            q = ec.g.mul(kp.priv)
            assert privs[i] == kp.priv and pubs[i] == (q.x, q.y), (name, ne, i)
    # entropy given as a string: utf8 by default, hex on request (genKeyPair's entropyEnc)
    a, _ = gec.gen_key_pair_batch(["0123456789abcdef0123456789abcdef0123456789abcdef"], entropy_enc="hex")
    assert a[0] == ec.gen_key_pair(bytes.fromhex("0123456789abcdef0123456789abcdef0123456789abcdef")).priv
    b, _ = gec.gen_key_pair_batch(["0123456789abcdef0123456789abcdef"])
    assert b[0] == ec.gen_key_pair(b"0123456789abcdef0123456789abcdef").priv
    with pytest.raises(EllipticError, match="Not enough entropy"):
        gec.gen_key_pair_batch([b"short"])
