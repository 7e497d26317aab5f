"""Shared EdDSA / X25519 parity-case generators (test helper)."""
import gzip
import json
import os
import random

P = 2**255 - 19
G = os.path.join(os.path.dirname(__file__), "golden")
ORDER8 = ["26e8958fc2b227b045c3f489f2ef98f0d5dfac05d3c63339b13802886d53fc05",
          "c7176a703d4dd84fba3c0b760d10670f2a2053fa2c39ccc64ec7fd7792ac037a"]


def ed_items(limit=None, seed=1):
    """(R, S, A, msg) byte tuples: reference sign.input vectors (valid + forged), S >= n, small-order
    and non-canonical points, undecodable R / A."""
    data = json.load(gzip.open(os.path.join(G, "ed25519_sign_input.json.gz"), "rt"))
    n = 0x1000000000000000000000000000000014DEF9DEA2F79CD65812631A5CF5D3ED
    rnd = random.Random(seed)
    items = []
    # default: the first 128 vectors and every 16th after (message lengths up to 1023); the signing tests use all 1024
    allv = data["vectors"]
    vecs = [allv[i] for i in list(range(128)) + list(range(128, len(allv), 16))] if limit is None else allv[:limit]
    for v in vecs:
        sig, pk, msg = bytes.fromhex(v["sig"]), bytes.fromhex(v["pk"]), bytes.fromhex(v["msg"])
        items.append((sig[:32], sig[32:], pk, msg))
        f = bytearray(msg) if msg else bytearray(b"x")
        f[-1] = (f[-1] + 1) & 255
        items.append((sig[:32], sig[32:], pk, bytes(f)))
    v = data["vectors"][3]
    sig, pk, msg = bytes.fromhex(v["sig"]), bytes.fromhex(v["pk"]), bytes.fromhex(v["msg"])
    items.append((sig[:32], (int.from_bytes(sig[32:], "little") + n).to_bytes(32, "little"), pk, msg))
    items.append((sig[:32], n.to_bytes(32, "little"), pk, msg))
    small = [(1).to_bytes(32, "little"), (P - 1).to_bytes(32, "little"), (0).to_bytes(32, "little"),
             bytes.fromhex(ORDER8[0]), bytes.fromhex(ORDER8[1]), (P + 1).to_bytes(32, "little"), P.to_bytes(32, "little"),
             bytes([1] + [0] * 30 + [0x80]), bytes([0xEE] + [0xFF] * 30 + [0x7F])]
    for sp in small:
        items.append((sp, sig[32:], pk, msg))
        items.append((sig[:32], sig[32:], sp, msg))
        items.append((sp, (0).to_bytes(32, "little"), sp, msg))
    for _ in range(30):
        y = rnd.randrange(2**256).to_bytes(32, "little")
        items.append((y, sig[32:], pk, msg))
        items.append((sig[:32], sig[32:], y, msg))
    return items


def ed_expected(ed, it):
    from oracle.ref_py.bn import RefError
    r, s, a, m = it
    try:
        return int(ed.verify(m, r + s, a))
    except RefError as ex:
        return {"invalid point": 2, "Assertion failed": 5}[ex.args[0]]


def x_items(n_order, seed=2, count=60):
    rnd = random.Random(seed)
    its = [(rnd.randrange(1, n_order), rnd.randrange(2**256) if t % 3 else rnd.randrange(P)) for t in range(count)]
    its += [(6, 9), (0, 9), (1, 9), (n_order - 1, 9), (5, 0), (5, 1), (5, P - 1), (5, P), (5, P + 9), (5, 2**256 - 1),
            (7, int.from_bytes(bytes.fromhex(ORDER8[0]), "little") & (2**255 - 1)), (7, 39382357235489614581723060781553021112529911719440698176882885853963445705823)]
    return its


def x_expected(ec, c, k, x):
    from oracle.ref_py.bn import RefError
    from oracle.ref_py.ec import KeyPair
    try:
        return 1, KeyPair(ec, priv=k).derive(c.point(x, 1))
    except RefError as ex:
        return {"Assertion failed": 5, "public point not validated": 3}[ex.args[0]], 0
