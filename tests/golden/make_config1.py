#!/usr/bin/env python
"""Mints tests/golden/secp256k1_verify_1024.json.gz -- BASELINE.json configs[0] / SURVEY 8d "config 1":
1024 fixed (msgHash, sig, pub) triples for `ec.verify` on secp256k1 in the reference's own argument forms
(hex message hash; signature as DER hex or {r, s}; key as SEC1 hex -- uncompressed 04, hybrid 06/07,
compressed 02/03 -- or {x, y}).  The reference holds no golden triple for secp256k1 verify, so
the file is minted from the oracle (oracle/ref_py, a restatement of lib/elliptic/ec/index.js:188-229 pinned
to the reference's fixtures by tests/test_oracle_golden.py) and every item a stock ECDSA library can
express is cross-checked against OpenSSL (`cryptography`).

  d_i = SHA256("eb200/key" || i) mod (n - 1) + 1,  e_i = SHA256("eb200/msg" || i),  RFC 6979 nonces.
  0..767     valid, uncompressed key, {r, s}
  768..895   valid in the other encodings: DER signature / compressed key / hybrid key / {x, y} key / high s
  896..1023  16 kinds of invalid or special items, 8 of each (see KINDS)

Run in the build container:  python tests/golden/make_config1.py
"""
import gzip
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "secp256k1_verify_1024.json.gz")

from oracle.ref_py.bn import RefError            # noqa: E402
from oracle.ref_py.ec import EC                  # noqa: E402

ec = EC("secp256k1")
N, P, G = ec.n, ec.curve.p, ec.g
KINDS = ["flip_e", "flip_r", "flip_s", "wrong_key", "r_zero", "s_zero", "r_eq_n", "s_ge_n", "r_plus_n_candidate",
         "q_is_g", "result_infinity", "off_curve_key", "x_ge_p", "bad_hybrid_parity", "bad_der", "no_sqrt_compressed"]


def h(tag, i):
    return int.from_bytes(hashlib.sha256(tag + b"%d" % i).digest(), "big")


def der(r, s):
    def integer(v):
        b = v.to_bytes((v.bit_length() + 7) // 8 or 1, "big")
        if b[0] & 0x80:
            b = b"\x00" + b
        return b"\x02" + bytes([len(b)]) + b
    body = integer(r) + integer(s)
    return (b"\x30" + bytes([len(body)]) + body).hex()


def sec1(q, mode):
    x, y = q.get_x(), q.get_y()
    if mode == "compressed":
        return ("03" if y & 1 else "02") + "%064x" % x
    tag = {"uncompressed": "04", "hybrid": "07" if y & 1 else "06"}[mode]
    return tag + "%064x%064x" % (x, y)


def forge(q, i):
    """(e, r, s) that verifies under a key of unknown discrete log: R = u1 G + u2 Q."""
    u1, u2 = h(b"eb200/u1", i) % (N - 1) + 1, h(b"eb200/u2", i) % (N - 1) + 1
    R = G.mul(u1).add(q.mul(u2))
    r = R.get_x() % N
    s = r * pow(u2, -1, N) % N
    return u1 * s % N, r, s


def main():
    items = []
    for i in range(1024):
        d = h(b"eb200/key", i) % (N - 1) + 1
        e = h(b"eb200/msg", i)
        q = G.mul(d)
        sig = ec.sign(e.to_bytes(32, "big"), d)
        r, s = sig.r, sig.s
        msg, sg, key, kind = "%064x" % e, {"r": "%064x" % r, "s": "%064x" % s}, sec1(q, "uncompressed"), "valid"
        if 768 <= i < 896:
            v = i % 5
            kind = ["valid_der", "valid_compressed", "valid_hybrid", "valid_xy", "valid_high_s"][v]
            if v == 0:
                sg = der(r, s)
            elif v == 1:
                key = sec1(q, "compressed")
            elif v == 2:
                key = sec1(q, "hybrid")
            elif v == 3:
                key = {"x": "%064x" % q.get_x(), "y": "%064x" % q.get_y()}
            else:
                hs = s if s > N // 2 else N - s          # no low-s rule on verify
                sg = der(r, hs) if i % 2 else {"r": "%064x" % r, "s": "%064x" % hs}
        elif i >= 896:
            kind = KINDS[(i - 896) % 16]
            bit = h(b"eb200/bit", i) % 255
            if kind == "flip_e":
                msg = "%064x" % (e ^ (1 << bit))
            elif kind == "flip_r":
                sg = {"r": "%064x" % (r ^ (1 << bit)), "s": "%064x" % s}
            elif kind == "flip_s":
                sg = der(r, s ^ (1 << bit))
            elif kind == "wrong_key":
                key = sec1(G.mul(d + 1), "compressed")
            elif kind == "r_zero":
                sg = {"r": "00", "s": "%064x" % s}
            elif kind == "s_zero":
                sg = {"r": "%064x" % r, "s": "0"}
            elif kind == "r_eq_n":
                sg = {"r": "%064x" % N, "s": "%064x" % s}
            elif kind == "s_ge_n":
                sg = {"r": "%064x" % r, "s": "%064x" % (s + N)}
            elif kind == "r_plus_n_candidate":
                # R with n <= x(R) < p: r = x(R) - n, accepted only through the second eqXToP candidate (short.js:908-925)
                x = N + h(b"eb200/x", i) % (P - N)
                while True:
                    try:
                        R = ec.curve.point_from_x(x, i & 1)
                        break
                    except RefError:
                        x = N + (x + 1 - N) % (P - N)
                s2 = h(b"eb200/s", i) % (N - 1) + 1
                r2 = R.get_x() - N
                u1, u2 = e * pow(s2, -1, N) % N, r2 * pow(s2, -1, N) % N
                q2 = R.add(G.mul(u1).neg()).mul(pow(u2, -1, N))          # Q = (R - u1 G) / u2
                sg = {"r": "%064x" % r2, "s": "%064x" % s2}
                key = sec1(q2, "uncompressed" if i % 2 else "compressed")
            elif kind == "q_is_g":
                sig1 = ec.sign(e.to_bytes(32, "big"), 1)
                sg, key = der(sig1.r, sig1.s), sec1(G, "uncompressed")
            elif kind == "result_infinity":
                # Q = -(u1 / u2) G  =>  u1 G + u2 Q = O  =>  false (ec/index.js:222-223)
                u1, u2 = e * pow(s, -1, N) % N, r * pow(s, -1, N) % N
                key = sec1(G.mul((N - u1) * pow(u2, -1, N) % N), "uncompressed")
            elif kind == "off_curve_key":
                # not validated by the reference (ec/key.js:95): the verdict is whatever its own schedule computes
                key = {"x": "%064x" % q.get_x(), "y": "%064x" % ((q.get_y() + 1 + (i >> 4)) % P)}
            elif kind == "x_ge_p":
                # coordinates >= p are reduced on entry (short.js:258-268): a key with a small x, given as x + p
                x = h(b"eb200/sx", i) % (2**32 + 900)
                while True:
                    try:
                        q3 = ec.curve.point_from_x(x, i & 1)
                        break
                    except RefError:
                        x = (x + 1) % (2**32 + 900)
                e3, r3, s3 = forge(q3, i)
                msg, sg = "%064x" % e3, {"r": "%064x" % r3, "s": "%064x" % s3}
                key = {"x": "%064x" % (q3.get_x() + P), "y": "%064x" % q3.get_y()}
            elif kind == "bad_hybrid_parity":
                key = ("06" if q.get_y() & 1 else "07") + "%064x%064x" % (q.get_x(), q.get_y())
            elif kind == "bad_der":
                good = bytes.fromhex(der(r, s))
                sg = (good[:1] + bytes([good[1] + 1]) + good[2:]).hex()      # wrong outer length
            elif kind == "no_sqrt_compressed":
                x = h(b"eb200/nx", i) % P
                while pow((x * x * x + 7) % P, (P - 1) // 2, P) == 1:
                    x = (x + 1) % P
                key = "02" + "%064x" % x
        try:
            exp = bool(ec.verify(msg, sg, key, "hex"))
        except RefError as ex:
            exp = "throw:" + ex.args[0]
        items.append({"i": i, "kind": kind, "msg": msg, "sig": sg, "pub": key, "expected": exp, "openssl": openssl(msg, sg, key)})
    kinds = {}
    for it in items:
        kinds.setdefault(it["kind"], []).append(it["expected"])
    summary = {k: {v: [str(x) for x in vals].count(v) for v in sorted(set(map(str, vals)))} for k, vals in kinds.items()}
    json.dump({"source": "minted by tests/golden/make_config1.py from oracle/ref_py (SURVEY 8d config 1); "
                         "reference call: new EC('secp256k1').verify(msg, sig, pub, 'hex')",
               "summary": summary, "items": items}, gzip.open(OUT, "wt"), separators=(",", ":"))
    print(json.dumps(summary, indent=1))
    agree = sum(1 for it in items if it["openssl"] is not None and it["openssl"] == it["expected"])
    print("openssl cross-checked:", sum(it["openssl"] is not None for it in items), "agree:", agree)
    assert all(it["openssl"] is None or it["openssl"] == it["expected"] for it in items)


def openssl(msg, sg, key):
    """OpenSSL's verdict where a stock ECDSA library can express the item (on-curve key in SEC1 or x/y form with
    coordinates < p, r and s in [1, n-1]); None otherwise."""
    from cryptography.exceptions import InvalidSignature
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import ec as cec, utils as cutils
    try:
        if isinstance(key, dict):
            x, y = int(key["x"], 16), int(key["y"], 16)
            if x >= P or y >= P:
                return None
            pk = cec.EllipticCurvePublicNumbers(x, y, cec.SECP256K1()).public_key()
        else:
            b = bytes.fromhex(key)
            if b[0] in (6, 7):
                if (b[0] == 6) != (b[-1] % 2 == 0):
                    return None
                b = b"\x04" + b[1:]
            pk = cec.EllipticCurvePublicKey.from_encoded_point(cec.SECP256K1(), b)
    except ValueError:
        return None
    if isinstance(sg, dict):
        r, s = int(sg["r"], 16), int(sg["s"], 16)
        if not (1 <= r < N and 1 <= s < N):
            return None                     # OpenSSL rejects these at the encoding layer; the reference returns false
        sigb = cutils.encode_dss_signature(r, s)
    else:
        sigb = bytes.fromhex(sg)
        try:
            cutils.decode_dss_signature(sigb)
        except ValueError:
            return None
    try:
        pk.verify(sigb, bytes.fromhex(msg), cec.ECDSA(cutils.Prehashed(hashes.SHA256())))
        return True
    except InvalidSignature:
        return False


if __name__ == "__main__":
    main()
