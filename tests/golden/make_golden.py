#!/usr/bin/env python
"""Extracts the reference's own fixtures/KATs for the hot path into small committed files.

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs (tests/golden/):
    ed25519_sign_input.json.gz   test/fixtures/sign.input: all 1024 lines (secret, pk, msg, sig)
    ed25519_derivation.json      test/fixtures/derivation-fixtures.js (256 entries: secret, a, A, A_P)
    secp256k1_precomputed.json   digest + samples of lib/elliptic/precomputed/secp256k1.js
    ecdsa_kats.json              test/ecdsa-test.js: Maxwell vectors (:352-451), RFC 6979 (:135-350),
                                 Wycheproof p192 (:492-534); test/curve-test.js KATs (:90-112, :298-356)
"""
import gzip
import hashlib
import json
import os
import re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def sign_input():
    lines = open(os.path.join(REF, "test/fixtures/sign.input")).read().split("\n")
    lines = [l for l in lines if l]
    keep = list(range(len(lines)))          # all 1024: EdDSA sign / verify KATs with message lengths 0..1023
    out = []
    for i in keep:
        sk_pk, pk, msg, sig_msg, _ = lines[i].split(":")
        out.append({"i": i, "secret": sk_pk[:64], "pk": pk, "msg": msg, "sig": sig_msg[:128]})
    with gzip.open(os.path.join(OUT, "ed25519_sign_input.json.gz"), "wt") as f:
        json.dump({"source": "test/fixtures/sign.input", "total_lines": len(lines), "vectors": out}, f)
    return len(out)


def derivation():
    src = open(os.path.join(REF, "test/fixtures/derivation-fixtures.js")).read()
    body = src[src.index("["):src.rindex("]") + 1]
    body = body.replace("'", '"')
    body = re.sub(r",(\s*[}\]])", r"\1", body)
    data = json.loads(body)
    out = [{"secret_hex": d["secret_hex"], "A_hex": d["A_hex"], "a_hex": d["a_hex"],
            "x": d["A_P"]["x"], "y": d["A_P"]["y"]} for d in data]
    json.dump({"source": "test/fixtures/derivation-fixtures.js", "vectors": out},
              open(os.path.join(OUT, "ed25519_derivation.json"), "w"))
    return len(out)


def precomputed():
    src = open(os.path.join(REF, "lib/elliptic/precomputed/secp256k1.js")).read()
    hexes = re.findall(r"'([0-9a-f]+)'", src)
    pts = [(int(hexes[i], 16), int(hexes[i + 1], 16)) for i in range(0, len(hexes), 2)]
    assert len(pts) == 65 + 127, len(pts)
    doubles, naf = pts[:65], pts[65:]
    dig = lambda ps: hashlib.sha256(b"".join(x.to_bytes(32, "big") + y.to_bytes(32, "big") for x, y in ps)).hexdigest()
    json.dump({"source": "lib/elliptic/precomputed/secp256k1.js", "doubles_step": 4, "naf_wnd": 7,
               "doubles_sha256": dig(doubles), "naf_sha256": dig(naf),
               "doubles_first": ["%064x" % v for v in doubles[0]], "doubles_last": ["%064x" % v for v in doubles[-1]],
               "naf_first": ["%064x" % v for v in naf[0]], "naf_last": ["%064x" % v for v in naf[-1]]},
              open(os.path.join(OUT, "secp256k1_precomputed.json"), "w"), indent=1)


def js_str(expr):
    """'aa' + 'bb' -> aabb"""
    return "".join(re.findall(r"'([^']*)'", expr))


def kats():
    t = open(os.path.join(REF, "test/ecdsa-test.js")).read()
    out = {"source": "test/ecdsa-test.js, test/curve-test.js"}
    # Maxwell's trick vectors
    sec = t[t.index("describe('Maxwell"):t.index("vectors.forEach")]
    msg = js_str(re.search(r"var msg =\s*([^;]+);", sec).group(1))
    vecs = []
    for m in re.finditer(r"curve: (p\d+),\s*pub: ([^,]+(?:\+[^,]+)*),\s*message: msg,\s*sig: ([^,]+),\s*result: (true|false)", sec):
        vecs.append({"curve": m.group(1), "pub": js_str(m.group(2)), "msg": msg, "sig": js_str(m.group(3)),
                     "result": m.group(4) == "true"})
    assert len(vecs) == 8, len(vecs)
    out["maxwell"] = vecs
    # RFC 6979
    sec = t[t.index("describe('RFC6979 vector'"):t.index("describe('Maxwell")]
    rfc = []
    for blk in re.finditer(r"test\(\{\s*name: '([^']+)',\s*curve: elliptic\.curves\.(\w+),\s*key: ([^,]+),\s*pub: \{\s*x: ([^,]+),\s*y: ([^,]+),\s*\},\s*cases: \[(.*?)\],\s*\}\);", sec, re.S):
        cases = []
        for c in re.finditer(r"message: '(\w+)',\s*hash: hash\.(\w+),\s*r: ([^,]+(?:\+[^,]+)*),\s*s: ([^,]+(?:\+[^,]+)*),", blk.group(6)):
            cases.append({"message": c.group(1), "hash": c.group(2), "r": js_str(c.group(3)), "s": js_str(c.group(4))})
        rfc.append({"name": blk.group(1), "curve": blk.group(2), "key": js_str(blk.group(3)),
                    "x": js_str(blk.group(4)), "y": js_str(blk.group(5)), "cases": cases})
    assert len(rfc) == 5 and all(len(r["cases"]) >= 3 for r in rfc), [(r["curve"], len(r["cases"])) for r in rfc]
    out["rfc6979"] = rfc
    # Wycheproof p192
    sec = t[t.index("Wycheproof special hash case with hex"):]
    out["wycheproof_p192"] = {
        "msg": js_str(re.search(r"var msg =\s*([^;]+);", sec).group(1)),
        "sig": js_str(re.search(r"var sig = ([^;]+);", sec).group(1)),
        "pub": js_str(re.search(r"var pub = ([^;]+);", sec).group(1)), "result": True}
    c = open(os.path.join(REF, "test/curve-test.js")).read()
    out["sec1"] = []
    for name in ("shortPointEvenY", "shortPointOddY"):
        blk = c[c.index("var " + name):]
        blk = blk[:blk.index("};")]
        out["sec1"].append({"x": js_str(re.search(r"x: ([^,]+),", blk).group(1)), "y": js_str(re.search(r"y: ([^,]+),", blk).group(1)),
                            "compact": js_str(re.search(r"compactEncoded:\s*([^,]+(?:\+[^,]+)*),", blk).group(1)),
                            "encoded": js_str(re.search(r"\bencoded:\s*([^,]+(?:\+[^,]+)*),", blk).group(1)),
                            "hybrid": js_str(re.search(r"hybrid:\s*([^,]+(?:\+[^,]+)*),", blk).group(1))})
    out["curve25519_g_mul_6"] = re.search(r"g\.mul\(new BN\('6'\)\).*?\n\s*x: '([0-9a-f]+)'", c, re.S).group(1)
    m = re.search(r"new Uint8Array\(\[([^\]]+)\]\)", c)
    out["ed25519_point_from_y"] = {"y_le_bytes": [int(v) for v in m.group(1).split(",")], "odd": True,
                                   "x": re.search(r"var target = '([0-9a-f]+)';", c).group(1)}
    m = re.search(r"beta\.fromRed\(\)\.toString\(16\),\s*'([0-9a-f]+)'", c)
    out["secp256k1_beta"] = m.group(1) if m else None
    m = re.search(r"lambda\.toString\(16\),\s*'([0-9a-f]+)'", c)
    out["secp256k1_lambda"] = m.group(1) if m else None
    json.dump(out, open(os.path.join(OUT, "ecdsa_kats.json"), "w"), indent=1)
    return len(vecs), sum(len(r["cases"]) for r in rfc)


if __name__ == "__main__":
    print("sign.input vectors:", sign_input())
    print("derivation vectors:", derivation())
    precomputed()
    print("maxwell / rfc6979 cases:", kats())
