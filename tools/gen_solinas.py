#!/usr/bin/env python
"""Generates elliptic_b200/csrc/solinas_gen.inc: the FIPS 186-4 D.2.3 / D.2.4 reductions of a double-width product
modulo p256 / p384, written COLUMN-WISE: every output word is one signed 64-bit sum of input words (the word vectors
s1..s9 / s1..s10 of the standard read down their columns), carries move between columns as arithmetic shifts, and the
word that leaves the top is folded back twice with 2^(32N) = K (mod p), K given by its signed digits.  Against the
vector-wise form (nine / ten N-word carry chains joined at the end) this needs no carry bookkeeping: 110 instead of
150 instructions per p256 reduction on sm_100a.

A small scale factor (1, 3, 4, 8: the constants of the a = -3 doubling) can be applied to the value on the way: the
column sums have the headroom, and the top word grows to at most +-32, still far from wrapping twice.

Range argument (checked by tests/test_hostemu_k256.py on edge operands): value = w + top 2^(32N), w in [0, 2^(32N)),
top in [-4, 4] (p256) / [-1, 3] (p384; exact, from the column polynomials); V2 = w + top K lies in (-p, 2p), so its own top word t2 is -1, 0 or 1; V3 = low(V2)
+ t2 K lies in [0, 2^(32N)) with no carry or borrow (t2 = -1 means low(V2) >= 2^(32N) - 6 K; t2 = +1 means low(V2) <
6 K); V3 < 2p, so one conditional subtraction finishes (sp_final).
"""
import os

Z = None
P256 = dict(
    name="p256", N=8,
    terms=[(1, list(range(8))),
           (2, [Z, Z, Z, 11, 12, 13, 14, 15]),
           (2, [Z, Z, Z, 12, 13, 14, 15, Z]),
           (1, [8, 9, 10, Z, Z, Z, 14, 15]),
           (1, [9, 10, 11, 13, 14, 15, 13, 8]),
           (-1, [11, 12, 13, Z, Z, Z, 8, 10]),
           (-1, [12, 13, 14, 15, Z, Z, 9, 11]),
           (-1, [13, 14, 15, 8, 9, 10, Z, 12]),
           (-1, [14, 15, Z, 9, 10, 11, Z, 13])],
    K={0: 1, 3: -1, 6: -1, 7: 1})              # 2^256 = 2^224 - 2^192 - 2^96 + 1
P384 = dict(
    name="p384", N=12,
    terms=[(1, list(range(12))),
           (2, [Z, Z, Z, Z, 21, 22, 23, Z, Z, Z, Z, Z]),
           (1, list(range(12, 24))),
           (1, [21, 22, 23, 12, 13, 14, 15, 16, 17, 18, 19, 20]),
           (1, [Z, 23, Z, 20, 12, 13, 14, 15, 16, 17, 18, 19]),
           (1, [Z, Z, Z, Z, 20, 21, 22, 23, Z, Z, Z, Z]),
           (1, [20, Z, Z, 21, 22, 23, Z, Z, Z, Z, Z, Z]),
           (-1, [23, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22]),
           (-1, [Z, 20, 21, 22, 23, Z, Z, Z, Z, Z, Z, Z]),
           (-1, [Z, Z, Z, 23, 23, Z, Z, Z, Z, Z, Z, Z])],
    K={0: 1, 1: -1, 3: 1, 4: 1})                # 2^384 = 2^128 + 2^96 - 2^32 + 1


def columns(cfg):
    cols = []
    for j in range(cfg["N"]):
        coef = {}
        for k, vec in cfg["terms"]:
            if vec[j] is not None:
                coef[vec[j]] = coef.get(vec[j], 0) + k
        cols.append({i: k for i, k in coef.items() if k})
    return cols


def expr(coef):
    pos = sorted((i, k) for i, k in coef.items() if k > 0)
    neg = sorted((i, -k) for i, k in coef.items() if k < 0)
    def term(i, k):
        return "(s64)c[%d]" % i if k == 1 else "%d * (s64)c[%d]" % (k, i)
    s = " + ".join(term(i, k) for i, k in pos) or "(s64)0"
    for i, k in neg:
        s += " - " + term(i, k)
    return s


def fold(N, K, src, dst, top, lines):
    """dst[j] = src[j] + top * K[j] with signed carries; returns the name of the value leaving the top"""
    prev = None
    for j in range(N):
        e = "(s64)%s[%d]" % (src, j)
        if j in K:
            e += (" + " if K[j] > 0 else " - ") + top
        if prev:
            e += " + (%s >> 32)" % prev
        lines.append("  const s64 %s%d = %s;" % (dst, j, e))
        prev = "%s%d" % (dst, j)
    return prev


def gen(cfg):
    N, name = cfg["N"], cfg["name"]
    L = ["// v[%d] = the double-width value c[%d] folded to [0, 2^%d), congruent mod %s (one conditional subtraction left)"
         % (N, 2 * N, 32 * N, name),
         "// scale (1, 3, 4 or 8; a compile-time constant at every call site) multiplies the value first: k a b costs one",
         "// 64-bit multiply per column here instead of two or three modular doublings of the reduced product",
         "EB_HD void solinas_%s(u32* v, const u32* c, const int scale = 1) {" % name,
         "  typedef long long s64;"]
    cols = columns(cfg)
    prev = None
    for j, coef in enumerate(cols):
        e = "scale * (%s)" % expr(coef)
        if prev:
            e += " + (%s >> 32)" % prev
        L.append("  const s64 t%d = %s;" % (j, e))
        prev = "t%d" % j
    L.append("  const s64 top = %s >> 32;" % prev)
    L.append("  const u32 w[%d] = {%s};" % (N, ", ".join("(u32)t%d" % j for j in range(N))))
    last = fold(N, cfg["K"], "w", "u", "top", L)
    L.append("  const s64 top2 = %s >> 32;" % last)
    L.append("  const u32 x[%d] = {%s};" % (N, ", ".join("(u32)u%d" % j for j in range(N))))
    # the second fold only happens when low(V2) wrapped: |top| K / 2^(32N) of the residues, i.e. about 2^-29 (p256)
    L.append("  if (top2 != 0) {")
    inner = []
    fold(N, cfg["K"], "x", "y", "top2", inner)
    L += ["  " + l for l in inner]
    for j in range(N):
        L.append("    v[%d] = (u32)y%d;" % (j, j))
    L.append("  } else {")
    for j in range(N):
        L.append("    v[%d] = x[%d];" % (j, j))
    L.append("  }")
    L.append("}")
    return L


def main():
    out = ["// GENERATED by tools/gen_solinas.py -- do not edit."]
    for cfg in (P256, P384):
        out += gen(cfg)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "elliptic_b200", "csrc", "solinas_gen.inc")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    # self-check of the column tables against integers
    import random
    for cfg, p in ((P256, 2**256 - 2**224 + 2**192 + 2**96 - 1), (P384, 2**384 - 2**128 - 2**96 + 2**32 - 1)):
        N = cfg["N"]
        rnd = random.Random(1)
        for it in range(4000):
            scale = (1, 3, 4, 8)[it % 4]
            val = rnd.randrange(p * p) if rnd.random() < 0.8 else (p - 1) ** 2 - rnd.randrange(3)
            c = [(val >> (32 * i)) & 0xffffffff for i in range(2 * N)]
            tot = scale * sum(coef * c[i] << (32 * j) for j, col in enumerate(columns(cfg)) for i, coef in col.items())
            assert tot % p == scale * val % p, cfg["name"]
            top = tot >> (32 * N)
            w = tot & ((1 << 32 * N) - 1)
            Kint = sum(d << (32 * j) for j, d in cfg["K"].items())
            assert (1 << 32 * N) % p == Kint % p
            v2 = w + top * Kint
            t2 = v2 >> (32 * N)
            assert t2 in (-1, 0, 1), (cfg["name"], top, t2)
            v3 = (v2 & ((1 << 32 * N) - 1)) + t2 * Kint
            assert 0 <= v3 < (1 << 32 * N) and v3 < 2 * p and v3 % p == scale * val % p
        # worst cases of the scaled top word: |top| <= 8 * 4 (p256), so |top| K < 2^(32N - 27): the folds cannot wrap twice
        assert 8 * 5 * Kint < (1 << (32 * N - 20))
    print(path)


if __name__ == "__main__":
    main()
