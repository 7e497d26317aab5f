"""Synthetic workload generator for bench.py and the large-batch tests.

BASELINE.md config 2: secp256k1 ECDSA verify, N signatures, SoA big-endian
e,r,s (32 B each) + pub x||y (64 B): `n_keys` distinct keys x N/n_keys messages,
SHA-256 counter stream seeded with 0xE1110002, one item in `corrupt_every`
corrupted by a single bit flip in e, r or s.  Self-contained integer code (no
oracle, no product code): signatures are produced by walking the nonce
(k -> k+1, R -> R+G) with batched modular inversions, so 2^20 signatures take
seconds in pure Python.  Expected statuses: TRUE unless corrupted.
"""
import hashlib
import os

import numpy as np

P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8


def _stream(seed, tag, i):
    return int.from_bytes(hashlib.sha256(b"eb200/%08x/%s/%d" % (seed, tag, i)).digest(), "big")


def _batch_inv(vals, m):
    """Montgomery's trick; all vals non-zero mod m."""
    n = len(vals)
    pref = [1] * (n + 1)
    for i, v in enumerate(vals):
        pref[i + 1] = pref[i] * v % m
    inv = pow(pref[n], -1, m)
    out = [0] * n
    for i in range(n - 1, -1, -1):
        out[i] = inv * pref[i] % m
        inv = inv * vals[i] % m
    return out


def _jac_dbl(X, Y, Z):
    if Y == 0 or Z == 0:
        return 1, 1, 0
    A = X * X % P; B = Y * Y % P; C = B * B % P
    D = 2 * ((X + B) * (X + B) - A - C) % P
    E = 3 * A % P
    X3 = (E * E - 2 * D) % P
    return X3, (E * (D - X3) - 8 * C) % P, 2 * Y * Z % P


def _jac_madd(X1, Y1, Z1, x2, y2):
    if Z1 == 0:
        return x2, y2, 1
    Z2 = Z1 * Z1 % P
    U2 = x2 * Z2 % P; S2 = y2 * Z2 * Z1 % P
    H = (U2 - X1) % P; R = (S2 - Y1) % P
    if H == 0:
        return _jac_dbl(X1, Y1, Z1) if R == 0 else (1, 1, 0)
    H2 = H * H % P; H3 = H2 * H % P; V = X1 * H2 % P
    X3 = (R * R - H3 - 2 * V) % P
    return X3, (R * (V - X3) - Y1 * H3) % P, Z1 * H % P


def scalar_mul_g(k):
    X, Y, Z = 1, 1, 0
    for i in range(k.bit_length() - 1, -1, -1):
        X, Y, Z = _jac_dbl(X, Y, Z)
        if (k >> i) & 1:
            X, Y, Z = _jac_madd(X, Y, Z, GX, GY)
    zi = pow(Z, -1, P)
    return X * zi * zi % P, Y * zi * zi * zi % P


def gen_secp256k1_verify(n_items, seed=0xE1110002, n_keys=4096, corrupt_every=64, cache_dir=None):
    """Returns dict of uint8 arrays e,r,s (n,32), pub (n,64) and expected (n,)."""
    n_keys = min(n_keys, n_items)
    per_key = (n_items + n_keys - 1) // n_keys
    if cache_dir:
        path = os.path.join(cache_dir, "k256_%x_%d_%d_%d.npz" % (seed, n_items, n_keys, corrupt_every))
        if os.path.exists(path):
            z = np.load(path)
            return {k: z[k] for k in z.files}
    d = [_stream(seed, b"key", j) % (N - 1) + 1 for j in range(n_keys)]
    k = [_stream(seed, b"nonce", j) % (N - (per_key + 2)) + 1 for j in range(n_keys)]
    Q = [scalar_mul_g(x) for x in d]
    R = [scalar_mul_g(x) for x in k]
    e_out = np.zeros((n_items, 32), np.uint8)
    r_out = np.zeros((n_items, 32), np.uint8)
    s_out = np.zeros((n_items, 32), np.uint8)
    pub_out = np.zeros((n_items, 64), np.uint8)
    expected = np.ones(n_items, np.uint8)
    pubs = [x.to_bytes(32, "big") + y.to_bytes(32, "big") for x, y in Q]
    for m in range(per_key):
        kinv = _batch_inv(k, N)
        for j in range(n_keys):
            i = j * per_key + m          # key-major order
            if i >= n_items:
                continue
            e = _stream(seed, b"msg", i)
            r = R[j][0] % N
            s = kinv[j] * (e + r * d[j]) % N
            if r == 0 or s == 0:         # astronomically unlikely; keep the item invalid
                expected[i] = 0
            if corrupt_every and i % corrupt_every == corrupt_every - 1:
                c = _stream(seed, b"corrupt", i)
                which, bit = c % 3, (c >> 8) % 256
                if which == 0:
                    e ^= 1 << bit
                elif which == 1:
                    r ^= 1 << bit
                else:
                    s ^= 1 << bit
                expected[i] = 0
            e_out[i] = np.frombuffer(e.to_bytes(32, "big"), np.uint8)
            r_out[i] = np.frombuffer(r.to_bytes(32, "big"), np.uint8)
            s_out[i] = np.frombuffer(s.to_bytes(32, "big"), np.uint8)
            pub_out[i] = np.frombuffer(pubs[j], np.uint8)
        if m + 1 < per_key:
            # R_j += G (affine, batched inversion); k_j += 1
            den = [(GX - x) % P for x, _ in R]
            inv = _batch_inv(den, P)
            for j in range(n_keys):
                x1, y1 = R[j]
                lam = (GY - y1) * inv[j] % P
                x3 = (lam * lam - x1 - GX) % P
                R[j] = (x3, (lam * (x1 - x3) - y1) % P)
                k[j] += 1
    out = dict(e=e_out, r=r_out, s=s_out, pub=pub_out, expected=expected)
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        np.savez(path, **out)
    return out


if __name__ == "__main__":
    import sys
    import time
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 14
    t = time.time()
    dset = gen_secp256k1_verify(n)
    print("generated", n, "in %.1fs" % (time.time() - t), "valid", int(dset["expected"].sum()))
