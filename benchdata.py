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


def _cpus():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _pmap(fn, items, min_items=64):
    """map over worker processes (fork) -- the generators are pure Python and the GPU boxes have many cores."""
    items = list(items)
    w = min(_cpus(), 32, max(1, len(items) // 8))
    if w <= 1 or len(items) < min_items or os.environ.get("EB200_GEN_SERIAL"):
        return [fn(x) for x in items]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(w) as pool:
        return pool.map(fn, items, chunksize=max(1, len(items) // (4 * w)))


def _smul_job(args):
    return scalar_mul_g(*args)


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


CURVES = {
    # name: (p, n, a, Gx, Gy, byte length)   -- public SECG / NIST parameters
    "secp256k1": (P, N, 0, GX, GY, 32),
    "p256": (2**256 - 2**224 + 2**192 + 2**96 - 1,
             0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551, -3,
             0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
             0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5, 32),
    "p384": (2**384 - 2**128 - 2**96 + 2**32 - 1,
             0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFC7634D81F4372DDF581A0DB248B0A77AECEC196ACCC52973, -3,
             0xAA87CA22BE8B05378EB1C71EF320AD746E1D3B628BA79B9859F741E082542A385502F25DBF55296C3A545E3872760AB7,
             0x3617DE4A96262C6F5D9E98BF9292DC29F8F41DBD289A147CE9DA3113B5F0B8C00A60B1CE1D7E819D7A431D7C90EA0E5F, 48),
    "p521": (2**521 - 1,
             0x1FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFA51868783BF2F966B7FCC0148F709A5D03BB5C9B8899C47AEBB6FB71E91386409, -3,
             0xC6858E06B70404E9CD9E3ECB662395B4429C648139053FB521F828AF606B4D3DBAA14B5E77EFE75928FE1DC127A2FFA8DE3348B3C1856A429BF97E7E31C2E5BD66,
             0x11839296A789A3BC0045C8A5FB42C7D1BD998F54449579B446817AFBD17273E662C97EE72995EF42640C550B9013FAD0761353C7086A272C24088BE94769FD16650, 66),
}


def _jac_dbl(X, Y, Z, P=P, a=0):
    if Y == 0 or Z == 0:
        return 1, 1, 0
    A = X * X % P; B = Y * Y % P; C = B * B % P
    D = 2 * ((X + B) * (X + B) - A - C) % P
    E = (3 * A + a * pow(Z, 4, P)) % P
    X3 = (E * E - 2 * D) % P
    return X3, (E * (D - X3) - 8 * C) % P, 2 * Y * Z % P


def _jac_madd(X1, Y1, Z1, x2, y2, P=P, a=0):
    if Z1 == 0:
        return x2, y2, 1
    Z2 = Z1 * Z1 % P
    U2 = x2 * Z2 % P; S2 = y2 * Z2 * Z1 % P
    H = (U2 - X1) % P; R = (S2 - Y1) % P
    if H == 0:
        return _jac_dbl(X1, Y1, Z1, P, a) if R == 0 else (1, 1, 0)
    H2 = H * H % P; H3 = H2 * H % P; V = X1 * H2 % P
    X3 = (R * R - H3 - 2 * V) % P
    return X3, (R * (V - X3) - Y1 * H3) % P, Z1 * H % P


def scalar_mul_g(k, curve="secp256k1"):
    P, _, a, GX, GY, _ = CURVES[curve]
    X, Y, Z = 1, 1, 0
    for i in range(k.bit_length() - 1, -1, -1):
        X, Y, Z = _jac_dbl(X, Y, Z, P, a)
        if (k >> i) & 1:
            X, Y, Z = _jac_madd(X, Y, Z, GX, GY, P, a)
    zi = pow(Z, -1, P)
    return X * zi * zi % P, Y * zi * zi * zi % P


def gen_secp256k1_verify(n_items, seed=0xE1110002, n_keys=4096, corrupt_every=64, cache_dir=None):
    return gen_ecdsa_verify("secp256k1", n_items, seed, n_keys, corrupt_every, cache_dir)


def _ecdsa_job(args):
    """Items of keys [j0, j1): rows j0*per_key .. min(n_items, j1*per_key) (key-major order)."""
    curve, seed, n_items, n_keys, per_key, corrupt_every, j0, j1 = args
    P, N, _a, GX, GY, LEN = CURVES[curve]
    nk = j1 - j0
    d = [_stream(seed, b"key", j) % (N - 1) + 1 for j in range(j0, j1)]
    k = [_stream(seed, b"nonce", j) % (N - (per_key + 2)) + 1 for j in range(j0, j1)]
    Q = [scalar_mul_g(x, curve) for x in d]
    R = [scalar_mul_g(x, curve) for x in k]
    lo, hi = j0 * per_key, min(n_items, j1 * per_key)
    rows = max(0, hi - lo)
    e_out = np.zeros((rows, LEN), np.uint8)
    r_out = np.zeros((rows, LEN), np.uint8)
    s_out = np.zeros((rows, LEN), np.uint8)
    pub_out = np.zeros((rows, 2 * LEN), np.uint8)
    expected = np.ones(rows, np.uint8)
    pubs = [x.to_bytes(LEN, "big") + y.to_bytes(LEN, "big") for x, y in Q]
    for m in range(per_key):
        kinv = _batch_inv(k, N)
        for jj in range(nk):
            i = (j0 + jj) * per_key + m          # key-major order
            if i >= n_items:
                continue
            e = _stream(seed, b"msg", i)
            if LEN > 64:       # p521: a 512-bit digest (longer values would be shortened by _truncateToN)
                e = (e << 256) | _stream(seed, b"msg2", i)
            elif LEN > 32:
                e = (e << (8 * (LEN - 32))) | (_stream(seed, b"msg2", i) >> (8 * (64 - LEN)))
            r = R[jj][0] % N
            s = kinv[jj] * (e + r * d[jj]) % N
            if r == 0 or s == 0:         # astronomically unlikely; keep the item invalid
                expected[i - lo] = 0
            if corrupt_every and i % corrupt_every == corrupt_every - 1:
                c = _stream(seed, b"corrupt", i)
                which, bit = c % 3, (c >> 8) % (8 * LEN - 1)
                if which == 0:
                    e ^= 1 << bit
                elif which == 1:
                    r ^= 1 << bit
                else:
                    s ^= 1 << bit
                expected[i - lo] = 0
            e_out[i - lo] = np.frombuffer(e.to_bytes(LEN, "big"), np.uint8)
            r_out[i - lo] = np.frombuffer(r.to_bytes(LEN, "big"), np.uint8)
            s_out[i - lo] = np.frombuffer(s.to_bytes(LEN, "big"), np.uint8)
            pub_out[i - lo] = np.frombuffer(pubs[jj], np.uint8)
        if m + 1 < per_key:
            # R_j += G (affine, batched inversion); k_j += 1
            den = [(GX - x) % P for x, _ in R]
            inv = _batch_inv(den, P)
            for jj in range(nk):
                x1, y1 = R[jj]
                lam = (GY - y1) * inv[jj] % P
                x3 = (lam * lam - x1 - GX) % P
                R[jj] = (x3, (lam * (x1 - x3) - y1) % P)
                k[jj] += 1
    return e_out, r_out, s_out, pub_out, expected


def gen_ecdsa_verify(curve, n_items, seed=0xE1110002, n_keys=4096, corrupt_every=64, cache_dir=None):
    """Returns dict of uint8 arrays e,r,s (n,len), pub (n,2*len) and expected (n,)."""
    n_keys = min(n_keys, n_items)
    per_key = (n_items + n_keys - 1) // n_keys
    if cache_dir:
        path = os.path.join(cache_dir, "%s_%x_%d_%d_%d.npz" % (curve, seed, n_items, n_keys, corrupt_every))
        if os.path.exists(path):
            z = np.load(path)
            return {k: z[k] for k in z.files}
    step = max(8, n_keys // 128)
    parts = _pmap(_ecdsa_job, [(curve, seed, n_items, n_keys, per_key, corrupt_every, j0, min(n_keys, j0 + step))
                               for j0 in range(0, n_keys, step)], min_items=2)
    out = dict(zip(("e", "r", "s", "pub", "expected"), (np.concatenate([p[k] for p in parts]) for k in range(5))))
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


# ---------------------------------------------------------------------------------------------
# ed25519 EdDSA verify -- R,S,A (wire format) + h = SHA512(R||A||M) mod n.
N_ED = 0x1000000000000000000000000000000014DEF9DEA2F79CD65812631A5CF5D3ED


def _ed_job(args):
    import nacl.signing
    seed, n_keys, corrupt_every, lo, hi = args
    n = hi - lo
    keys = {}
    R = np.zeros((n, 32), np.uint8); S = np.zeros((n, 32), np.uint8); A = np.zeros((n, 32), np.uint8)
    H = np.zeros((n, 32), np.uint8); M = np.zeros((n, 32), np.uint8)
    expected = np.ones(n, np.uint8)
    for i in range(lo, hi):
        j = i % n_keys
        if j not in keys:
            k = nacl.signing.SigningKey(hashlib.sha256(b"eb200/%08x/edkey/%d" % (seed, j)).digest())
            keys[j] = (k, bytes(k.verify_key))
        key, pub = keys[j]
        msg = hashlib.sha256(b"eb200/%08x/edmsg/%d" % (seed, i)).digest()
        sig = key.sign(msg).signature
        if corrupt_every and i % corrupt_every == corrupt_every - 1:
            msg = msg[:-1] + bytes([(msg[-1] + 1) & 255])
            expected[i - lo] = 0
        h = int.from_bytes(hashlib.sha512(sig[:32] + pub + msg).digest(), "little") % N_ED
        R[i - lo] = np.frombuffer(sig[:32], np.uint8)
        S[i - lo] = np.frombuffer(sig[32:], np.uint8)
        A[i - lo] = np.frombuffer(pub, np.uint8)
        H[i - lo] = np.frombuffer(h.to_bytes(32, "little"), np.uint8)
        M[i - lo] = np.frombuffer(msg, np.uint8)
    return R, S, A, H, M, expected


def gen_ed25519_verify(n_items, seed=0xE1110003, n_keys=4096, corrupt_every=64, cache_dir=None, with_msgs=False):
    """Signatures are produced with libsodium (PyNaCl) -- a generator, not the code under test.
    1/64 items are forged the way test/ed25519-test.js:75-77 does (last message byte + 1)."""
    import nacl.signing
    if cache_dir:
        path = os.path.join(cache_dir, "ed25519m_%x_%d_%d_%d.npz" % (seed, n_items, n_keys, corrupt_every))
        if os.path.exists(path):
            z = np.load(path)
            return {k: z[k] for k in z.files}
    n_keys = min(n_keys, n_items)
    step = max(1024, (n_items + 4 * 32 - 1) // (4 * 32))
    parts = _pmap(_ed_job, [(seed, n_keys, corrupt_every, lo, min(n_items, lo + step)) for lo in range(0, n_items, step)], min_items=2)
    R, S, A, H, M, expected = (np.concatenate([p[k] for p in parts]) for k in range(6))
    out = dict(R=R, S=S, A=A, h=H, msgs=M, expected=expected)
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        np.savez(path, **out)
    return out


# curve25519 ECDH derive -- priv, pubx big-endian; 1/256 twist points.
def gen_x25519_derive(n_items, seed=0xE1110004, n_pubs=4096, twist_every=256, cache_dir=None):
    import nacl.bindings
    P25 = 2**255 - 19
    if cache_dir:
        path = os.path.join(cache_dir, "x25519_%x_%d_%d_%d.npz" % (seed, n_items, n_pubs, twist_every))
        if os.path.exists(path):
            z = np.load(path)
            return {k: z[k] for k in z.files}
    n_pubs = min(n_pubs, n_items)
    pubs = []
    for j in range(n_pubs):
        sk = hashlib.sha256(b"eb200/%08x/xpub/%d" % (seed, j)).digest()
        pubs.append(int.from_bytes(nacl.bindings.crypto_scalarmult_base(sk), "little"))
    twists = []
    t = 2
    while len(twists) < 16:
        rhs = (t * t * t + 486662 * t * t + t) % P25
        if pow(rhs, (P25 - 1) // 2, P25) == P25 - 1:
            twists.append(t)
        t += 1
    priv = np.zeros((n_items, 32), np.uint8)
    pubx = np.zeros((n_items, 32), np.uint8)
    expected = np.ones(n_items, np.uint8)
    for i in range(n_items):
        k = _stream(seed, b"xpriv", i) % (N_ED - 1) + 1
        if twist_every and i % twist_every == twist_every - 1:
            x = twists[(i // twist_every) % len(twists)]
            expected[i] = 5          # the reference throws (bn.js sqrt assertion inside validate)
        else:
            x = pubs[i % n_pubs]
        priv[i] = np.frombuffer(k.to_bytes(32, "big"), np.uint8)
        pubx[i] = np.frombuffer(x.to_bytes(32, "big"), np.uint8)
    out = dict(priv=priv, pubx=pubx, expected=expected)
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        np.savez(path, **out)
    return out
