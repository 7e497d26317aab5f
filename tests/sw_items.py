"""Shared edge-case generators for the short-Weierstrass ECDSA parity tests (test helper)."""
import random


def sw_edge_items(ec, ln, seed=5, count=60, ebits=None):
    """(e, r, s, x, y) tuples covering ec/index.js:188-229 for a non-GLV curve."""
    n, G, P = ec.n, ec.g, ec.curve.p
    rnd = random.Random(seed)
    keys = [rnd.randrange(1, n) for _ in range(4)]
    pubs = [G.mul(d) for d in keys]
    items = []
    for t in range(count):
        d, Q = keys[t % 4], pubs[t % 4]
        e = rnd.randrange(2 ** (ebits or 8 * ln))
        sig = ec.sign(e, d)          # BN form: _truncateToN measures the value, not a padded array
        r, s = sig.r, sig.s
        k = t % 10
        if k == 1: e ^= 1 << rnd.randrange(8 * ln)
        if k == 2: r ^= 1 << rnd.randrange(8 * ln - 1)
        if k == 3: s = n - s
        if k == 4: r = 0
        if k == 5: s = n
        if k == 6: Q = pubs[(t + 1) % 4]
        if k == 7: s ^= 1 << rnd.randrange(8 * ln - 1)
        items.append((e, r, s, Q.x, Q.y))
    d, Q = keys[0], pubs[0]
    lim = 2 ** (ebits or 8 * ln)          # messages must not be shortened by _truncateToN (p521: below 2^520)
    for sign in (-1, 1):                  # e = -r d: R = O;  e = r d: u1*G == u2*Q
        while True:
            r, s = rnd.randrange(1, n), rnd.randrange(1, n)
            if (sign * r * d) % n < lim:
                break
        items.append(((sign * r * d) % n, r, s, Q.x, Q.y))
    items.append((5, r, s, Q.x, (Q.y + 1) % P))        # off-curve -> NEEDS_HOST
    sig = ec.sign(7, 1); items.append((7, sig.r, sig.s, G.x, G.y))      # Q = G
    mg = G.neg(); sig = ec.sign(8, n - 1); items.append((8, sig.r, sig.s, mg.x, mg.y))
    sig = ec.sign(0, d); items.append((0, sig.r, sig.s, Q.x, Q.y))                 # e = 0
    return items


def sw_off_curve_items(ec, ln, seed=4, count=12, ebits=None):
    """Un-validated off-curve keys (ec/key.js:95).  Even items are minted so that the reference's own
    schedule lands on x(R) == r (verdict TRUE); odd items are random (FALSE)."""
    n, P = ec.n, ec.curve.p
    rnd = random.Random(seed)
    items = []
    while len(items) < count:
        x, y = rnd.randrange(P), rnd.randrange(P)
        if len(items) % 4 == 2: y = 0
        Q = ec.curve.point(x, y)
        if ec.curve.validate(Q):
            continue
        lim = 2 ** (ebits or 8 * ln)
        if len(items) % 2:
            items.append((rnd.randrange(min(n, lim)), rnd.randrange(1, n), rnd.randrange(1, n), x, y))
            continue
        u1, u2 = rnd.randrange(1, n), rnd.randrange(1, n)
        R = ec.g.jmul_add(u1, Q, u2)
        if R.z % P == 0:
            continue
        zi = pow(R.z, -1, P)
        r = (R.x * zi * zi % P) % n
        if r == 0:
            continue
        s = r * pow(u2, -1, n) % n
        e = u1 * s % n
        if e >= lim:
            continue
        # u1, u2 recomputed by verify are the same residues, so the schedule and hence R repeat
        items.append((e, r, s, x, y))
    return items


def sw_expected(ec, ln, it, replay=True):
    """replay=False: the fast kernel alone flags off-curve keys 4; the product path (replay=True)
    re-runs them through the exact-replay kernel and returns the reference's verdict."""
    e, r, s, x, y = it
    n = ec.n
    if not (1 <= r < n and 1 <= s < n):
        return 0
    if not replay and not ec.curve.validate(ec.curve.point(x, y)):
        return 4
    return int(ec.verify(e if e < n else e - n, {"r": r, "s": s}, {"x": x, "y": y}))
