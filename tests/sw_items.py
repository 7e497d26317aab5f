"""Shared edge-case generators for the short-Weierstrass ECDSA parity tests (test helper)."""
import random


def sw_edge_items(ec, ln, seed=5, count=60):
    """(e, r, s, x, y) tuples covering ec/index.js:188-229 for a non-GLV curve."""
    n, G, P = ec.n, ec.g, ec.curve.p
    rnd = random.Random(seed)
    keys = [rnd.randrange(1, n) for _ in range(4)]
    pubs = [G.mul(d) for d in keys]
    items = []
    for t in range(count):
        d, Q = keys[t % 4], pubs[t % 4]
        e = rnd.randrange(2 ** (8 * ln))
        sig = ec.sign(e.to_bytes(ln, "big"), d)
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
    r, s = rnd.randrange(1, n), rnd.randrange(1, n)
    items.append(((-r * d) % n, r, s, Q.x, Q.y))       # R = O
    items.append(((r * d) % n, r, s, Q.x, Q.y))        # u1*G == u2*Q
    items.append((5, r, s, Q.x, (Q.y + 1) % P))        # off-curve -> NEEDS_HOST
    sig = ec.sign((7).to_bytes(ln, "big"), 1); items.append((7, sig.r, sig.s, G.x, G.y))      # Q = G
    mg = G.neg(); sig = ec.sign((8).to_bytes(ln, "big"), n - 1); items.append((8, sig.r, sig.s, mg.x, mg.y))
    sig = ec.sign(b"\x00" * ln, d); items.append((0, sig.r, sig.s, Q.x, Q.y))                 # e = 0
    return items


def sw_expected(ec, ln, it):
    e, r, s, x, y = it
    n = ec.n
    if not (1 <= r < n and 1 <= s < n):
        return 0
    if not ec.curve.validate(ec.curve.point(x, y)):
        return 4
    return int(ec.verify((e if e < n else e - n).to_bytes(ln, "big"), {"r": r, "s": s}, {"x": x, "y": y}))
