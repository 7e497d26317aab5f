"""recoverPubKey parity cases (test helper)."""
import random

P = 2**256 - 2**32 - 977


def rec_items(ec, seed=12, count=60):
    n = ec.n
    P = ec.curve.p
    mbits = min(n.bit_length() - 1, 512)        # messages _truncateToN leaves alone
    rnd = random.Random(seed)
    items, truth = [], {}
    for t in range(count):
        d = rnd.randrange(1, n)
        m = rnd.randrange(2**mbits)
        sig = ec.sign(m, d)
        Q = ec.g.mul(d)
        for j in range(4):
            if j == sig.recovery_param:
                truth[len(items)] = (Q.x, Q.y)
            items.append((m, sig.r, sig.s, j))
    items += [(5, 0, 7, 0), (5, 1, 7, 1), (0, 12345, 999, 0), (n + 3, 12345, 999, 1), (5, n - 1, 3, 2),
              (5, P - n + 5, 3, 2), (5, 3, 0, 0), (5, n, 3, 0), (5, 2, 3, 3), (7, 1, 1, 2)]
    return items, truth


def rec_expected(ec, it):
    from oracle.ref_py.bn import RefError
    from oracle.ref_py.signature import Signature
    m, r, s, j = it
    sg = Signature.__new__(Signature)
    sg.r, sg.s, sg.recovery_param = r, s, None
    try:
        Q = ec.recover_pub_key(m, sg, j)
        return (7, None) if Q.is_infinity() else (1, (Q.x, Q.y))
    except RefError as ex:
        return {"invalid point": 2, "Unable to find sencond key candinate": 8, "Assertion failed": 5}[ex.args[0]], None
