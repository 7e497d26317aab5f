"""The generic `.curve` batch API on the GPU: short Weierstrass curves with run-time parameters
(`new elliptic.curve.short({p, a, b})`, lib/elliptic/curve/short.js:10-24) -- Point.add / dbl / mul / mulAdd /
validate batches against the oracle's ShortCurve, starting with the toy curve of the reference's own test
(test/curve-test.js:9-22), then the NIST / SECG parameter sets given as plain numbers, cross-checked with the
presets' tuned kernels."""
import random

import pytest

pytestmark = pytest.mark.gpu


def _aff(p):
    return None if p.is_infinity() else (p.get_x(), p.get_y())


def test_reference_example_curve(native):
    from elliptic_b200.curve import ShortCurve
    from elliptic_b200.ec import NeedsReferencePath
    from oracle.ref_py.short import ShortCurve as RefCurve
    ref = RefCurve({"p": 0x1d, "a": 4, "b": 0x14})
    cv = ShortCurve("1d", "4", "14")
    p = ("18", "16")                                            # curve.point('18', '16')
    P = ref.point(0x18, 0x16)
    assert cv.validate_batch([p]) == [True]
    d = cv.dbl_batch([p])
    assert d == [_aff(P.dbl())] and cv.validate_batch(d) == [True]
    assert cv.add_batch(d, [p]) == [_aff(P.dbl().add(P))]
    dd = cv.add_batch(d, d)
    assert dd == cv.add_batch(cv.add_batch(cv.add_batch([p], [p]), [p]), [p]) == [_aff(P.add(P).add(P).add(P))]
    # every pair of points of the curve, and every multiple
    pts = [(x, y) for x in range(29) for y in range(29) if (y * y - x ** 3 - 4 * x - 20) % 29 == 0]
    A = [a for a in pts for _ in pts]
    B = [b for _ in pts for b in pts]
    got = cv.add_batch(A, B)
    assert got == [_aff(ref.point(*a).add(ref.point(*b))) for a, b in zip(A, B)] and None in got
    ks = list(range(40))
    for q in pts[:6]:
        assert cv.mul_batch([q] * len(ks), ks) == [_aff(ref.point(*q).mul(k)) if k else None for k in ks]
    assert cv.validate_batch([(3, 3), pts[0]]) == [False, True]
    with pytest.raises(NeedsReferencePath):
        cv.mul_batch([(3, 3)], [5])


@pytest.mark.parametrize("name", ["secp256k1", "p256", "p384", "p521", "p192", "p224"])
def test_named_parameter_sets_through_the_runtime_path(native, name):
    from elliptic_b200.curve import ShortCurve
    from elliptic_b200.ec import EC as GpuEC
    from oracle.ref_py.ec import EC
    ec = EC(name)
    c = ec.curve
    cv = ShortCurve(c.p, c.a, c.b)
    rnd = random.Random(17)
    n = 48 if name != "p521" else 16
    ds = [rnd.randrange(1, ec.n) for _ in range(n)]
    pts = GpuEC(name).g_mul_batch(ds)                                   # tuned preset kernel: d * G
    G = (ec.g.x, ec.g.y)
    assert cv.mul_batch([G] * n, ds) == pts                             # run-time path agrees with the preset path
    k1 = [rnd.randrange(2 ** (8 * cv.len)) for _ in range(n)]
    k2 = [rnd.randrange(ec.n) for _ in range(n)]
    k2[0] = 0
    got = cv.mul_add_batch([G] * n, k1, pts, k2)
    assert got == GpuEC(name).mul_add_batch(k1, pts, k2)                # G.mulAdd(k1, P, k2) on the preset kernels
    for i in range(0, n, 6):
        P = c.point(*pts[i])
        w = ec.g.mul(k1[i] % ec.n).add(P.mul(k2[i]))
        assert got[i] == _aff(w)
    assert cv.dbl_batch(pts[:8]) == [_aff(c.point(*q).dbl()) for q in pts[:8]]
    assert cv.add_batch(pts[:8], pts[8:16]) == [_aff(c.point(*a).add(c.point(*b))) for a, b in zip(pts[:8], pts[8:16])]
    neg = [(x, (c.p - y) % c.p) for x, y in pts[:4]]
    assert cv.add_batch(pts[:4], neg) == [None] * 4                     # P + (-P)
    assert cv.mul_batch(pts[:2], [ec.n, ec.n + 1]) == [None, pts[1]]
    assert cv.validate_batch(pts[:4] + [(pts[0][0], (pts[0][1] + 1) % c.p)]) == [True] * 4 + [False]
