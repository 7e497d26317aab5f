"""ORACLE (test infrastructure only).

Restates lib/elliptic/curve/base.js (BaseCurve/BasePoint strategies) and
lib/elliptic/curve/short.js (ShortCurve, affine Point, Jacobian JPoint) with
the reference's control flow -- including the places where its behaviour is
not a pure group-law function (un-validated points).
Field values are canonical residues (Python ints); see bn.py.
"""

from .bn import Red, RefError, ref_assert, div_round
from .utils import get_naf, get_jsf, to_array


def _get(lst, i):
    # JS `naf[j][i] | 0` on a missing index -> 0
    return lst[i] if 0 <= i < len(lst) else 0


class Precomputed:
    __slots__ = ("naf", "doubles", "beta")

    def __init__(self, naf=None, doubles=None, beta=None):
        self.naf = naf          # (wnd, [points])
        self.doubles = doubles  # (step, [points])
        self.beta = beta


class Point:
    """Affine point, lib/elliptic/curve/short.js:251-271 (+ BasePoint)."""

    type = "affine"

    def __init__(self, curve, x, y):
        self.curve = curve
        self.precomputed = None
        if x is None and y is None:
            self.x = self.y = None
            self.inf = True
        else:
            # short.js:258-268 toRed -> umod p (quirk Q2: oversize coords)
            self.x = curve.red.conv(x)
            self.y = curve.red.conv(y)
            self.inf = False

    # -- BasePoint (base.js:256-370) ------------------------------------
    def validate(self):
        return self.curve.validate(self)

    def precompute(self, power):
        """BasePoint.precompute, base.js:312-327."""
        if self.precomputed:
            return self
        pre = Precomputed()
        pre.naf = self._get_naf_points(8)
        pre.doubles = self._get_doubles(4, power)
        pre.beta = self._get_beta()
        self.precomputed = pre
        return self

    def _has_doubles(self, k):
        """base.js:329-338."""
        if not self.precomputed or not self.precomputed.doubles:
            return False
        step, pts = self.precomputed.doubles
        return len(pts) >= -(-(k.bit_length() + 1) // step)

    def _get_doubles(self, step=None, power=None):
        """base.js:340-355."""
        if self.precomputed and self.precomputed.doubles:
            return self.precomputed.doubles
        doubles = [self]
        acc = self
        i = 0
        while i < power:
            for _ in range(step):
                acc = acc.dbl()
            doubles.append(acc)
            i += step
        return (step, doubles)

    def _get_naf_points(self, wnd):
        """base.js:357-370."""
        if self.precomputed and self.precomputed.naf:
            return self.precomputed.naf
        res = [self]
        mx = (1 << wnd) - 1
        dbl = None if mx == 1 else self.dbl()
        for i in range(1, mx):
            res.append(res[i - 1].add(dbl))
        return (wnd, res)

    def _get_beta(self):
        """Point._getBeta, short.js:282-310."""
        c = self.curve
        if not c.endo:
            return None
        pre = self.precomputed
        if pre and pre.beta:
            return pre.beta
        beta = c.point(c.red.mul(self.x, c.endo["beta"]), self.y)
        if pre:
            def endo_mul(p):
                return c.point(c.red.mul(p.x, c.endo["beta"]), p.y)
            pre.beta = beta
            beta.precomputed = Precomputed(
                naf=pre.naf and (pre.naf[0], [endo_mul(p) for p in pre.naf[1]]),
                doubles=pre.doubles and (pre.doubles[0], [endo_mul(p) for p in pre.doubles[1]]),
            )
        return beta

    # -- short.js Point -----------------------------------------------------
    def is_infinity(self):
        return self.inf

    def eq(self, p):
        """short.js:452-456."""
        return self is p or (self.inf == p.inf and
                             (self.inf or (self.x == p.x and self.y == p.y)))

    def neg(self, _precompute=False):
        """short.js:458-480."""
        if self.inf:
            return self
        res = self.curve.point(self.x, self.curve.red.neg(self.y))
        if _precompute and self.precomputed:
            pre = self.precomputed
            res.precomputed = Precomputed(
                naf=pre.naf and (pre.naf[0], [p.neg() for p in pre.naf[1]]),
                doubles=pre.doubles and (pre.doubles[0], [p.neg() for p in pre.doubles[1]]),
            )
        return res

    def add(self, p):
        """short.js:365-392 (affine add; slope 0 needs no inversion)."""
        red = self.curve.red
        if self.inf:
            return p
        if p.inf:
            return self
        if self.eq(p):
            return self.dbl()
        if self.neg().eq(p):
            return self.curve.point(None, None)
        if self.x == p.x:
            return self.curve.point(None, None)
        c = red.sub(self.y, p.y)
        if c != 0:
            c = red.mul(c, red.invm(red.sub(self.x, p.x)))
        nx = red.sub(red.sub(red.sqr(c), self.x), p.x)
        ny = red.sub(red.mul(c, red.sub(self.x, nx)), self.y)
        return self.curve.point(nx, ny)

    def dbl(self):
        """short.js:394-412."""
        red = self.curve.red
        if self.inf:
            return self
        ys1 = red.add(self.y, self.y)
        if ys1 == 0:
            return self.curve.point(None, None)
        a = self.curve.a
        x2 = red.sqr(self.x)
        dyinv = red.invm(ys1)
        c = red.mul(red.add(red.add(red.add(x2, x2), x2), a), dyinv)
        nx = red.sub(red.sqr(c), red.add(self.x, self.x))
        ny = red.sub(red.mul(c, red.sub(self.x, nx)), self.y)
        return self.curve.point(nx, ny)

    def get_x(self):
        return self.x

    def get_y(self):
        return self.y

    def mul(self, k):
        """short.js:422-432."""
        if self.is_infinity():
            return self
        elif self._has_doubles(k):
            return self.curve._fixed_naf_mul(self, k)
        elif self.curve.endo:
            return self.curve._endo_wnaf_mul_add([self], [k])
        else:
            return self.curve._wnaf_mul(self, k)

    def mul_add(self, k1, p2, k2):
        """short.js:434-441."""
        if self.curve.endo:
            return self.curve._endo_wnaf_mul_add([self, p2], [k1, k2])
        return self.curve._wnaf_mul_add(1, [self, p2], [k1, k2], 2)

    def jmul_add(self, k1, p2, k2):
        """short.js:443-450."""
        if self.curve.endo:
            return self.curve._endo_wnaf_mul_add([self, p2], [k1, k2], True)
        return self.curve._wnaf_mul_add(1, [self, p2], [k1, k2], 2, True)

    def to_j(self):
        """short.js:482-488."""
        if self.inf:
            return self.curve.jpoint(None, None, None)
        return JPoint(self.curve, self.x, self.y, 1, z_one=True)

    def encode(self, compact=False):
        """BasePoint._encode, base.js:298-306."""
        ln = (self.curve.p.bit_length() + 7) // 8
        x = self.get_x().to_bytes(ln, "big")
        if compact:
            return bytes([0x02 if self.get_y() % 2 == 0 else 0x03]) + x
        return b"\x04" + x + self.get_y().to_bytes(ln, "big")


class JPoint:
    """Jacobian point, short.js:490-509."""

    type = "jacobian"

    def __init__(self, curve, x, y, z, z_one=False):
        self.curve = curve
        self.precomputed = None
        if x is None and y is None and z is None:
            self.x = 1
            self.y = 1
            self.z = 0
            z_one = False
        else:
            self.x = x
            self.y = y
            self.z = z
        # short.js:508: zOne is an object-identity test against curve.one, so
        # it is only ever true for points made by Point.toJ().
        self.z_one = z_one

    def is_infinity(self):
        return self.z == 0  # short.js:935-938

    def to_p(self):
        """short.js:516-526."""
        red = self.curve.red
        if self.is_infinity():
            return self.curve.point(None, None)
        zinv = red.invm(self.z)
        zinv2 = red.sqr(zinv)
        ax = red.mul(self.x, zinv2)
        ay = red.mul(red.mul(self.y, zinv2), zinv)
        return self.curve.point(ax, ay)

    def neg(self):
        return JPoint(self.curve, self.x, self.curve.red.neg(self.y), self.z)

    def add(self, p):
        """short.js:532-567  (12M + 4S)."""
        red = self.curve.red
        if self.is_infinity():
            return p
        if p.is_infinity():
            return self
        pz2 = red.sqr(p.z)
        z2 = red.sqr(self.z)
        u1 = red.mul(self.x, pz2)
        u2 = red.mul(p.x, z2)
        s1 = red.mul(self.y, red.mul(pz2, p.z))
        s2 = red.mul(p.y, red.mul(z2, self.z))
        h = red.sub(u1, u2)
        r = red.sub(s1, s2)
        if h == 0:
            if r != 0:
                return self.curve.jpoint(None, None, None)
            return self.dbl()
        h2 = red.sqr(h)
        h3 = red.mul(h2, h)
        v = red.mul(u1, h2)
        nx = red.sub(red.sub(red.add(red.sqr(r), h3), v), v)
        ny = red.sub(red.mul(r, red.sub(v, nx)), red.mul(s1, h3))
        nz = red.mul(red.mul(self.z, p.z), h)
        return JPoint(self.curve, nx, ny, nz)

    def mixed_add(self, p):
        """short.js:569-603  (8M + 3S)."""
        red = self.curve.red
        if self.is_infinity():
            return p.to_j()
        if p.is_infinity():
            return self
        z2 = red.sqr(self.z)
        u1 = self.x
        u2 = red.mul(p.x, z2)
        s1 = self.y
        s2 = red.mul(red.mul(p.y, z2), self.z)
        h = red.sub(u1, u2)
        r = red.sub(s1, s2)
        if h == 0:
            if r != 0:
                return self.curve.jpoint(None, None, None)
            return self.dbl()
        h2 = red.sqr(h)
        h3 = red.mul(h2, h)
        v = red.mul(u1, h2)
        nx = red.sub(red.sub(red.add(red.sqr(r), h3), v), v)
        ny = red.sub(red.mul(r, red.sub(v, nx)), red.mul(s1, h3))
        nz = red.mul(self.z, h)
        return JPoint(self.curve, nx, ny, nz)

    def dblp(self, pw):
        """short.js:605-654."""
        if pw == 0:
            return self
        if self.is_infinity():
            return self
        c = self.curve
        if c.zero_a or c.three_a:
            r = self
            for _ in range(pw):
                r = r.dbl()
            return r
        red = c.red
        a = c.a
        tinv = c.tinv
        jx, jy, jz = self.x, self.y, self.z
        jz4 = red.sqr(red.sqr(jz))
        jyd = red.add(jy, jy)
        for i in range(pw):
            jx2 = red.sqr(jx)
            jyd2 = red.sqr(jyd)
            jyd4 = red.sqr(jyd2)
            cc = red.add(red.add(red.add(jx2, jx2), jx2), red.mul(a, jz4))
            t1 = red.mul(jx, jyd2)
            nx = red.sub(red.sqr(cc), red.add(t1, t1))
            t2 = red.sub(t1, nx)
            dny = red.mul(cc, t2)
            dny = red.sub(red.add(dny, dny), jyd4)
            nz = red.mul(jyd, jz)
            if i + 1 < pw:
                jz4 = red.mul(jz4, jyd4)
            jx, jz, jyd = nx, nz, dny
        return JPoint(c, jx, red.mul(jyd, tinv), jz)

    def dbl(self):
        """short.js:656-666."""
        if self.is_infinity():
            return self
        if self.curve.zero_a:
            return self._zero_dbl()
        elif self.curve.three_a:
            return self._three_dbl()
        return self._dbl()

    def _zero_dbl(self):
        """short.js:668-737 (mdbl-2007-bl when Z is the `one` object, else
        dbl-2009-l).  Both yield the same (X3, Y3, Z3)."""
        red = self.curve.red
        if self.z_one:
            xx = red.sqr(self.x)
            yy = red.sqr(self.y)
            yyyy = red.sqr(yy)
            s = red.sub(red.sub(red.sqr(red.add(self.x, yy)), xx), yyyy)
            s = red.add(s, s)
            m = red.add(red.add(xx, xx), xx)
            t = red.sub(red.sub(red.sqr(m), s), s)
            yyyy8 = red.add(yyyy, yyyy)
            yyyy8 = red.add(yyyy8, yyyy8)
            yyyy8 = red.add(yyyy8, yyyy8)
            nx = t
            ny = red.sub(red.mul(m, red.sub(s, t)), yyyy8)
            nz = red.add(self.y, self.y)
        else:
            a = red.sqr(self.x)
            b = red.sqr(self.y)
            c = red.sqr(b)
            d = red.sub(red.sub(red.sqr(red.add(self.x, b)), a), c)
            d = red.add(d, d)
            e = red.add(red.add(a, a), a)
            f = red.sqr(e)
            c8 = red.add(c, c)
            c8 = red.add(c8, c8)
            c8 = red.add(c8, c8)
            nx = red.sub(red.sub(f, d), d)
            ny = red.sub(red.mul(e, red.sub(d, nx)), c8)
            nz = red.mul(self.y, self.z)
            nz = red.add(nz, nz)
        return JPoint(self.curve, nx, ny, nz)

    def _three_dbl(self):
        """short.js:739-800 (a = -3)."""
        red = self.curve.red
        if self.z_one:
            xx = red.sqr(self.x)
            yy = red.sqr(self.y)
            yyyy = red.sqr(yy)
            s = red.sub(red.sub(red.sqr(red.add(self.x, yy)), xx), yyyy)
            s = red.add(s, s)
            m = red.add(red.add(red.add(xx, xx), xx), self.curve.a)
            t = red.sub(red.sub(red.sqr(m), s), s)
            nx = t
            yyyy8 = red.add(yyyy, yyyy)
            yyyy8 = red.add(yyyy8, yyyy8)
            yyyy8 = red.add(yyyy8, yyyy8)
            ny = red.sub(red.mul(m, red.sub(s, t)), yyyy8)
            nz = red.add(self.y, self.y)
        else:
            delta = red.sqr(self.z)
            gamma = red.sqr(self.y)
            beta = red.mul(self.x, gamma)
            alpha = red.mul(red.sub(self.x, delta), red.add(self.x, delta))
            alpha = red.add(red.add(alpha, alpha), alpha)
            beta4 = red.add(beta, beta)
            beta4 = red.add(beta4, beta4)
            beta8 = red.add(beta4, beta4)
            nx = red.sub(red.sqr(alpha), beta8)
            nz = red.sub(red.sub(red.sqr(red.add(self.y, self.z)), gamma), delta)
            ggamma8 = red.sqr(gamma)
            ggamma8 = red.add(ggamma8, ggamma8)
            ggamma8 = red.add(ggamma8, ggamma8)
            ggamma8 = red.add(ggamma8, ggamma8)
            ny = red.sub(red.mul(alpha, red.sub(beta4, nx)), ggamma8)
        return JPoint(self.curve, nx, ny, nz)

    def _dbl(self):
        """short.js:802-830 (generic a)."""
        red = self.curve.red
        a = self.curve.a
        jx, jy, jz = self.x, self.y, self.z
        jz4 = red.sqr(red.sqr(jz))
        jx2 = red.sqr(jx)
        jy2 = red.sqr(jy)
        c = red.add(red.add(red.add(jx2, jx2), jx2), red.mul(a, jz4))
        jxd4 = red.add(jx, jx)
        jxd4 = red.add(jxd4, jxd4)
        t1 = red.mul(jxd4, jy2)
        nx = red.sub(red.sqr(c), red.add(t1, t1))
        t2 = red.sub(t1, nx)
        jyd8 = red.sqr(jy2)
        jyd8 = red.add(jyd8, jyd8)
        jyd8 = red.add(jyd8, jyd8)
        jyd8 = red.add(jyd8, jyd8)
        ny = red.sub(red.mul(c, t2), jyd8)
        nz = red.mul(red.add(jy, jy), jz)
        return JPoint(self.curve, nx, ny, nz)

    def eq_x_to_p(self, x):
        """JPoint.eqXToP, short.js:908-925 ("Maxwell's trick")."""
        c = self.curve
        red = c.red
        zs = red.sqr(self.z)
        rx = red.mul(red.conv(x), zs)
        if self.x == rx:
            return True
        xc = x
        t = red.mul(c.red_n, zs)
        while True:
            xc += c.n
            if xc >= c.p:
                return False
            rx = red.add(rx, t)
            if self.x == rx:
                return True

    def mul(self, k):
        return self.curve._wnaf_mul(self, k)

    def _get_naf_points(self, wnd):
        return Point._get_naf_points(self, wnd)


class ShortCurve:
    """ShortCurve (short.js:10-24) on top of BaseCurve (base.js:9-41)."""

    type = "short"

    def __init__(self, conf):
        self.p = conf["p"]
        # base.js:14 -- BN.red(name) or BN.mont(p); both are value-transparent.
        self.red = Red(self.p)
        self.n = conf.get("n")
        self._bit_length = self.n.bit_length() if self.n else 0
        # base.js:32-40 generalized Maxwell trick
        adjust = self.n and self.p // self.n
        if not adjust or adjust > 100:
            self.red_n = None
            self._maxwell_trick = False
        else:
            self._maxwell_trick = True
            self.red_n = self.red.conv(self.n)
        self.a = self.red.conv(conf["a"])
        self.b = self.red.conv(conf["b"])
        self.tinv = self.red.invm(2)
        self.zero_a = self.a == 0
        self.three_a = (self.a - self.p) == -3
        self.g = None
        if conf.get("g"):
            self.g = self.point(conf["g"][0], conf["g"][1])
            pre = conf.get("g_pre")
            if pre:
                # Point.fromJSON, short.js:328-352
                self.g.precomputed = Precomputed(
                    doubles=(pre["doubles"]["step"],
                             [self.g] + [self.point(x, y) for x, y in pre["doubles"]["points"]]),
                    naf=(pre["naf"]["wnd"],
                         [self.g] + [self.point(x, y) for x, y in pre["naf"]["points"]]),
                )
        self.endo = self._get_endomorphism(conf)

    def _get_endomorphism(self, conf):
        """short.js:28-75.  Only the preset form (beta, lambda, basis given,
        curves.js:187-198) is restated; derived-basis curves are not on the
        path."""
        if not self.zero_a or not self.g or not self.n or self.p % 3 != 1:
            return None
        ref_assert("beta" in conf and "lambda" in conf and "basis" in conf,
                   "oracle: only preset endomorphism supported")
        return {"beta": self.red.conv(conf["beta"]), "lambda": conf["lambda"],
                "basis": conf["basis"]}

    def point(self, x, y):
        return Point(self, x, y)

    def jpoint(self, x, y, z):
        return JPoint(self, x, y, z)

    def _endo_split(self, k):
        """short.js:168-185."""
        v1, v2 = self.endo["basis"]
        c1 = div_round(v2["b"] * k, self.n)
        c2 = div_round(-v1["b"] * k, self.n)
        p1 = c1 * v1["a"]
        p2 = c2 * v2["a"]
        q1 = c1 * v1["b"]
        q2 = c2 * v2["b"]
        k1 = k - p1 - p2
        k2 = -(q1 + q2)
        return k1, k2

    def point_from_x(self, x, odd):
        """short.js:187-204."""
        red = self.red
        x = red.conv(x)
        y2 = red.add(red.add(red.mul(red.sqr(x), x), red.mul(x, self.a)), self.b)
        y = red.sqrt(y2)
        if red.sub(red.sqr(y), y2) != 0:
            raise RefError("invalid point")
        is_odd = y & 1
        if (odd and not is_odd) or (not odd and is_odd):
            y = red.neg(y)
        return self.point(x, y)

    def validate(self, point):
        """short.js:206-216."""
        red = self.red
        if point.inf:
            return True
        x, y = point.x, point.y
        ax = red.mul(self.a, x)
        rhs = red.add(red.add(red.mul(red.sqr(x), x), ax), self.b)
        return red.sub(red.sqr(y), rhs) == 0

    def decode_point(self, data, enc=None):
        """BaseCurve.decodePoint, base.js:270-292."""
        b = to_array(data, enc)
        ln = (self.p.bit_length() + 7) // 8
        if len(b) and b[0] in (0x04, 0x06, 0x07) and len(b) - 1 == 2 * ln:
            if b[0] == 0x06:
                ref_assert(b[-1] % 2 == 0)
            elif b[0] == 0x07:
                ref_assert(b[-1] % 2 == 1)
            return self.point(int.from_bytes(bytes(b[1:1 + ln]), "big"),
                              int.from_bytes(bytes(b[1 + ln:1 + 2 * ln]), "big"))
        elif len(b) and b[0] in (0x02, 0x03) and len(b) - 1 == ln:
            return self.point_from_x(int.from_bytes(bytes(b[1:1 + ln]), "big"), b[0] == 0x03)
        raise RefError("Unknown point format")

    # -- multiplication strategies (base.js) ------------------------------
    def _fixed_naf_mul(self, p, k):
        """BaseCurve._fixedNafMul, base.js:52-84."""
        ref_assert(p.precomputed)
        step, pts = p._get_doubles()
        naf = get_naf(k, 1, self._bit_length)
        I = (1 << (step + 1)) - (2 if step % 2 == 0 else 1)
        I //= 3
        repr_ = []
        for j in range(0, len(naf), step):
            naf_w = 0
            for l in range(j + step - 1, j - 1, -1):
                naf_w = (naf_w << 1) + _get(naf, l)
            repr_.append(naf_w)
        a = self.jpoint(None, None, None)
        b = self.jpoint(None, None, None)
        for i in range(I, 0, -1):
            for j in range(len(repr_)):
                naf_w = repr_[j]
                if naf_w == i:
                    b = b.mixed_add(pts[j])
                elif naf_w == -i:
                    b = b.mixed_add(pts[j].neg())
            a = a.add(b)
        return a.to_p()

    def _wnaf_mul(self, p, k):
        """BaseCurve._wnafMul, base.js:86-126."""
        w, wnd = p._get_naf_points(4)
        naf = get_naf(k, w, self._bit_length)
        acc = self.jpoint(None, None, None)
        i = len(naf) - 1
        while i >= 0:
            l = 0
            while i >= 0 and naf[i] == 0:
                l += 1
                i -= 1
            if i >= 0:
                l += 1
            acc = acc.dblp(l)
            if i < 0:
                break
            z = naf[i]
            if p.type == "affine":
                if z > 0:
                    acc = acc.mixed_add(wnd[(z - 1) >> 1])
                else:
                    acc = acc.mixed_add(wnd[(-z - 1) >> 1].neg())
            else:
                if z > 0:
                    acc = acc.add(wnd[(z - 1) >> 1])
                else:
                    acc = acc.add(wnd[(-z - 1) >> 1].neg())
            i -= 1
        return acc.to_p() if p.type == "affine" else acc

    def _wnaf_mul_add(self, def_w, points, coeffs, ln, jacobian_result=False):
        """BaseCurve._wnafMulAdd, base.js:128-253."""
        red = self.red
        wnd_width = [None] * ln
        wnd = [None] * ln
        naf = [None] * ln
        mx = 0
        for i in range(ln):
            wnd_width[i], wnd[i] = points[i]._get_naf_points(def_w)
        i = ln - 1
        while i >= 1:
            a = i - 1
            b = i
            if wnd_width[a] != 1 or wnd_width[b] != 1:
                naf[a] = get_naf(coeffs[a], wnd_width[a], self._bit_length)
                naf[b] = get_naf(coeffs[b], wnd_width[b], self._bit_length)
                mx = max(len(naf[a]), mx)
                mx = max(len(naf[b]), mx)
                i -= 2
                continue
            comb = [points[a], None, None, points[b]]
            if points[a].y == points[b].y:
                comb[1] = points[a].add(points[b])
                comb[2] = points[a].to_j().mixed_add(points[b].neg())
            elif points[a].y == red.neg(points[b].y):
                comb[1] = points[a].to_j().mixed_add(points[b])
                comb[2] = points[a].add(points[b].neg())
            else:
                comb[1] = points[a].to_j().mixed_add(points[b])
                comb[2] = points[a].to_j().mixed_add(points[b].neg())
            index = [-3, -1, -5, -7, 0, 7, 5, 1, 3]
            jsf = get_jsf(coeffs[a], coeffs[b])
            mx = max(len(jsf[0]), mx)
            naf[a] = [0] * mx
            naf[b] = [0] * mx
            for j in range(mx):
                ja = _get(jsf[0], j)
                jb = _get(jsf[1], j)
                naf[a][j] = index[(ja + 1) * 3 + (jb + 1)]
                naf[b][j] = 0
            wnd[a] = comb
            i -= 2
        # NB base.js:153: when ln is odd (single-point GLV: ln = 2 always even)
        acc = self.jpoint(None, None, None)
        tmp = [0] * ln
        i = mx
        while i >= 0:
            k = 0
            while i >= 0:
                zero = True
                for j in range(ln):
                    tmp[j] = _get(naf[j], i) if naf[j] is not None else 0
                    if tmp[j] != 0:
                        zero = False
                if not zero:
                    break
                k += 1
                i -= 1
            if i >= 0:
                k += 1
            acc = acc.dblp(k)
            if i < 0:
                break
            for j in range(ln):
                z = tmp[j]
                if z == 0:
                    continue
                elif z > 0:
                    p = wnd[j][(z - 1) >> 1]
                else:
                    p = wnd[j][(-z - 1) >> 1].neg()
                if p.type == "affine":
                    acc = acc.mixed_add(p)
                else:
                    acc = acc.add(p)
            i -= 1
        if jacobian_result:
            return acc
        return acc.to_p()

    def _endo_wnaf_mul_add(self, points, coeffs, jacobian_result=False):
        """ShortCurve._endoWnafMulAdd, short.js:218-249."""
        npoints = []
        ncoeffs = []
        for i in range(len(points)):
            k1, k2 = self._endo_split(coeffs[i])
            p = points[i]
            beta = p._get_beta()
            if k1 < 0:
                k1 = -k1
                p = p.neg(True)
            if k2 < 0:
                k2 = -k2
                beta = beta.neg(True)
            npoints += [p, beta]
            ncoeffs += [k1, k2]
        return self._wnaf_mul_add(1, npoints, ncoeffs, len(npoints), jacobian_result)
