"""ORACLE (test infrastructure only).  Restates lib/elliptic/curve/edwards.js
for the extended-coordinate twisted case the path uses (ed25519: a = -1,
c = 1), plus the base.js strategies it inherits."""

from .bn import Red, RefError, ref_assert
from .utils import get_naf
from .short import Precomputed, _get


class EPoint:
    """edwards.js:114-144 (extended coordinates X:Y:Z:T)."""

    type = "projective"

    def __init__(self, curve, x=None, y=None, z=None, t=None):
        self.curve = curve
        self.precomputed = None
        red = curve.red
        if x is None and y is None and z is None:
            self.x, self.y, self.z, self.t = 0, 1, 1, 0
            self.z_one = True
        else:
            self.x = red.conv(x)
            self.y = red.conv(y)
            self.z_one = z is None
            self.z = 1 if z is None else red.conv(z)
            self.t = None if t is None else red.conv(t)
            if curve.extended and self.t is None:
                self.t = red.mul(self.x, self.y)
                if not self.z_one:
                    self.t = red.mul(self.t, red.invm(self.z))

    def validate(self):
        return self.curve.validate(self)

    def is_infinity(self):
        """edwards.js:167-172."""
        return self.x == 0 and (self.y == self.z or (self.z_one and self.y == self.curve.c))

    def _ext_dbl(self):
        """edwards.js:174-205  (4M + 4S)."""
        red = self.curve.red
        a = red.sqr(self.x)
        b = red.sqr(self.y)
        c = red.sqr(self.z)
        c = red.add(c, c)
        d = self.curve._mul_a(a)
        e = red.sub(red.sub(red.sqr(red.add(self.x, self.y)), a), b)
        g = red.add(d, b)
        f = red.sub(g, c)
        h = red.sub(d, b)
        return EPoint(self.curve, red.mul(e, f), red.mul(g, h), red.mul(f, g), red.mul(e, h))

    def dbl(self):
        if self.is_infinity():
            return self
        ref_assert(self.curve.extended, "oracle: only extended Edwards restated")
        return self._ext_dbl()

    def _ext_add(self, p):
        """edwards.js:279-309  (add-2008-hwcd-3)."""
        red = self.curve.red
        a = red.mul(red.sub(self.y, self.x), red.sub(p.y, p.x))
        b = red.mul(red.add(self.y, self.x), red.add(p.y, p.x))
        c = red.mul(red.mul(self.t, self.curve.dd), p.t)
        d = red.mul(self.z, red.add(p.z, p.z))
        e = red.sub(b, a)
        f = red.sub(d, c)
        g = red.add(d, c)
        h = red.add(b, a)
        return EPoint(self.curve, red.mul(e, f), red.mul(g, h), red.mul(f, g), red.mul(e, h))

    def add(self, p):
        """edwards.js:350-360."""
        if self.is_infinity():
            return p
        if p.is_infinity():
            return self
        return self._ext_add(p)

    mixed_add = add  # edwards.js:434

    def neg(self, _pre=False):
        red = self.curve.red
        return EPoint(self.curve, red.neg(self.x), self.y, self.z,
                      None if self.t is None else red.neg(self.t))

    def normalize(self):
        """edwards.js:377-390."""
        if self.z_one:
            return self
        red = self.curve.red
        zi = red.invm(self.z)
        self.x = red.mul(self.x, zi)
        self.y = red.mul(self.y, zi)
        if self.t is not None:
            self.t = red.mul(self.t, zi)
        self.z = 1
        self.z_one = True
        return self

    to_p = normalize

    def get_x(self):
        self.normalize()
        return self.x

    def get_y(self):
        self.normalize()
        return self.y

    def eq(self, other):
        """edwards.js:409-413."""
        return self is other or (self.get_x() == other.get_x() and self.get_y() == other.get_y())

    def eq_x_to_p(self, x):
        """edwards.js:415-431."""
        c = self.curve
        red = c.red
        rx = red.mul(red.conv(x), self.z)
        if self.x == rx:
            return True
        xc = x
        t = red.mul(c.red_n, self.z)
        while True:
            xc += c.n
            if xc >= c.p:
                return False
            rx = red.add(rx, t)
            if self.x == rx:
                return True

    def mul(self, k):
        """edwards.js:362-367."""
        if self._has_doubles(k):
            return self.curve._fixed_naf_mul(self, k)
        return self.curve._wnaf_mul(self, k)

    def mul_add(self, k1, p, k2):
        return self.curve._wnaf_mul_add(1, [self, p], [k1, k2], 2, False)

    def jmul_add(self, k1, p, k2):
        return self.curve._wnaf_mul_add(1, [self, p], [k1, k2], 2, True)

    def dblp(self, k):
        """BasePoint.dblp, base.js:376-381."""
        r = self
        for _ in range(k):
            r = r.dbl()
        return r

    def to_j(self):
        return self

    def _get_beta(self):
        return None

    # shared BasePoint machinery
    from .short import Point as _P
    precompute = _P.precompute
    _has_doubles = _P._has_doubles
    _get_doubles = _P._get_doubles
    _get_naf_points = _P._get_naf_points
    del _P


class EdwardsCurve:
    """edwards.js:10-27 (+ base.js:9-41)."""

    type = "edwards"

    def __init__(self, conf):
        self.twisted = conf["a"] != 1
        self.m_one_a = self.twisted and conf["a"] == -1
        self.extended = self.m_one_a
        self.p = conf["p"]
        self.red = Red(self.p)
        self.n = conf.get("n")
        self._bit_length = self.n.bit_length() if self.n else 0
        adjust = self.n and self.p // self.n
        if not adjust or adjust > 100:
            self.red_n, self._maxwell_trick = None, False
        else:
            self._maxwell_trick = True
            self.red_n = self.red.conv(self.n)
        self.a = conf["a"] % self.p
        self.c = self.red.conv(conf["c"])
        self.c2 = self.red.sqr(self.c)
        self.d = self.red.conv(conf["d"])
        self.dd = self.red.add(self.d, self.d)
        self.one_c = conf["c"] == 1
        self.zero_a = self.three_a = False
        self.endo = None
        self.g = self.point(conf["g"][0], conf["g"][1]) if conf.get("g") else None

    def _mul_a(self, num):
        return self.red.neg(num) if self.m_one_a else self.red.mul(self.a, num)

    def point(self, x=None, y=None, z=None, t=None):
        return EPoint(self, x, y, z, t)

    def jpoint(self, x, y, z, t=None):
        return self.point(x, y, z, t)

    def point_from_x(self, x, odd):
        """edwards.js:46-69."""
        red = self.red
        x = red.conv(x)
        x2 = red.sqr(x)
        rhs = red.sub(self.c2, red.mul(self.a, x2))
        lhs = red.sub(1, red.mul(red.mul(self.c2, self.d), x2))
        y2 = red.mul(rhs, red.invm(lhs))
        y = red.sqrt(y2)
        if red.sub(red.sqr(y), y2) != 0:
            raise RefError("invalid point")
        is_odd = bool(y & 1)
        if (odd and not is_odd) or (not odd and is_odd):
            y = red.neg(y)
        return self.point(x, y)

    def point_from_y(self, y, odd):
        """edwards.js:71-97."""
        red = self.red
        y = red.conv(y)
        y2 = red.sqr(y)
        lhs = red.sub(y2, self.c2)
        rhs = red.sub(red.mul(red.mul(y2, self.d), self.c2), self.a)
        x2 = red.mul(lhs, red.invm(rhs))
        if x2 == 0:
            if odd:
                raise RefError("invalid point")
            return self.point(0, y)
        x = red.sqrt(x2)
        if red.sub(red.sqr(x), x2) != 0:
            raise RefError("invalid point")
        if bool(x & 1) != bool(odd):
            x = red.neg(x)
        return self.point(x, y)

    def validate(self, point):
        """edwards.js:99-112."""
        red = self.red
        if point.is_infinity():
            return True
        point.normalize()
        x2 = red.sqr(point.x)
        y2 = red.sqr(point.y)
        lhs = red.add(red.mul(x2, self.a), y2)
        rhs = red.mul(self.c2, red.add(1, red.mul(red.mul(self.d, x2), y2)))
        return lhs == rhs

    from .short import ShortCurve as _S
    decode_point = _S.decode_point          # BaseCurve.decodePoint (base.js:270-292) is shared by all curve types
    _fixed_naf_mul = _S._fixed_naf_mul
    _wnaf_mul = _S._wnaf_mul
    _wnaf_mul_add = _S._wnaf_mul_add
    del _S
