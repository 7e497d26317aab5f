"""ORACLE (test infrastructure only).  Restates lib/elliptic/curve/mont.js."""

from .bn import Red
from .utils import to_array


class MPoint:
    """mont.js:30-43 (x-only projective X:Z)."""

    type = "projective"

    def __init__(self, curve, x=None, z=None):
        self.curve = curve
        self.precomputed = None
        if x is None and z is None:
            self.x, self.z = 1, 0
        else:
            self.x = curve.red.conv(x)
            self.z = curve.red.conv(z)

    def precompute(self, power=None):
        return self  # mont.js:58-60 no-op

    def validate(self):
        return self.curve.validate(self)

    def is_infinity(self):
        return self.z == 0

    def dbl(self):
        """mont.js:82-101."""
        red = self.curve.red
        a = red.add(self.x, self.z)
        aa = red.sqr(a)
        b = red.sub(self.x, self.z)
        bb = red.sqr(b)
        c = red.sub(aa, bb)
        nx = red.mul(aa, bb)
        nz = red.mul(c, red.add(bb, red.mul(self.curve.a24, c)))
        return MPoint(self.curve, nx, nz)

    def diff_add(self, p, diff):
        """mont.js:107-128."""
        red = self.curve.red
        a = red.add(self.x, self.z)
        b = red.sub(self.x, self.z)
        c = red.add(p.x, p.z)
        d = red.sub(p.x, p.z)
        da = red.mul(d, a)
        cb = red.mul(c, b)
        nx = red.mul(diff.z, red.sqr(red.add(da, cb)))
        nz = red.mul(diff.x, red.sqr(red.sub(da, cb)))
        return MPoint(self.curve, nx, nz)

    def mul(self, k):
        """mont.js:130-153 (ladder over the bits of k, MSB first, no clamping)."""
        a = self
        b = self.curve.point(None, None)
        c = self
        for i in range(k.bit_length() - 1, -1, -1):
            if (k >> i) & 1 == 0:
                a = a.diff_add(b, c)
                b = b.dbl()
            else:
                b = a.diff_add(b, c)
                a = a.dbl()
        return b

    def normalize(self):
        """mont.js:167-171 (invm(0) == 0, so infinity -> x = 0)."""
        red = self.curve.red
        self.x = red.mul(self.x, red.invm(self.z))
        self.z = 1
        return self

    def get_x(self):
        self.normalize()
        return self.x

    def eq(self, other):
        return self.get_x() == other.get_x()


class MontCurve:
    """mont.js:9-17."""

    type = "mont"

    def __init__(self, conf):
        self.p = conf["p"]
        self.red = Red(self.p)
        self.n = conf.get("n")
        self._bit_length = self.n.bit_length() if self.n else 0
        adjust = self.n and self.p // self.n
        self._maxwell_trick = bool(adjust and adjust <= 100)
        self.red_n = self.red.conv(self.n) if self._maxwell_trick else None
        self.a = self.red.conv(conf["a"])
        self.b = self.red.conv(conf["b"])
        self.i4 = self.red.invm(4)
        self.two = 2
        self.a24 = self.red.mul(self.i4, self.red.add(self.a, self.two))
        self.g = self.point(conf["g"][0], 1) if conf.get("g") else None

    def point(self, x, z):
        return MPoint(self, x, z)

    def decode_point(self, data, enc=None):
        """mont.js:46-48: big-endian bytes -> x, z = 1."""
        return self.point(int.from_bytes(bytes(to_array(data, enc)), "big"), 1)

    def validate(self, point):
        """mont.js:21-28.  NB Red.sqrt throws 'Assertion failed' for a
        non-residue when p = 1 mod 4 (bn.py), so twist points never reach the
        `return false`."""
        red = self.red
        x = point.normalize().x
        x2 = red.sqr(x)
        rhs = red.add(red.add(red.mul(x2, x), red.mul(x2, self.a)), x)
        y = red.sqrt(rhs)
        return red.sqr(y) == rhs
