"""ORACLE (test infrastructure only -- never imported by the product package).

Python restatement of the handful of bn.js behaviours the hot path depends on.
bn.js 4.11.9 is a third-party dependency of the reference (package.json:48,
package-lock.json:598-600); its source is vendored in the reference bundle at
dist/elliptic.js:3950-7382 and every function below cites that copy.

Values are plain Python ints (the canonical residue bn.js holds after
`fromRed()`); only behaviour that is *observable* at the path's outputs is
restated: rounding of divRound, invm(0) == 0, which inputs make sqrt fail.
A global op counter records field multiplications / squarings / inversions so
BASELINE.md's per-operation work table can be re-emitted from committed code.
"""

COUNT = {"M": 0, "S": 0, "I": 0}


class RefError(Exception):
    """An `Error` the reference would throw; .args[0] is the JS message."""


def ref_assert(cond, msg="Assertion failed"):
    """minimalistic-assert (dist:8832-8835) and bn.js's private assert
    (dist:3956-3958): throw new Error(msg || 'Assertion failed')."""
    if not cond:
        raise RefError(msg)


def reset_count():
    for k in COUNT:
        COUNT[k] = 0


def snapshot_count():
    return dict(COUNT)


def div_round(a, b):
    """BN.prototype.divRound, dist/elliptic.js:6387-6404.

    divmod truncates toward zero (sign of mod follows the dividend); the
    rounding rule is "half rounds up in magnitude" unless the divisor is odd
    and the remainder is exactly floor(b/2).
    """
    assert b > 0
    neg = a < 0
    q, m = divmod(abs(a), b)
    if neg:
        q, m = -q, -m
    if m == 0:
        return q
    mod = m - b if q < 0 else m  # dist:6393 (dm.div.negative ? mod - num : mod)
    # NB: for a negative dividend with |a| < b the quotient is 0 (not negative)
    # and bn.js keeps the negative remainder; cmp below is a signed compare.
    half = b >> 1
    r2 = b & 1
    if mod < half or (r2 == 1 and mod == half):
        return q
    return q - 1 if q < 0 else q + 1


def invmp(a, p):
    """BN.prototype._invmp, dist/elliptic.js:6518-6582 (binary ext. GCD).

    Observable behaviour: a^-1 mod p for gcd(a,p)=1, and **0 for a == 0**
    (the loop never runs and x2 = 0 is returned, dist:6568-6579).
    """
    a %= p
    if a == 0:
        return 0
    return pow(a, -1, p)


class Red:
    """Reduction context: BN.red(name) / BN.mont(m), dist/elliptic.js:7078-7381.

    K256 / P25519 pseudo-Mersenne folding (dist:6904-7051) and Montgomery
    (dist:7312-7381) only change the internal representation; `fromRed()`
    always yields the canonical residue, which is what this class stores.
    """

    def __init__(self, m):
        self.m = m

    def conv(self, x):
        # Red.prototype.convertTo, dist:7292-7296: num.umod(m)  (quirk Q2)
        return x % self.m

    def add(self, a, b):
        return (a + b) % self.m

    def sub(self, a, b):
        return (a - b) % self.m

    def neg(self, a):
        return (-a) % self.m

    def mul(self, a, b):
        COUNT["M"] += 1
        return (a * b) % self.m

    def sqr(self, a):
        COUNT["S"] += 1
        return (a * a) % self.m

    def invm(self, a):
        # Red.prototype.invm, dist:7234-7242
        COUNT["I"] += 1
        return invmp(a, self.m)

    def pow(self, a, e):
        """Red.prototype.pow, dist:7244-7290 (window 4). Result-exact; the op
        count follows the reference's windowing so fm totals are comparable."""
        if e == 0:
            return 1
        if e == 1:
            return a
        wnd = [1, a]
        for i in range(2, 16):
            wnd.append(self.mul(wnd[i - 1], a))
        res = None  # stands for wnd[0] identity object
        current = 0
        current_len = 0
        nbits = e.bit_length()
        for pos in range(nbits - 1, -1, -1):
            bit = (e >> pos) & 1
            if res is not None:
                res = self.sqr(res)
            if bit == 0 and current == 0:
                current_len = 0
                continue
            current = (current << 1) | bit
            current_len += 1
            if current_len != 4 and pos != 0:
                continue
            res = self.mul(1 if res is None else res, wnd[current])
            current_len = 0
            current = 0
        return 1 if res is None else res

    def sqrt(self, a):
        """Red.prototype.sqrt, dist:7177-7232.

        p % 4 == 3 -> a^((p+1)/4); otherwise Tonelli-Shanks.  For a
        non-residue the returned value is garbage and callers detect it by
        squaring (short.js:194-195, edwards.js:90-91, mont.js:27).
        """
        if a == 0:
            return 0
        m = self.m
        if m & 3 == 3:
            return self.pow(a, (m + 1) >> 2)
        q = m - 1
        s = 0
        while q != 0 and q & 1 == 0:
            s += 1
            q >>= 1
        one = 1
        n_one = m - 1
        lpow = (m - 1) >> 1
        z = m.bit_length()
        z = (2 * z * z) % m
        while self.pow(z, lpow) != n_one:
            z = self.add(z, n_one)
        c = self.pow(z, q)
        r = self.pow(a, (q + 1) >> 1)
        t = self.pow(a, q)
        mm = s
        while t != one:
            tmp = t
            i = 0
            while tmp != one:
                tmp = self.sqr(tmp)
                i += 1
                if i >= mm:
                    # bn.js `assert(i < m)` (dist:7220): for a non-residue
                    # t = a^q has order exactly 2^s, so this ALWAYS fires and
                    # the caller sees Error('Assertion failed') -- not its own
                    # 'invalid point' / validate()==false path.
                    raise RefError("Assertion failed")
            b = self.pow(c, 1 << (mm - i - 1))
            r = self.mul(r, b)
            c = self.sqr(b)
            t = self.mul(t, c)
            mm = i
        return r
