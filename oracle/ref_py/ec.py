"""ORACLE (test infrastructure only).  Restates lib/elliptic/ec/index.js and
lib/elliptic/ec/key.js (ECDSA verify/sign, ECDH derive, key recovery)."""

import hmac as _hmac

from . import curves
from .bn import RefError, ref_assert
from .signature import Signature, _bn, _parse_hex
from .utils import to_array


class HmacDRBG:
    """hmac-drbg 1.0.1, dist/elliptic.js:8686-8800 (RFC 6979 nonce source)."""

    def __init__(self, hash_fn, entropy, nonce, pers=b""):
        self.h = hash_fn
        out = hash_fn().digest_size
        self.K = b"\x00" * out
        self.V = b"\x01" * out
        self._update(bytes(entropy) + bytes(nonce) + bytes(pers))

    def _mac(self, *parts):
        m = _hmac.new(self.K, digestmod=self.h)
        for p in parts:
            m.update(p)
        return m.digest()

    def _update(self, seed=None):
        self.K = self._mac(self.V, b"\x00", seed or b"")
        self.V = self._mac(self.V)
        if not seed:
            return
        self.K = self._mac(self.V, b"\x01", seed)
        self.V = self._mac(self.V)

    def generate(self, ln):
        temp = b""
        while len(temp) < ln:
            self.V = self._mac(self.V)
            temp += self.V
        self._update(None)
        return temp[:ln]


class KeyPair:
    """ec/key.js:7-18."""

    def __init__(self, ec, priv=None, pub=None, pub_enc=None):
        self.ec = ec
        self.priv = None
        self.pub = None
        if priv is not None:
            # _importPrivate, ec/key.js:76-82 (reduced mod n)
            self.priv = _bn(priv) % ec.curve.n
        if pub is not None:
            self._import_public(pub, pub_enc)

    def _import_public(self, key, enc):
        """ec/key.js:84-99.  NB: {x,y} and uncompressed keys are NOT checked
        to be on the curve (quirk Q1)."""
        c = self.ec.curve
        # JS truthiness: a BN object (even 0) and a non-empty string are truthy
        tr = lambda v: v is not None and v != "" and not (isinstance(v, (bytes, list)) and len(v) == 0)
        if isinstance(key, dict) and (tr(key.get("x")) or tr(key.get("y"))):
            if c.type == "mont":
                ref_assert(tr(key.get("x")), "Need x coordinate")
                self.pub = c.point(_bn(key["x"]), 1)
                return
            ref_assert(tr(key.get("x")) and tr(key.get("y")), "Need both x and y coordinate")
            self.pub = c.point(_bn(key["x"]), _bn(key["y"]))
            return
        if hasattr(key, "curve") and hasattr(key, "is_infinity"):
            self.pub = key  # already a point (KeyPair.fromPublic with a Point -> decodePoint would choke; tests pass points via getPublic())
            return
        self.pub = c.decode_point(key, enc)

    def get_public(self):
        if self.pub is None:
            self.pub = self.ec.g.mul(self.priv)
        return self.pub

    def derive(self, pub):
        """ec/key.js:102-107 (ECDH)."""
        if not pub.validate():
            ref_assert(pub.validate(), "public point not validated")
        return pub.mul(self.priv).get_x()


class EC:
    """ec/index.js:13-40."""

    def __init__(self, name, hash_fn=None):
        preset = curves.get(name)
        self.curve = preset.curve
        self.n = self.curve.n
        self.nh = self.n >> 1
        self.g = preset.g
        self.g.precompute(self.n.bit_length() + 1)
        self.hash = hash_fn or preset.hash

    def key_from_private(self, priv):
        return priv if isinstance(priv, KeyPair) else KeyPair(self, priv=priv)

    def key_from_public(self, pub, enc=None):
        return pub if isinstance(pub, KeyPair) else KeyPair(self, pub=pub, pub_enc=enc)

    def gen_key_pair(self, entropy, pers=b""):
        """ec/index.js:55-79 with options.entropy given (bytes, after entropyEnc decoding): HmacDRBG(hash, entropy,
        nonce = n.toArray(), pers); first candidate priv <= n - 2, plus one."""
        entropy = bytes(entropy)
        ref_assert(len(entropy) >= 24, "Not enough entropy. Minimum is: 192 bits")     # hmac-drbg ctor, dist:8708-8710
        nbytes = (self.n.bit_length() + 7) // 8
        drbg = HmacDRBG(self.hash, entropy, self.n.to_bytes(nbytes, "big"), pers)
        ns2 = self.n - 2
        while True:
            priv = int.from_bytes(drbg.generate(nbytes), "big")
            if priv > ns2:
                continue
            return KeyPair(self, priv=priv + 1)

    def _truncate_to_n(self, msg, trunc_only=False, bit_length=None):
        """ec/index.js:81-108 (quirk Q3).  msg: int (BN), hex str or bytes."""
        if isinstance(msg, int):
            v = msg
            byte_length = (v.bit_length() + 7) // 8
        elif isinstance(msg, str):
            byte_length = (len(msg) + 1) >> 1
            v = _parse_hex(msg)
        else:
            byte_length = len(msg)
            v = int.from_bytes(bytes(msg), "big")
        if bit_length is None:
            bit_length = byte_length * 8
        delta = bit_length - self.n.bit_length()
        if delta > 0:
            v >>= delta
        if not trunc_only and v >= self.n:
            return v - self.n
        return v

    def sign(self, msg, key, canonical=False, k_fn=None, pers=b"", msg_bit_length=None):
        """ec/index.js:110-186."""
        key = self.key_from_private(key)
        msg = self._truncate_to_n(msg, False, msg_bit_length)
        nbytes = (self.n.bit_length() + 7) // 8
        bkey = key.priv.to_bytes(nbytes, "big")
        nonce = msg.to_bytes(nbytes, "big")
        drbg = HmacDRBG(self.hash, bkey, nonce, pers)
        ns1 = self.n - 1
        it = 0
        while True:
            k = k_fn(it) if k_fn else int.from_bytes(drbg.generate(nbytes), "big")
            it += 1
            k = self._truncate_to_n(k, True)
            if k <= 1 or k >= ns1:
                continue
            kp = self.g.mul(k)
            if kp.is_infinity():
                continue
            kpx = kp.get_x()
            r = kpx % self.n
            if r == 0:
                continue
            s = (pow(k, -1, self.n) * (r * key.priv + msg)) % self.n
            if s == 0:
                continue
            rp = (1 if kp.get_y() & 1 else 0) | (2 if kpx != r else 0)
            if canonical and s > self.nh:
                s = self.n - s
                rp ^= 1
            return Signature({"r": r, "s": s, "recoveryParam": rp})

    def verify(self, msg, signature, key, enc=None, msg_bit_length=None):
        """ec/index.js:188-229.  Returns bool or raises RefError (quirk Q5)."""
        msg = self._truncate_to_n(msg, False, msg_bit_length)
        key = self.key_from_public(key, enc)
        signature = Signature(signature, "hex")
        r, s = signature.r, signature.s
        if r < 1 or r >= self.n:
            return False
        if s < 1 or s >= self.n:
            return False
        sinv = pow(s, -1, self.n)          # s.invm(n): egcd, dist:6624
        u1 = (sinv * msg) % self.n
        u2 = (sinv * r) % self.n
        if not self.curve._maxwell_trick:
            p = self.g.mul_add(u1, key.get_public(), u2)
            if p.is_infinity():
                return False
            return p.get_x() % self.n == r
        p = self.g.jmul_add(u1, key.get_public(), u2)
        if p.is_infinity():
            return False
        return p.eq_x_to_p(r)

    def recover_pub_key(self, msg, signature, j, enc=None):
        """ec/index.js:231-259."""
        ref_assert((3 & j) == j, "The recovery param is more than two bits")
        signature = Signature(signature, enc)
        n = self.n
        e = _bn(msg)
        r, s = signature.r, signature.s
        is_y_odd = j & 1
        is_second = j >> 1
        if r >= self.curve.p % n and is_second:
            raise RefError("Unable to find sencond key candinate")
        if is_second:
            rp = self.curve.point_from_x(r + n, is_y_odd)
        else:
            rp = self.curve.point_from_x(r, is_y_odd)
        # BN.invm -> egcd (dist:6436-6516, 6624-6626): for r = 0 (mod n) the returned cofactor is 0
        r_inv = pow(signature.r % n, -1, n) if signature.r % n else 0
        s1 = ((n - e) * r_inv) % n
        s2 = (s * r_inv) % n
        return self.g.mul_add(s1, rp, s2)
