"""ORACLE (test infrastructure only).  Restates lib/elliptic/eddsa/{index,key,
signature}.js for ed25519."""

import hashlib

from . import curves
from .bn import ref_assert
from .utils import to_array, int_from_le


def _bytes(x):
    """utils.parseBytes (utils.js:112-116): hex string or byte array."""
    return bytes(to_array(x, "hex")) if isinstance(x, str) else bytes(x)


class EDDSA:
    """eddsa/index.js:11-25."""

    def __init__(self, curve="ed25519"):
        ref_assert(curve == "ed25519", "only tested with ed25519 so far")
        c = curves.get(curve).curve
        self.curve = c
        self.g = c.g
        self.g.precompute(c.n.bit_length() + 1)
        self.encoding_length = (c.n.bit_length() + 7) // 8
        self.hash = hashlib.sha512

    def hash_int(self, *parts):
        """eddsa/index.js:65-70."""
        h = self.hash()
        for p in parts:
            h.update(bytes(p))
        return int_from_le(h.digest()) % self.curve.n

    def encode_point(self, point):
        """eddsa/index.js:94-98."""
        enc = bytearray(point.get_y().to_bytes(self.encoding_length, "little"))
        enc[-1] |= 0x80 if point.get_x() & 1 else 0
        return bytes(enc)

    def decode_point(self, data):
        """eddsa/index.js:100-109."""
        b = bytearray(_bytes(data))
        x_is_odd = (b[-1] & 0x80) != 0
        b[-1] &= 0x7F
        return self.curve.point_from_y(int_from_le(b), x_is_odd)

    def priv_from_secret(self, secret):
        """eddsa/key.js:52-71."""
        h = self.hash(_bytes(secret)).digest()
        a = bytearray(h[:self.encoding_length])
        a[0] &= 248
        a[-1] &= 127
        a[-1] |= 64
        return int_from_le(a), h[self.encoding_length:]

    def sign(self, message, secret):
        """eddsa/index.js:34-44."""
        message = _bytes(message)
        priv, prefix = self.priv_from_secret(secret)
        pub_bytes = self.encode_point(self.g.mul(priv))
        r = self.hash_int(prefix, message)
        R = self.g.mul(r)
        r_enc = self.encode_point(R)
        s_ = self.hash_int(r_enc, pub_bytes, message) * priv
        S = (r + s_) % self.curve.n
        return r_enc + S.to_bytes(self.encoding_length, "little")

    def verify(self, message, sig, pub):
        """eddsa/index.js:52-63.  Returns bool or raises RefError."""
        message = _bytes(message)
        sig = _bytes(sig)
        # eddsa/signature.js:23-24
        ref_assert(len(sig) == self.encoding_length * 2, "Signature has invalid size")
        r_enc, s_enc = sig[:self.encoding_length], sig[self.encoding_length:]
        S = int_from_le(s_enc)
        if S >= self.curve.n:
            return False
        pub_bytes = _bytes(pub)
        h = self.hash_int(r_enc, pub_bytes, message)
        SG = self.g.mul(S)
        R = self.decode_point(r_enc)          # sig.R()  (may throw)
        A = self.decode_point(pub_bytes)      # key.pub() (may throw)
        RplusAh = R.add(A.mul(h))
        return RplusAh.eq(SG)
