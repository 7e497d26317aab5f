"""Host-side mirror of the reference's ECDSA object for the accelerated path.

`EC(name)` corresponds to `new elliptic.ec(name)` (lib/elliptic/ec/index.js:13-40);
`verify(msg, sig, key)` keeps the reference's argument forms and error
behaviour (ec/index.js:188-229) and `verify_batch` is the new batch entry
point (`EC#verifyBatch(msgs, sigs, pubs) -> Uint8Array`).  Parsing
(hex / byte arrays / DER / SEC1) is done here exactly as the reference's JS
does it; all curve arithmetic happens in libelliptic_b200.so on the GPU.
"""
import ctypes
import re

import numpy as np

from . import _native as nat

_CURVES = {
    # name -> C-ABI id, byte length of a field element, group order n, field prime p (curves.js:43-206)
    "secp256k1": dict(id=nat.CURVE_SECP256K1, len=32,
                      n=0xfffffffffffffffffffffffffffffffebaaedce6af48a03bbfd25e8cd0364141,
                      p=2**256 - 2**32 - 977),
    "p256": dict(id=nat.CURVE_P256, len=32,
                 n=0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551,
                 p=2**256 - 2**224 + 2**192 + 2**96 - 1),
    "curve25519": dict(id=nat.CURVE_CURVE25519, len=32,
                       n=0x1000000000000000000000000000000014def9dea2f79cd65812631a5cf5d3ed, p=2**255 - 19),
    "p384": dict(id=nat.CURVE_P384, len=48,
                 n=0xffffffffffffffffffffffffffffffffffffffffffffffffc7634d81f4372ddf581a0db248b0a77aecec196accc52973,
                 p=2**384 - 2**128 - 2**96 + 2**32 - 1),
    "p521": dict(id=nat.CURVE_P521, len=66,
                 n=0x1fffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffa51868783bf2f966b7fcc0148f709a5d03bb5c9b8899c47aebb6fb71e91386409,
                 p=2**521 - 1),
    "ed25519": dict(id=nat.CURVE_ED25519, len=32,
                    n=0x1000000000000000000000000000000014def9dea2f79cd65812631a5cf5d3ed, p=2**255 - 19),
    "p192": dict(id=nat.CURVE_P192, len=24, n=0xffffffffffffffffffffffff99def836146bc9b1b4d22831, p=0xfffffffffffffffffffffffffffffffeffffffffffffffff),
    "p224": dict(id=nat.CURVE_P224, len=28, n=0xffffffffffffffffffffffffffff16a2e0b8f03e13dd29455c5c2a3d, p=0xffffffffffffffffffffffffffffffff000000000000000000000001),
}


_SHORT = ("secp256k1", "p256", "p384", "p521", "p192", "p224")
# presets whose points are (x, y) pairs with batch sign / keygen / mul / mulAdd / ECDH entry points: the six short
# curves and the twisted Edwards preset (new elliptic.ec('ed25519'), test/ecdsa-test.js:130, test/ecdh-test.js:26)
_XY = _SHORT + ("ed25519",)


class EllipticError(Exception):
    """An `Error` the reference would have thrown (message = the JS message)."""


class NeedsReferencePath(EllipticError):
    """The reference's result for this item is not a group-law function of the
    inputs (un-validated off-curve public key); the engine does
    not guess -- run the reference's single-item path for it."""


_THROW_MSG = {
    nat.ST_THROW_INVALID_POINT: "invalid point",
    nat.ST_THROW_NOT_VALIDATED: "public point not validated",
    nat.ST_THROW_ASSERT: "Assertion failed",
    nat.ST_THROW_POINT_FORMAT: "Unknown point format",
    nat.ST_THROW_SECOND_KEY: "Unable to find sencond key candinate",
    nat.ST_THROW_SIG_FORMAT: "Signature without r or s",
}


def _to_array(msg, enc=None):
    """minimalistic-crypto-utils.toArray (reference dist/elliptic.js:8847-8876)."""
    if isinstance(msg, (bytes, bytearray)):
        return bytes(msg)
    if isinstance(msg, (list, tuple, np.ndarray)):
        return bytes(int(b) & 0xFF for b in msg)
    if not msg:
        return b""
    if isinstance(msg, str):
        if enc == "hex":
            msg = re.sub(r"[^a-zA-Z0-9]+", "", msg)
            if len(msg) % 2:
                msg = "0" + msg
            out = bytearray()
            for i in range(0, len(msg), 2):
                try:
                    out.append(int(msg[i:i + 2], 16))
                except ValueError:
                    out.append(0)
            return bytes(out)
        out = bytearray()
        for ch in msg:
            c = ord(ch)
            if c >> 8:
                out.append(c >> 8)
            out.append(c & 0xFF)
        return bytes(out)
    raise TypeError("unsupported input form")


def _parse_hex(s):
    """bn.js 4.11.9 `new BN(str, 16)` (dist/elliptic.js:4135-4157 parseHex, :4003-4017): whitespace is dropped, a
    leading '-' negates, and a character outside [0-9a-fA-F] contributes (charCode - 48) & 0xf instead of throwing."""
    s = "".join(s.split())
    neg = s.startswith("-")
    if neg:
        s = s[1:]
    v = 0
    for ch in s:
        c = ord(ch) - 48
        if 49 <= c <= 54:
            d = c - 49 + 10
        elif 17 <= c <= 22:
            d = c - 17 + 10
        else:
            d = c & 0xF
        v = (v << 4) | d
    return -v if neg else v


def _bn(v):
    """`new BN(v, 16)` for int / hex string / byte array."""
    if isinstance(v, int):
        return v
    if isinstance(v, str):
        return _parse_hex(v)
    return int.from_bytes(bytes(v), "big")


def parse_der(data):
    """Signature._importDER (ec/signature.js:73-134).  Returns (r, s) or None."""
    n = len(data)
    pos = 0

    def byte(i):
        return data[i] if 0 <= i < n else None

    def get_length():
        nonlocal pos
        initial = byte(pos)
        pos += 1
        if initial is None:
            return 0
        if not (initial & 0x80):
            return initial
        octets = initial & 0xF
        if octets == 0 or octets > 4:
            return None
        if byte(pos) == 0:
            return None
        val = 0
        off = pos
        for _ in range(octets):
            val = ((val << 8) | (byte(off) or 0)) & 0xFFFFFFFF
            off += 1
        if val <= 0x7F:
            return None
        pos = off
        return val

    if byte(pos) != 0x30:
        return None
    pos += 1
    ln = get_length()
    if ln is None or ln + pos != n:
        return None
    if byte(pos) != 0x02:
        return None
    pos += 1
    rlen = get_length()
    if rlen is None or ((byte(pos) or 0) & 0x80):
        return None
    r = data[pos:pos + rlen]
    pos += rlen
    if byte(pos) != 0x02:
        return None
    pos += 1
    slen = get_length()
    if slen is None or n != slen + pos or ((byte(pos) or 0) & 0x80):
        return None
    s = data[pos:pos + slen]
    for v in (r, s):
        if len(v) and v[0] == 0 and not (len(v) > 1 and v[1] & 0x80):
            return None
    return int.from_bytes(r, "big"), int.from_bytes(s, "big")


class EC:
    def __init__(self, curve="secp256k1", device=0):
        if curve not in _CURVES:
            raise EllipticError("Unknown curve " + str(curve))   # ec/index.js:19-20
        self.name = curve
        self._c = _CURVES[curve]
        self.n = self._c["n"]
        self._len = self._c["len"]
        self._device = device

    # ---- reference-compatible scalar preparation (host side, cheap) -------------
    def _truncate_to_n(self, msg, msg_bit_length=None):
        """EC._truncateToN (ec/index.js:81-108), including its single conditional `- n`."""
        if isinstance(msg, int):
            v = msg
            byte_length = (v.bit_length() + 7) // 8
        elif isinstance(msg, str):
            byte_length = (len(msg) + 1) >> 1
            v = _parse_hex(msg)
        else:
            b = _to_array(msg)
            byte_length = len(b)
            v = int.from_bytes(b, "big")
        bit_length = byte_length * 8 if msg_bit_length is None else msg_bit_length
        delta = bit_length - self.n.bit_length()
        if delta > 0:
            v >>= delta
        if v >= self.n:
            v -= self.n
        return v

    def _signature(self, sig):
        """new Signature(sig, 'hex') (ec/signature.js:8-22)."""
        if isinstance(sig, dict):
            if not (sig.get("r") and sig.get("s")):
                raise EllipticError("Signature without r or s")
            return _bn(sig["r"]), _bn(sig["s"])
        if hasattr(sig, "r") and hasattr(sig, "s"):
            return int(sig.r), int(sig.s)
        rs = parse_der(_to_array(sig, "hex"))
        if rs is None:
            raise EllipticError("Signature without r or s")
        return rs

    def _public(self, key, enc=None):
        """KeyPair._importPublic (ec/key.js:84-99) -> (fmt, bytes)."""
        ln = self._len
        if isinstance(key, dict) and (key.get("x") or key.get("y")):
            if not (key.get("x") and key.get("y")):
                raise EllipticError("Need both x and y coordinate")
            x, y = _bn(key["x"]), _bn(key["y"])
            if x < 0 or y < 0:
                raise EllipticError("red works only with positives")
            # toRed reduces oversize coordinates mod p (short.js:258-268); the engine
            # does that for anything that fits the wire width, wider values here.
            if x >> (8 * ln) or y >> (8 * ln):
                x %= self._c["p"]
                y %= self._c["p"]
            return nat.PUB_XY, x.to_bytes(ln, "big") + y.to_bytes(ln, "big")
        b = _to_array(key, enc)
        if len(b) and b[0] in (4, 6, 7) and len(b) - 1 == 2 * ln:
            if (b[0] == 6 and b[-1] % 2 != 0) or (b[0] == 7 and b[-1] % 2 != 1):
                raise EllipticError("Assertion failed")            # base.js:278-281
            return nat.PUB_XY, b[1:]
        if len(b) and b[0] in (2, 3) and len(b) - 1 == ln:
            return nat.PUB_SEC1_33, b
        raise EllipticError("Unknown point format")                # base.js:291

    # ---- batch entry points ---------------------------------------------------
    def verify_batch_packed(self, e, r, s, pub, pub_fmt=nat.PUB_XY):
        """Packed form: e, r, s are (n, len) uint8 arrays (big-endian), pub is
        (n, 2*len).  Returns the per-item status bytes (see _native.ST_*)."""
        lib = nat.init(self._device)
        e = np.ascontiguousarray(e, dtype=np.uint8)
        r = np.ascontiguousarray(r, dtype=np.uint8)
        s = np.ascontiguousarray(s, dtype=np.uint8)
        pub = np.ascontiguousarray(pub, dtype=np.uint8)
        n = e.shape[0]
        if not (e.shape == (n, self._len) and r.shape == e.shape and s.shape == e.shape):      # raw pointers go down
            raise ValueError("e, r, s must be (n, %d) uint8 arrays" % self._len)
        pb = {nat.PUB_XY: 2 * self._len, nat.PUB_SEC1_65: 1 + 2 * self._len, nat.PUB_SEC1_33: 1 + self._len}[pub_fmt]
        if pub.shape != (n, pb):
            raise ValueError("pub must be (n, %d) for this format, got %r" % (pb, pub.shape))
        status = np.empty(n, dtype=np.uint8)
        nat.check(lib.eb200_ecdsa_verify_batch(
            self._c["id"], n, e.ctypes.data, r.ctypes.data, s.ctypes.data, pub.ctypes.data,
            pub_fmt, status.ctypes.data))
        return status

    def verify_batch_der_packed(self, e, ders, pub, pub_fmt=nat.PUB_XY):
        """e: (n, len) uint8 truncated hashes; ders: list of DER byte strings (parsed on the GPU exactly as
        Signature._importDER, ec/signature.js:73-134); pub: (n, k) uint8 in `pub_fmt`.  Returns statuses."""
        lib = nat.init(self._device)
        n = len(ders)
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum([len(d) for d in ders])
        blob = np.frombuffer(b"".join(bytes(d) for d in ders) + b"\x00", np.uint8)
        e = np.ascontiguousarray(e, np.uint8); pub = np.ascontiguousarray(pub, np.uint8)
        pb = {nat.PUB_XY: 2 * self._len, nat.PUB_SEC1_65: 1 + 2 * self._len, nat.PUB_SEC1_33: 1 + self._len}[pub_fmt]
        if e.shape != (n, self._len) or pub.shape != (n, pb):
            raise EllipticError("verify_batch_der_packed: e must be (n, %d) and pub (n, %d)" % (self._len, pb))
        status = np.zeros(n, np.uint8)
        nat.check(lib.eb200_ecdsa_verify_batch_der(self._c["id"], n, e.ctypes.data, blob.ctypes.data, off.ctypes.data,
                                                   pub.ctypes.data, pub_fmt, status.ctypes.data))
        return status

    def verify_batch(self, msgs, sigs, keys, enc=None, msg_bit_length=None):
        """EC#verifyBatch: lists of the reference's own argument forms.
        Returns a uint8 array of statuses; items whose *parsing* throws in the
        reference raise here, like a loop over `verify` would at that item."""
        n = len(msgs)
        ln = self._len
        e = np.zeros((n, ln), np.uint8)
        r = np.zeros((n, ln), np.uint8)
        s = np.zeros((n, ln), np.uint8)
        groups = {}          # pub format -> (item indices, key bytes)
        early = {}
        for i in range(n):
            ev = self._truncate_to_n(msgs[i], msg_bit_length)
            fmt, pb = self._public(keys[i], enc)
            rv, sv = self._signature(sigs[i])
            if rv < 1 or rv >= self.n or sv < 1 or sv >= self.n:
                early[i] = nat.ST_FALSE       # ec/index.js:199-202 (after the key import, which may throw)
                rv = sv = 0
            e[i] = np.frombuffer(ev.to_bytes(ln, "big"), np.uint8)
            r[i] = np.frombuffer(rv.to_bytes(ln, "big"), np.uint8)
            s[i] = np.frombuffer(sv.to_bytes(ln, "big"), np.uint8)
            g = groups.setdefault(fmt, ([], []))
            g[0].append(i)
            g[1].append(pb)
        st = np.zeros(n, np.uint8)
        for fmt, (idx, pbs) in groups.items():
            idx = np.asarray(idx)
            pub = np.frombuffer(b"".join(pbs), np.uint8).reshape(len(pbs), -1)
            st[idx] = self.verify_batch_packed(e[idx], r[idx], s[idx], pub, fmt)
        for i, v in early.items():
            if st[i] in (nat.ST_TRUE, nat.ST_FALSE, nat.ST_NEEDS_HOST):
                st[i] = v
        return st

    # ---- signing ---------------------------------------------------------------------------------------------
    def sign_batch(self, msgs, privs, canonical=False, msg_bit_length=None, pers=None, pers_enc=None, k=None):
        """Batch of EC.prototype.sign (ec/index.js:110-186) with the curve's default hash.
        pers / pers_enc: the `pers` / `persEnc` options (one personalisation string for the batch; persEnc defaults
        to 'utf8' as in ec/index.js:150).  k: the `k` option as a callable k(item, iter) -> nonce (int / hex / bytes);
        without it the nonces are RFC 6979 (HMAC-DRBG on the GPU).  Returns (r list, s list, recoveryParam array)."""
        if self.name not in _XY:
            raise EllipticError("sign_batch: not available on " + self.name)
        lib = nat.init(self._device)
        n, ln = len(msgs), self._len
        e = np.zeros((n, ln), np.uint8); d = np.zeros((n, ln), np.uint8)
        for i in range(n):
            ev = self._truncate_to_n(msgs[i], msg_bit_length)                  # includes the single `- n` (ec/index.js:105-106)
            if ev >> (8 * ln):
                raise EllipticError("byte array longer than desired length")   # msg.toArray('be', bytes), dist bn.js toArrayLike
            e[i] = np.frombuffer(ev.to_bytes(ln, "big"), np.uint8)
            d[i] = np.frombuffer((_bn(privs[i]) % self.n).to_bytes(ln, "big"), np.uint8)   # _importPrivate
        r = np.zeros((n, ln), np.uint8); s = np.zeros((n, ln), np.uint8)
        rec = np.zeros(n, np.uint8); st = np.zeros(n, np.uint8)
        flags = 1 if canonical else 0
        if k is not None:
            todo = np.arange(n)
            for it in range(1 << 16):
                kb = np.zeros((len(todo), ln), np.uint8)
                for j, i in enumerate(todo):
                    kv = _bn(k(int(i), it))
                    # _truncateToN(k, true) on a BN wider than the field: shift by its own byte length (ec/index.js:96-103)
                    delta = ((kv.bit_length() + 7) // 8) * 8 - self.n.bit_length()
                    if kv.bit_length() > 8 * ln and delta > 0:
                        kv >>= delta
                    kb[j] = np.frombuffer(kv.to_bytes(ln, "big"), np.uint8)
                es, ds = np.ascontiguousarray(e[todo]), np.ascontiguousarray(d[todo])
                rr = np.zeros((len(todo), ln), np.uint8); ss = np.zeros((len(todo), ln), np.uint8)
                cc = np.zeros(len(todo), np.uint8); tt = np.zeros(len(todo), np.uint8)
                nat.check(lib.eb200_ecdsa_sign_batch_k(self._c["id"], len(todo), es.ctypes.data, ds.ctypes.data, kb.ctypes.data, flags,
                                                       rr.ctypes.data, ss.ctypes.data, cc.ctypes.data, tt.ctypes.data))
                ok = tt == nat.ST_TRUE
                r[todo[ok]], s[todo[ok]], rec[todo[ok]], st[todo[ok]] = rr[ok], ss[ok], cc[ok], tt[ok]
                if not bool(((tt == nat.ST_TRUE) | (tt == nat.ST_RETRY)).all()):
                    raise nat.NativeError("sign_batch: unexpected status")
                todo = todo[~ok]
                if not len(todo):
                    break
        elif pers is not None:
            pb = np.frombuffer(bytes(_to_array(pers, pers_enc or "utf8")) + b"\x00", np.uint8)
            nat.check(lib.eb200_ecdsa_sign_batch_pers(self._c["id"], n, e.ctypes.data, d.ctypes.data, pb.ctypes.data, pb.size - 1, flags,
                                                      r.ctypes.data, s.ctypes.data, rec.ctypes.data, st.ctypes.data))
        else:
            nat.check(lib.eb200_ecdsa_sign_batch(self._c["id"], n, e.ctypes.data, d.ctypes.data, flags,
                                                 r.ctypes.data, s.ctypes.data, rec.ctypes.data, st.ctypes.data))
        if not bool((st == nat.ST_TRUE).all()):
            bad = int(np.flatnonzero(st != nat.ST_TRUE)[0])
            raise nat.NativeError("sign_batch: item %d returned status %d" % (bad, int(st[bad])))
        return ([int.from_bytes(r[i].tobytes(), "big") for i in range(n)],
                [int.from_bytes(s[i].tobytes(), "big") for i in range(n)], rec)

    def gen_key_pair_batch(self, entropies, entropy_enc=None, pers=None, pers_enc=None):
        """Batch of EC.prototype.genKeyPair({entropy, entropyEnc, pers, persEnc}) (ec/index.js:55-79): every item's key
        comes from its own HMAC-DRBG(hash, entropy, nonce = n, pers).  All entropies must have the same byte length
        (>= 24, the reference's 'Not enough entropy' assertion).  Returns (private keys, public points (x, y))."""
        if self.name not in _XY:
            raise EllipticError("gen_key_pair_batch: not available on " + self.name)
        lib = nat.init(self._device)
        ents = [bytes(_to_array(x, entropy_enc or "utf8")) for x in entropies]
        n, ln = len(ents), self._len
        if not n:
            return [], []
        ne = len(ents[0])
        if any(len(x) < 24 for x in ents):
            raise EllipticError("Not enough entropy. Minimum is: 192 bits")     # hmac-drbg ctor, dist:8708-8710
        if any(len(x) != ne for x in ents):
            raise EllipticError("gen_key_pair_batch: entropies of one batch must have the same length")
        eb = np.frombuffer(b"".join(ents), np.uint8).reshape(n, ne)
        pb = np.frombuffer(bytes(_to_array(pers, pers_enc or "utf8")) + b"\x00", np.uint8) if pers is not None else None
        priv = np.zeros((n, ln), np.uint8); pub = np.zeros((n, 2 * ln), np.uint8); st = np.zeros(n, np.uint8)
        nat.check(lib.eb200_ec_keygen_batch(self._c["id"], n, eb.ctypes.data, ne, pb.ctypes.data if pb is not None else None,
                                            (pb.size - 1) if pb is not None else 0, priv.ctypes.data, pub.ctypes.data, st.ctypes.data))
        if not bool((st == nat.ST_TRUE).all()):
            raise nat.NativeError("gen_key_pair_batch: unexpected status")
        return ([int.from_bytes(priv[i].tobytes(), "big") for i in range(n)],
                [(int.from_bytes(pub[i, :ln].tobytes(), "big"), int.from_bytes(pub[i, ln:].tobytes(), "big")) for i in range(n)])

    def sign(self, msg, priv, canonical=False, pers=None, pers_enc=None, k=None):
        """EC.prototype.sign (ec/index.js:110-186); k: the reference's options.k(iter)."""
        r, s, rec = self.sign_batch([msg], [priv], canonical, None, pers, pers_enc, (lambda i, it: k(it)) if k else None)
        return {"r": r[0], "s": s[0], "recoveryParam": int(rec[0])}

    # ---- public-key recovery -----------------------------------------------------------------------------
    def recover_pub_key_batch(self, msgs, sigs, js, enc=None):
        """Batch of EC.prototype.recoverPubKey (ec/index.js:231-259).  msgs as `new BN(msg)` takes them
        (int / hex / bytes, NOT truncated), sigs as Signature takes them, js the recovery params.
        Returns (points, statuses): points[i] = (x, y), None for the point at infinity / a throw."""
        if self.name not in _SHORT:
            raise EllipticError("recover_pub_key_batch: short curves only")
        lib = nat.init(self._device)
        n, ln = len(msgs), self._len
        e = np.zeros((n, ln), np.uint8); r = np.zeros((n, ln), np.uint8); s = np.zeros((n, ln), np.uint8)
        rid = np.zeros(n, np.uint8)
        for i in range(n):
            if (3 & js[i]) != js[i]:
                raise EllipticError("The recovery param is more than two bits")      # ec/index.js:232
            rv, sv = self._signature_enc(sigs[i], enc)
            ev = _bn(msgs[i]) if not isinstance(msgs[i], (bytes, bytearray, list, tuple)) else int.from_bytes(_to_array(msgs[i]), "big")
            if rv >> (8 * ln):
                raise NeedsReferencePath("r does not fit the curve's field width")
            e[i] = np.frombuffer((ev % self.n).to_bytes(ln, "big"), np.uint8)
            r[i] = np.frombuffer(rv.to_bytes(ln, "big"), np.uint8)
            s[i] = np.frombuffer((sv % self.n).to_bytes(ln, "big"), np.uint8)
            rid[i] = js[i]
        out = np.zeros((n, 2 * ln), np.uint8)
        st = np.zeros(n, np.uint8)
        nat.check(lib.eb200_ecdsa_recover_batch(self._c["id"], n, e.ctypes.data, r.ctypes.data, s.ctypes.data,
                                                rid.ctypes.data, out.ctypes.data, st.ctypes.data))
        pts = [(int.from_bytes(out[i, :ln].tobytes(), "big"), int.from_bytes(out[i, ln:].tobytes(), "big"))
               if st[i] == nat.ST_TRUE else None for i in range(n)]
        return pts, st

    def recover_pub_key(self, msg, signature, j, enc=None):
        pts, st = self.recover_pub_key_batch([msg], [signature], [j], enc)
        if st[0] in (nat.ST_TRUE, nat.ST_INFINITY):
            return pts[0]
        raise EllipticError(_THROW_MSG.get(int(st[0]), "status %d" % int(st[0])))

    # ---- curve.point(...).mul / mulAdd batches (short.js:422-441) ---------------------------------------------
    def _scalars(self, ks):
        ln = self._len
        out = np.zeros((len(ks), ln), np.uint8)
        for i, k in enumerate(ks):
            k = _bn(k)
            if k < 0:
                raise EllipticError("negative scalars are not supported by the batch path")
            if k >> (8 * ln):
                k %= self.n          # same point for every on-curve input
            out[i] = np.frombuffer(k.to_bytes(ln, "big"), np.uint8)
        return out

    def _points(self, pts):
        """curve.point(x, y) (short.js:251-271): coordinates reduced mod p, not validated."""
        ln = self._len
        out = np.zeros((len(pts), 2 * ln), np.uint8)
        for i, pt in enumerate(pts):
            x, y = (pt["x"], pt["y"]) if isinstance(pt, dict) else pt
            out[i, :ln] = np.frombuffer((_bn(x) % self._c["p"]).to_bytes(ln, "big"), np.uint8)
            out[i, ln:] = np.frombuffer((_bn(y) % self._c["p"]).to_bytes(ln, "big"), np.uint8)
        return out

    def _mul_common(self, k1, k2, pts):
        if self.name == "curve25519":
            raise EllipticError("Not supported on Montgomery curve")      # mont.js: mulAdd / jumlAdd throw; mul: x_mul_batch
        if self.name not in _XY:
            raise EllipticError("mul/mulAdd batches: not available on " + self.name)
        lib = nat.init(self._device)
        n, ln = len(k2), self._len
        out = np.zeros((n, 2 * ln), np.uint8)
        st = np.zeros(n, np.uint8)
        if k1 is None:
            nat.check(lib.eb200_scalar_mul_batch(self._c["id"], n, k2.ctypes.data, pts.ctypes.data if pts is not None else None,
                                                 out.ctypes.data, st.ctypes.data))
        else:
            nat.check(lib.eb200_mul_add_batch(self._c["id"], n, k1.ctypes.data, k2.ctypes.data, pts.ctypes.data,
                                              out.ctypes.data, st.ctypes.data))
        if bool((st == nat.ST_NEEDS_HOST).any()):      # (only the Edwards preset reports it; the short curves replay)
            raise NeedsReferencePath("point %d is not on the curve; the reference does not validate it" % int(np.flatnonzero(st == nat.ST_NEEDS_HOST)[0]))
        return [(int.from_bytes(out[i, :ln].tobytes(), "big"), int.from_bytes(out[i, ln:].tobytes(), "big"))
                if st[i] == nat.ST_TRUE else None for i in range(n)]

    def x_mul_batch(self, xs, ks):
        """curve25519: [curve.point(x).mul(k).getX()] (mont.js:130-153, 173-178) -- x-only points, no validation."""
        if self.name != "curve25519":
            raise EllipticError("x_mul_batch: curve25519 only")
        lib = nat.init(self._device)
        n = len(ks)
        kb = np.zeros((n, 32), np.uint8); xb = np.zeros((n, 32), np.uint8)
        for i in range(n):
            kv, xv = _bn(ks[i]), _bn(xs[i])
            if kv >> 256:
                raise NeedsReferencePath("scalar wider than 256 bits")
            kb[i] = np.frombuffer(kv.to_bytes(32, "big"), np.uint8)
            xb[i] = np.frombuffer((xv % self._c["p"] if xv >> 256 else xv).to_bytes(32, "big"), np.uint8)
        out = np.zeros((n, 32), np.uint8); st = np.zeros(n, np.uint8)
        nat.check(lib.eb200_x25519_mul_batch(n, kb.ctypes.data, xb.ctypes.data, out.ctypes.data, st.ctypes.data))
        return [int.from_bytes(out[i].tobytes(), "big") for i in range(n)]

    def g_mul_batch(self, ks):
        """[G.mul(k) for k in ks] (short.js:422-427, the keygen product ec/key.js:55-60): (x, y) or None = infinity."""
        return self._mul_common(None, self._scalars(ks), None)

    def mul_batch(self, points, ks):
        """[curve.point(x, y).mul(k)] (short.js:422-432): (x, y) or None = infinity."""
        return self._mul_common(None, self._scalars(ks), self._points(points))

    def mul_add_batch(self, k1s, p2s, k2s):
        """[G.mulAdd(k1, P2, k2)] (short.js:434-441): (x, y) or None = infinity."""
        return self._mul_common(self._scalars(k1s), self._scalars(k2s), self._points(p2s))

    def _signature_enc(self, sig, enc):
        """new Signature(sig, enc) as recoverPubKey calls it (enc passed through, ec/index.js:233)."""
        if isinstance(sig, dict) or (hasattr(sig, "r") and hasattr(sig, "s")):
            return self._signature(sig)
        rs = parse_der(_to_array(sig, enc))
        if rs is None:
            raise EllipticError("Signature without r or s")
        return rs

    # ---- ECDH --------------------------------------------------------------------------------------------
    def derive_batch(self, privs, pubs):
        """Batch of `ec.keyFromPrivate(priv).derive(ec.keyFromPublic(pub).getPublic())`
        (ec/key.js:76-82, 102-107) on curve25519.  privs: ints / hex / bytes; pubs: the peer's x as
        int / hex / big-endian bytes (mont.js:46-48).  Returns (list of int-or-None, statuses)."""
        if self.name != "curve25519":
            return self._derive_short(privs, pubs)
        lib = nat.init(self._device)
        n = len(privs)
        k = np.zeros((n, 32), np.uint8)
        x = np.zeros((n, 32), np.uint8)
        for i in range(n):
            kv = _bn(privs[i]) % self.n                      # _importPrivate: umod n
            xv = _bn(pubs[i])
            if xv >> 256:
                xv %= self._c["p"]
            k[i] = np.frombuffer(kv.to_bytes(32, "big"), np.uint8)
            x[i] = np.frombuffer(xv.to_bytes(32, "big"), np.uint8)
        out = np.zeros((n, 32), np.uint8)
        st = np.zeros(n, np.uint8)
        nat.check(lib.eb200_x25519_derive_batch(n, k.ctypes.data, x.ctypes.data, out.ctypes.data, st.ctypes.data))
        vals = [int.from_bytes(out[i].tobytes(), "big") if st[i] == nat.ST_TRUE else None for i in range(n)]
        return vals, st

    def derive_batch_packed(self, priv, pubx, out=None, status=None):
        """Packed curve25519 ECDH: priv, pubx are (n, 32) uint8 arrays (big-endian; priv already reduced mod n as
        _importPrivate does, ec/key.js:76-82).  Returns ((n, 32) shared x big-endian, statuses); `out` / `status`
        let the caller supply (and reuse) the result buffers, e.g. pinned ones."""
        if self.name != "curve25519":
            raise EllipticError("derive_batch_packed: curve25519 only (short curves: derive_batch)")
        lib = nat.init(self._device)
        priv = np.ascontiguousarray(priv, dtype=np.uint8)
        pubx = np.ascontiguousarray(pubx, dtype=np.uint8)
        n = priv.shape[0]
        if priv.shape != (n, 32) or pubx.shape != (n, 32):
            raise EllipticError("derive_batch_packed: (n, 32) arrays expected")
        out = np.empty((n, 32), np.uint8) if out is None else out
        st = np.empty(n, np.uint8) if status is None else status
        if out.shape != (n, 32) or out.dtype != np.uint8 or not out.flags.c_contiguous or st.shape != (n,) or st.dtype != np.uint8:
            raise EllipticError("derive_batch_packed: out must be a contiguous (n, 32) uint8 array, status (n,) uint8")
        nat.check(lib.eb200_x25519_derive_batch(n, priv.ctypes.data, pubx.ctypes.data, out.ctypes.data, st.ctypes.data))
        return out, st

    def _derive_short(self, privs, pubs):
        """Short curves: pubs are the peer's points as {x, y} / (x, y) (keyFromPublic(...).getPublic())."""
        lib = nat.init(self._device)
        n, ln = len(privs), self._len
        k = self._scalars([_bn(p) % self.n for p in privs])
        pts = self._points(pubs)
        out = np.zeros((n, ln), np.uint8)
        st = np.zeros(n, np.uint8)
        nat.check(lib.eb200_ecdh_derive_batch(self._c["id"], n, k.ctypes.data, pts.ctypes.data, out.ctypes.data, st.ctypes.data))
        vals = [int.from_bytes(out[i].tobytes(), "big") if st[i] == nat.ST_TRUE else None for i in range(n)]
        return vals, st

    def derive(self, priv, pub):
        """KeyPair.prototype.derive: the shared x as an int, or raises like the reference."""
        vals, st = self.derive_batch([priv], [pub])
        if st[0] == nat.ST_TRUE:
            return vals[0]
        raise EllipticError(_THROW_MSG.get(int(st[0]), "status %d" % int(st[0])))

    def verify(self, msg, signature, key, enc=None, options=None):
        """EC.prototype.verify (ec/index.js:188-229): bool, or raises."""
        mbl = (options or {}).get("msgBitLength")
        st = int(self.verify_batch([msg], [signature], [key], enc, mbl)[0])
        if st == nat.ST_TRUE:
            return True
        if st == nat.ST_FALSE:
            return False
        if st == nat.ST_NEEDS_HOST:
            raise NeedsReferencePath("public key is not on the curve; the reference does not validate it")
        raise EllipticError(_THROW_MSG.get(st, "status %d" % st))
