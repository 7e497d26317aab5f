"""ORACLE (test infrastructure only).  Restates lib/elliptic/ec/signature.js."""

from .bn import ref_assert
from .utils import to_array


class Signature:
    """ec/signature.js:8-22."""

    def __init__(self, options, enc=None):
        if isinstance(options, Signature):
            self.r, self.s, self.recovery_param = options.r, options.s, options.recovery_param
            return
        if self._import_der(options, enc):
            return
        ok = isinstance(options, dict) and options.get("r") and options.get("s")
        ref_assert(ok, "Signature without r or s")
        self.r = _bn(options["r"])
        self.s = _bn(options["s"])
        self.recovery_param = options.get("recoveryParam")

    def _import_der(self, data, enc):
        """ec/signature.js:73-134 (strict DER)."""
        if isinstance(data, dict):
            return False  # toArray({}) -> [] -> data[0] !== 0x30
        data = to_array(data, enc)
        g = lambda i: data[i] if 0 <= i < len(data) else None
        place = [0]

        def get_length():
            initial = g(place[0]); place[0] += 1
            if initial is None:
                return 0  # undefined & 0x80 == 0 -> returns undefined; treated as falsy below
            if not (initial & 0x80):
                return initial
            octet_len = initial & 0xF
            if octet_len == 0 or octet_len > 4:
                return False
            if g(place[0]) == 0x00:
                return False
            val = 0
            off = place[0]
            for _ in range(octet_len):
                val = ((val << 8) | (g(off) or 0)) & 0xFFFFFFFF
                off += 1
            if val <= 0x7F:
                return False
            place[0] = off
            return val

        first = g(place[0]); place[0] += 1
        if first != 0x30:
            return False
        ln = get_length()
        if ln is False:
            return False
        if ln + place[0] != len(data):
            return False
        t = g(place[0]); place[0] += 1
        if t != 0x02:
            return False
        rlen = get_length()
        if rlen is False:
            return False
        if ((g(place[0]) or 0) & 128) != 0:
            return False
        r = data[place[0]:rlen + place[0]]
        place[0] += rlen
        t = g(place[0]); place[0] += 1
        if t != 0x02:
            return False
        slen = get_length()
        if slen is False:
            return False
        if len(data) != slen + place[0]:
            return False
        if ((g(place[0]) or 0) & 128) != 0:
            return False
        s = data[place[0]:slen + place[0]]
        if len(r) and r[0] == 0:
            if len(r) > 1 and r[1] & 0x80:
                r = r[1:]
            else:
                return False
        if len(s) and s[0] == 0:
            if len(s) > 1 and s[1] & 0x80:
                s = s[1:]
            else:
                return False
        self.r = int.from_bytes(bytes(r), "big")
        self.s = int.from_bytes(bytes(s), "big")
        self.recovery_param = None
        return True

    def to_der(self):
        """ec/signature.js:149-176."""
        def arr(v):
            b = list(v.to_bytes(max(1, (v.bit_length() + 7) // 8), "big"))
            if b[0] & 0x80:
                b = [0] + b
            i = 0
            while i < len(b) - 1 and not b[i] and not (b[i + 1] & 0x80):
                i += 1
            return b[i:]

        def clen(out, n):
            if n < 0x80:
                out.append(n)
                return
            octets = 1 + ((n.bit_length() - 1) >> 3)
            out.append(octets | 0x80)
            for o in range(octets - 1, 0, -1):
                out.append((n >> (o << 3)) & 0xFF)
            out.append(n & 0xFF)

        r, s = arr(self.r), arr(self.s)
        out = [0x02]
        clen(out, len(r)); out += r
        out.append(0x02)
        clen(out, len(s)); out += s
        res = [0x30]
        clen(res, len(out))
        return bytes(res + out)


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
    """new BN(v, 16) for the forms callers use: int (BN), hex str, byte array."""
    if isinstance(v, int):
        return v
    if isinstance(v, str):
        return _parse_hex(v)
    return int.from_bytes(bytes(v), "big")
