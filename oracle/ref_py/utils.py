"""ORACLE (test infrastructure only).  Restates lib/elliptic/utils.js."""


def get_naf(num, w, bits):
    """utils.getNAF, lib/elliptic/utils.js:15-44.

    NB: this is not a textbook width-(w+1) NAF; digits satisfy
    |z| <= 2^w - 1 and sum(z_i 2^i) == num, which is all callers rely on.
    """
    naf = [0] * (max(num.bit_length(), bits) + 1)
    ws = 1 << (w + 1)
    k = num
    for i in range(len(naf)):
        mod = k & (ws - 1)
        if k & 1:
            if mod > (ws >> 1) - 1:
                z = (ws >> 1) - mod
            else:
                z = mod
            k -= z
        else:
            z = 0
        naf[i] = z
        k >>= 1
    return naf


def get_jsf(k1, k2):
    """utils.getJSF, lib/elliptic/utils.js:47-101 (joint sparse form)."""
    jsf = [[], []]
    d1 = 0
    d2 = 0
    while k1 > -d1 or k2 > -d2:
        m14 = ((k1 & 3) + d1) & 3
        m24 = ((k2 & 3) + d2) & 3
        if m14 == 3:
            m14 = -1
        if m24 == 3:
            m24 = -1
        if (m14 & 1) == 0:
            u1 = 0
        else:
            m8 = ((k1 & 7) + d1) & 7
            if (m8 == 3 or m8 == 5) and m24 == 2:
                u1 = -m14
            else:
                u1 = m14
        jsf[0].append(u1)
        if (m24 & 1) == 0:
            u2 = 0
        else:
            m8 = ((k2 & 7) + d2) & 7
            if (m8 == 3 or m8 == 5) and m14 == 2:
                u2 = -m24
            else:
                u2 = m24
        jsf[1].append(u2)
        if 2 * d1 == u1 + 1:
            d1 = 1 - d1
        if 2 * d2 == u2 + 1:
            d2 = 1 - d2
        k1 >>= 1
        k2 >>= 1
    return jsf


def to_array(msg, enc=None):
    """minimalistic-crypto-utils toArray, dist/elliptic.js:8847-8876.

    Arrays pass through (each element `| 0`); 'hex' strings are stripped of
    non-hex characters, left-padded to even length and parsed pairwise; other
    strings are taken per UTF-16 code unit (hi byte emitted only if non-zero).
    """
    if isinstance(msg, (bytes, bytearray, list, tuple)):
        return [int(b) for b in msg]
    if not msg:
        return []
    res = []
    if isinstance(msg, str):
        if enc == "hex":
            import re
            msg = re.sub(r"[^a-zA-Z0-9]+", "", msg)
            if len(msg) % 2 != 0:
                msg = "0" + msg
            for i in range(0, len(msg), 2):
                try:
                    res.append(int(msg[i:i + 2], 16))
                except ValueError:
                    res.append(0)  # parseInt -> NaN -> |0 at use sites
        else:
            for ch in msg:
                c = ord(ch)
                hi = c >> 8
                lo = c & 0xFF
                if hi:
                    res.extend([hi, lo])
                else:
                    res.append(lo)
    return res


def int_from_le(b):
    """utils.intFromLE, lib/elliptic/utils.js:118-121."""
    return int.from_bytes(bytes(b), "little")
