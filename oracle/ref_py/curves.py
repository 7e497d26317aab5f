"""ORACLE (test infrastructure only).  Restates lib/elliptic/curves.js presets.

Curve parameters are public constants (SECG / NIST / RFC 7748 / RFC 8032) as
listed at curves.js:43-206.  The secp256k1 G tables that the reference loads
from lib/elliptic/precomputed/secp256k1.js (doubles step 4, naf wnd 7) are
regenerated here from G; tests/test_oracle_golden.py checks them against a
digest of the reference's file.
"""

import hashlib

from .short import ShortCurve
from .bn import ref_assert

H = lambda s: int(s.replace(" ", ""), 16)

_CONF = {
    "p192": dict(type="short", p=H("ffffffff ffffffff ffffffff fffffffe ffffffff ffffffff"),
                 a=H("ffffffff ffffffff ffffffff fffffffe ffffffff fffffffc"),
                 b=H("64210519 e59c80e7 0fa7e9ab 72243049 feb8deec c146b9b1"),
                 n=H("ffffffff ffffffff ffffffff 99def836 146bc9b1 b4d22831"), hash="sha256",
                 g=(H("188da80e b03090f6 7cbf20eb 43a18800 f4ff0afd 82ff1012"),
                    H("07192b95 ffc8da78 631011ed 6b24cdd5 73f977a1 1e794811"))),
    "p224": dict(type="short", p=H("ffffffff ffffffff ffffffff ffffffff 00000000 00000000 00000001"),
                 a=H("ffffffff ffffffff ffffffff fffffffe ffffffff ffffffff fffffffe"),
                 b=H("b4050a85 0c04b3ab f5413256 5044b0b7 d7bfd8ba 270b3943 2355ffb4"),
                 n=H("ffffffff ffffffff ffffffff ffff16a2 e0b8f03e 13dd2945 5c5c2a3d"), hash="sha256",
                 g=(H("b70e0cbd 6bb4bf7f 321390b9 4a03c1d3 56c21122 343280d6 115c1d21"),
                    H("bd376388 b5f723fb 4c22dfe6 cd4375a0 5a074764 44d58199 85007e34"))),
    "p256": dict(type="short", p=H("ffffffff 00000001 00000000 00000000 00000000 ffffffff ffffffff ffffffff"),
                 a=H("ffffffff 00000001 00000000 00000000 00000000 ffffffff ffffffff fffffffc"),
                 b=H("5ac635d8 aa3a93e7 b3ebbd55 769886bc 651d06b0 cc53b0f6 3bce3c3e 27d2604b"),
                 n=H("ffffffff 00000000 ffffffff ffffffff bce6faad a7179e84 f3b9cac2 fc632551"), hash="sha256",
                 g=(H("6b17d1f2 e12c4247 f8bce6e5 63a440f2 77037d81 2deb33a0 f4a13945 d898c296"),
                    H("4fe342e2 fe1a7f9b 8ee7eb4a 7c0f9e16 2bce3357 6b315ece cbb64068 37bf51f5"))),
    "p384": dict(type="short",
                 p=H("ffffffff ffffffff ffffffff ffffffff ffffffff ffffffff ffffffff fffffffe ffffffff 00000000 00000000 ffffffff"),
                 a=H("ffffffff ffffffff ffffffff ffffffff ffffffff ffffffff ffffffff fffffffe ffffffff 00000000 00000000 fffffffc"),
                 b=H("b3312fa7 e23ee7e4 988e056b e3f82d19 181d9c6e fe814112 0314088f 5013875a c656398d 8a2ed19d 2a85c8ed d3ec2aef"),
                 n=H("ffffffff ffffffff ffffffff ffffffff ffffffff ffffffff c7634d81 f4372ddf 581a0db2 48b0a77a ecec196a ccc52973"),
                 hash="sha384",
                 g=(H("aa87ca22 be8b0537 8eb1c71e f320ad74 6e1d3b62 8ba79b98 59f741e0 82542a38 5502f25d bf55296c 3a545e38 72760ab7"),
                    H("3617de4a 96262c6f 5d9e98bf 9292dc29 f8f41dbd 289a147c e9da3113 b5f0b8c0 0a60b1ce 1d7e819d 7a431d7c 90ea0e5f"))),
    "p521": dict(type="short",
                 p=(1 << 521) - 1, a=(1 << 521) - 4,
                 b=H("00000051 953eb961 8e1c9a1f 929a21a0 b68540ee a2da725b 99b315f3 b8b48991 8ef109e1 56193951 ec7e937b 1652c0bd 3bb1bf07 3573df88 3d2c34f1 ef451fd4 6b503f00"),
                 n=H("000001ff ffffffff ffffffff ffffffff ffffffff ffffffff ffffffff ffffffff fffffffa 51868783 bf2f966b 7fcc0148 f709a5d0 3bb5c9b8 899c47ae bb6fb71e 91386409"),
                 hash="sha512",
                 g=(H("000000c6 858e06b7 0404e9cd 9e3ecb66 2395b442 9c648139 053fb521 f828af60 6b4d3dba a14b5e77 efe75928 fe1dc127 a2ffa8de 3348b3c1 856a429b f97e7e31 c2e5bd66"),
                    H("00000118 39296a78 9a3bc004 5c8a5fb4 2c7d1bd9 98f54449 579b4468 17afbd17 273e662c 97ee7299 5ef42640 c550b901 3fad0761 353c7086 a272c240 88be9476 9fd16650"))),
    "curve25519": dict(type="mont", p=(1 << 255) - 19, a=0x76d06, b=1,
                       n=H("1000000000000000 0000000000000000 14def9dea2f79cd6 5812631a5cf5d3ed"),
                       hash="sha256", g=(9,)),
    "ed25519": dict(type="edwards", p=(1 << 255) - 19, a=-1, c=1,
                    d=H("52036cee2b6ffe73 8cc740797779e898 00700a4d4141d8ab 75eb4dca135978a3"),
                    n=H("1000000000000000 0000000000000000 14def9dea2f79cd6 5812631a5cf5d3ed"),
                    hash="sha256",
                    g=(H("216936d3cd6e53fec0a4e231fdd6dc5c692cc7609525a7b2c9562d608f25d51a"),
                       H("6666666666666666666666666666666666666666666666666666666666666658"))),
    "secp256k1": dict(type="short", p=H("ffffffff ffffffff ffffffff ffffffff ffffffff ffffffff fffffffe fffffc2f"),
                      a=0, b=7,
                      n=H("ffffffff ffffffff ffffffff fffffffe baaedce6 af48a03b bfd25e8c d0364141"),
                      hash="sha256",
                      beta=H("7ae96a2b657c07106e64479eac3434e99cf0497512f58995c1396c28719501ee"),
                      **{"lambda": H("5363ad4cc05c30e0a5261c028812645a122e22ea20816678df02967c1b23bd72")},
                      basis=[dict(a=H("3086d221a7d46bcde86c90e49284eb15"), b=-H("e4437ed6010e88286f547fa90abfe4c3")),
                             dict(a=H("114ca50f7a8e2f3f657c1108d9d44cfd8"), b=H("3086d221a7d46bcde86c90e49284eb15"))],
                      g=(H("79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798"),
                         H("483ada7726a3c4655da4fbfc0e1108a8fd17b448a68554199c47d08ffb10d4b8"))),
}

HASHES = {"sha256": hashlib.sha256, "sha384": hashlib.sha384, "sha512": hashlib.sha512,
          "sha224": hashlib.sha224, "sha1": hashlib.sha1}


def _secp256k1_pre(conf):
    """Regenerate lib/elliptic/precomputed/secp256k1.js: doubles = 2^(4i) G for
    i = 1..65 (`:2-266`), naf = (2i+1) G for i = 1..127 (`:267-779`)."""
    tmp = ShortCurve(dict(conf, g=None))
    g = tmp.point(*conf["g"])
    doubles = []
    acc = g
    for _ in range(65):
        for _ in range(4):
            acc = acc.dbl()
        doubles.append((acc.x, acc.y))
    d = g.dbl()
    naf = []
    acc = g
    for _ in range(127):
        acc = acc.add(d)
        naf.append((acc.x, acc.y))
    return {"doubles": {"step": 4, "points": doubles}, "naf": {"wnd": 7, "points": naf}}


class PresetCurve:
    """curves.js:11-24."""

    def __init__(self, name, conf):
        self.name = name
        if conf["type"] == "short":
            self.curve = ShortCurve(conf)
        elif conf["type"] == "edwards":
            from .edwards import EdwardsCurve
            self.curve = EdwardsCurve(conf)
        else:
            from .mont import MontCurve
            self.curve = MontCurve(conf)
        self.g = self.curve.g
        self.n = self.curve.n
        self.hash = HASHES[conf["hash"]]
        ref_assert(self.g.validate(), "Invalid curve")
        ref_assert(self.g.mul(self.n).is_infinity(), "Invalid curve, G*N != O")


_CACHE = {}


def get(name):
    """defineCurve lazy getter, curves.js:27-41."""
    if name not in _CACHE:
        ref_assert(name in _CONF, "Unknown curve " + name)
        conf = dict(_CONF[name])
        if name == "secp256k1":
            conf["g_pre"] = _secp256k1_pre(conf)
        _CACHE[name] = PresetCurve(name, conf)
    return _CACHE[name]


NAMES = tuple(_CONF)
