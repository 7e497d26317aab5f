"""elliptic_b200 -- B200-native batch engine behind indutny/elliptic's `ec` API.

Host side (Python here because no Node.js toolchain exists in this image; the
N-API shim a maintainer would add is in binding/ and INTEGRATION.md) mirroring
`require('elliptic')` for the accelerated path (lib/elliptic.js:5-13).
"""
from .ec import EC as ec  # noqa: F401,N813  (reference export name)
from .eddsa import EDDSA as eddsa  # noqa: F401,N813
from . import _native  # noqa: F401
from . import curve  # noqa: F401   (curve.ShortCurve: run-time short Weierstrass parameters, lib/elliptic/curve/short.js)

version = "0.1.0"
