"""Host-side mirror of the reference's EdDSA object for the accelerated path.

`EDDSA('ed25519')` corresponds to `new elliptic.eddsa('ed25519')`
(lib/elliptic/eddsa/index.js:11-25); `verify(message, sig, pub)` keeps the reference's
argument forms and error behaviour (eddsa/index.js:52-63) and `verify_batch` is the new
batch entry point.  Byte/hex parsing and the SHA-512 of R || A || M are done here
(hashlib; the reference uses hash.js); all curve arithmetic runs on the GPU.
"""
import hashlib

import numpy as np

from . import _native as nat
from .ec import EllipticError, _to_array, _THROW_MSG

N_ED25519 = 0x1000000000000000000000000000000014DEF9DEA2F79CD65812631A5CF5D3ED


def _parse_bytes(x):
    """utils.parseBytes (lib/elliptic/utils.js:112-116)."""
    return _to_array(x, "hex") if isinstance(x, str) else _to_array(x)


class EDDSA:
    def __init__(self, curve="ed25519", device=0):
        if curve != "ed25519":
            raise EllipticError("only tested with ed25519 so far")      # eddsa/index.js:12
        self.encoding_length = 32
        self.n = N_ED25519
        self._device = device

    def hash_int(self, *parts):
        """EDDSA.hashInt (eddsa/index.js:65-70)."""
        h = hashlib.sha512()
        for p in parts:
            h.update(bytes(p))
        return int.from_bytes(h.digest(), "little") % self.n

    def verify_batch_packed(self, R, S, A, h):
        """R, S, A, h: (n, 32) uint8 arrays (little-endian wire forms; h = hashInt(R, A, M) < n)."""
        lib = nat.init(self._device)
        R, S, A, h = (np.ascontiguousarray(a, dtype=np.uint8) for a in (R, S, A, h))
        n = R.shape[0]
        assert R.shape == (n, 32) and S.shape == R.shape and A.shape == R.shape and h.shape == R.shape
        status = np.empty(n, np.uint8)
        nat.check(lib.eb200_eddsa_verify_batch(n, R.ctypes.data, S.ctypes.data, A.ctypes.data, h.ctypes.data,
                                               status.ctypes.data))
        return status

    def verify_batch_msgs_packed(self, R, S, A, msgs, msg_off):
        """Like verify_batch_packed but takes the raw messages (concatenated bytes + n+1 offsets);
        SHA-512 and the reduction mod n run on the GPU."""
        lib = nat.init(self._device)
        R, S, A = (np.ascontiguousarray(a, dtype=np.uint8) for a in (R, S, A))
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
        msg_off = np.ascontiguousarray(msg_off, dtype=np.uint64)
        n = R.shape[0]
        assert msg_off.shape == (n + 1,) and int(msg_off[n]) == msgs.size
        status = np.empty(n, np.uint8)
        nat.check(lib.eb200_eddsa_verify_batch_msgs(n, R.ctypes.data, S.ctypes.data, A.ctypes.data,
                                                    msgs.ctypes.data if msgs.size else None, msg_off.ctypes.data,
                                                    status.ctypes.data))
        return status

    def verify_batch(self, messages, sigs, pubs, gpu_hash=True):
        """EDDSA#verifyBatch: lists of the reference's own argument forms (hex strings / byte arrays).
        gpu_hash=False computes hashInt with hashlib on the host instead of on the GPU."""
        n = len(messages)
        R = np.zeros((n, 32), np.uint8)
        S = np.zeros((n, 32), np.uint8)
        A = np.zeros((n, 32), np.uint8)
        h = np.zeros((n, 32), np.uint8)
        if gpu_hash:
            ms = []
            for i in range(n):
                sig = _parse_bytes(sigs[i])
                if len(sig) != 2 * self.encoding_length:
                    raise EllipticError("Signature has invalid size")   # eddsa/signature.js:23-24
                pub = _parse_bytes(pubs[i])
                if len(pub) != self.encoding_length:
                    raise EllipticError("unsupported public key length %d" % len(pub))
                R[i] = np.frombuffer(sig[:32], np.uint8)
                S[i] = np.frombuffer(sig[32:], np.uint8)
                A[i] = np.frombuffer(pub, np.uint8)
                ms.append(_parse_bytes(messages[i]))
            off = np.zeros(n + 1, np.uint64)
            off[1:] = np.cumsum([len(m) for m in ms])
            return self.verify_batch_msgs_packed(R, S, A, np.frombuffer(b"".join(ms), np.uint8), off)
        for i in range(n):
            msg = _parse_bytes(messages[i])
            sig = _parse_bytes(sigs[i])
            if len(sig) != 2 * self.encoding_length:
                raise EllipticError("Signature has invalid size")       # eddsa/signature.js:23-24
            pub = _parse_bytes(pubs[i])
            if len(pub) != self.encoding_length:
                # decodePoint on another length reads a different y; not on the accelerated path
                raise EllipticError("unsupported public key length %d" % len(pub))
            R[i] = np.frombuffer(sig[:32], np.uint8)
            S[i] = np.frombuffer(sig[32:], np.uint8)
            A[i] = np.frombuffer(pub, np.uint8)
            h[i] = np.frombuffer(self.hash_int(sig[:32], pub, msg).to_bytes(32, "little"), np.uint8)
        return self.verify_batch_packed(R, S, A, h)

    def verify(self, message, sig, pub):
        """EDDSA.prototype.verify (eddsa/index.js:52-63): bool, or raises."""
        st = int(self.verify_batch([message], [sig], [pub])[0])
        if st == nat.ST_TRUE:
            return True
        if st == nat.ST_FALSE:
            return False
        raise EllipticError(_THROW_MSG.get(st, "status %d" % st))
