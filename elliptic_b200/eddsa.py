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
        if not (R.shape == (n, 32) and S.shape == R.shape and A.shape == R.shape and h.shape == R.shape):
            raise ValueError("R, S, A, h must be (n, 32) uint8 arrays")
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
        if not (R.shape == (n, 32) and S.shape == R.shape and A.shape == R.shape):
            raise ValueError("R, S, A must be (n, 32) uint8 arrays")
        if not (msg_off.shape == (n + 1,) and int(msg_off[n]) == msgs.size):
            raise ValueError("msg_off must hold n + 1 offsets, the last one equal to len(msgs)")
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

    def sign_batch_packed(self, secrets, msgs, msg_off, want_pub=False):
        """secrets: (n, 32) uint8; msgs: concatenated message bytes; msg_off: n + 1 uint64 offsets.
        Returns the (n, 64) signatures Rencoded || S (and the (n, 32) public keys with want_pub)."""
        lib = nat.init(self._device)
        secrets = np.ascontiguousarray(secrets, dtype=np.uint8)
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
        msg_off = np.ascontiguousarray(msg_off, dtype=np.uint64)
        n = secrets.shape[0]
        if secrets.shape != (n, 32) or msg_off.shape != (n + 1,) or int(msg_off[n]) != msgs.size:
            raise EllipticError("sign_batch_packed: secrets (n, 32), n + 1 offsets covering msgs expected")
        sig = np.empty((n, 64), np.uint8)
        pub = np.empty((n, 32), np.uint8) if want_pub else None
        st = np.empty(n, np.uint8)
        nat.check(lib.eb200_eddsa_sign_batch(n, secrets.ctypes.data, msgs.ctypes.data if msgs.size else None, msg_off.ctypes.data,
                                             sig.ctypes.data, pub.ctypes.data if want_pub else None, st.ctypes.data))
        if not bool((st == nat.ST_TRUE).all()):
            raise nat.NativeError("eddsa sign: unexpected status")
        return (sig, pub) if want_pub else sig

    def sign_batch(self, messages, secrets):
        """EDDSA#signBatch: lists of the reference's own argument forms (hex strings / byte arrays); secrets as
        eddsa.keyFromSecret takes them.  Returns a list of 64-byte signatures (sig.toBytes())."""
        n = len(messages)
        sec = np.zeros((n, 32), np.uint8)
        ms = []
        for i in range(n):
            sk = _parse_bytes(secrets[i])
            if len(sk) != 32:
                raise EllipticError("unsupported secret length %d" % len(sk))
            sec[i] = np.frombuffer(bytes(sk), np.uint8)
            ms.append(bytes(_parse_bytes(messages[i])))
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum([len(m) for m in ms])
        sig = self.sign_batch_packed(sec, np.frombuffer(b"".join(ms), np.uint8), off)
        return [sig[i].tobytes() for i in range(n)]

    def sign(self, message, secret):
        """EDDSA.prototype.sign (eddsa/index.js:34-44): the 64 signature bytes (sig.toBytes())."""
        return self.sign_batch([message], [secret])[0]

    def public_from_secret_batch(self, secrets):
        """key.getPublic('bytes') for a batch of secrets (a fixed-base multiplication each)."""
        secrets = np.ascontiguousarray(secrets, dtype=np.uint8)
        n = secrets.shape[0]
        _, pub = self.sign_batch_packed(secrets, np.zeros(0, np.uint8), np.zeros(n + 1, np.uint64), want_pub=True)
        return pub

    def verify(self, message, sig, pub):
        """EDDSA.prototype.verify (eddsa/index.js:52-63): bool, or raises."""
        st = int(self.verify_batch([message], [sig], [pub])[0])
        if st == nat.ST_TRUE:
            return True
        if st == nat.ST_FALSE:
            return False
        raise EllipticError(_THROW_MSG.get(st, "status %d" % st))
