"""ORACLE (test infrastructure only): ctypes loader for oracle/c/libk256_ref.so,
the C restatement of the reference's secp256k1 verify algorithm (CPU baseline)."""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_LIB = os.path.join(_DIR, "libk256_ref.so")


def build():
    subprocess.run(["make", "-C", _DIR, "-s"], check=True)
    return _LIB


def load():
    if not os.path.exists(_LIB):
        build()
    lib = ctypes.CDLL(_LIB)
    lib.k256_ref_verify_batch.argtypes = [ctypes.c_size_t] + [ctypes.c_void_p] * 5 + [ctypes.c_int]
    lib.k256_ref_fm_count.restype = ctypes.c_ulong
    return lib


def verify_batch(e, r, s, pub, threads=1):
    lib = load()
    e, r, s, pub = (np.ascontiguousarray(a, np.uint8) for a in (e, r, s, pub))
    n = e.shape[0]
    st = np.zeros(n, np.uint8)
    lib.k256_ref_verify_batch(n, e.ctypes.data, r.ctypes.data, s.ctypes.data, pub.ctypes.data, st.ctypes.data, threads)
    return st
