"""The C-ABI library loads on a CPU-only box, exports every symbol include/elliptic_b200.h
declares, and refuses to compute without a GPU (no fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from elliptic_b200 import _native, build
    build.build()
    lib = _native.load()
    hdr = open(os.path.join(ROOT, "include", "elliptic_b200.h")).read()
    declared = set(re.findall(r"\b(eb200_\w+)\s*\(", hdr))
    assert declared and declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from elliptic_b200 import _native
    lib = _native.load()
    import ctypes
    dev0 = (ctypes.c_int * 1)(0)
    assert lib.eb200_init(dev0, 1, 0) == _native.ERR_NO_DEVICE
    assert lib.eb200_init(None, 0, 0) == _native.ERR_NO_DEVICE and lib.eb200_device_count() == 0
    import numpy as np
    from elliptic_b200.ec import EC
    z = np.zeros((1, 32), np.uint8)
    with pytest.raises(_native.NativeError):
        EC("secp256k1").verify_batch_packed(z, z, z, np.zeros((1, 64), np.uint8))


def test_product_package_does_not_import_the_oracle():
    import subprocess
    import sys
    code = "import sys; import elliptic_b200, elliptic_b200.ec; assert not [m for m in sys.modules if m.startswith('oracle')]"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    # no product source file imports, includes or links anything under oracle/ (comments may mention the oracle)
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|#\s*include\s+[\"<][^\">]*oracle)", re.M)
    for top in ("elliptic_b200", "include", "binding"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".js", ".inc")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not pat.search(src), os.path.join(dirpath, f)
                    assert "libk256_ref" not in src and "c_oracle" not in src, os.path.join(dirpath, f)


def test_host_side_parsing_matches_reference_forms():
    from elliptic_b200.ec import EC, parse_der, EllipticError
    from oracle.ref_py.signature import Signature
    from oracle.ref_py.ec import EC as RefEC
    import random
    ec, ref = EC("secp256k1"), RefEC("secp256k1")
    rnd = random.Random(4)
    for _ in range(200):
        r, s = rnd.randrange(1, 2**rnd.choice([8, 64, 255, 256])), rnd.randrange(1, 2**rnd.choice([8, 128, 256]))
        der = Signature({"r": r, "s": s}).to_der()
        assert parse_der(der) == (r, s)
        bad = bytearray(der); bad[rnd.randrange(len(bad))] ^= 1 << rnd.randrange(8)
        got = parse_der(bytes(bad))
        chk = Signature.__new__(Signature)
        ok = chk._import_der(bytes(bad), None)
        assert (got is not None) == bool(ok) and (got is None or got == (chk.r, chk.s))
    for m in (b"\x00" * 32, b"\xff" * 32, b"\xff" * 40, "abc", "00" * 40, 12345, 2**300 + 5):
        assert ec._truncate_to_n(m) == ref._truncate_to_n(m)
    with pytest.raises(EllipticError):
        ec._public("05" + "00" * 32, "hex")


def test_napi_shim_compiles_against_the_header():
    """binding/elliptic_b200_napi.c (the N-API addon a Node.js maintainer builds) must stay in step with
    include/elliptic_b200.h: compile it (no Node.js here: against binding/node_api_min.h) with warnings as errors,
    and check that it binds every host-pointer batch entry point of the header."""
    import subprocess
    src = os.path.join(ROOT, "binding", "elliptic_b200_napi.c")
    subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "binding"),
                    "-I" + os.path.join(ROOT, "include"), src], check=True)
    code = open(src).read()
    hdr = open(os.path.join(ROOT, "include", "elliptic_b200.h")).read()
    declared = set(re.findall(r"\b(eb200_\w+)\s*\(", hdr))
    not_for_js = {n for n in declared if n.endswith("_dev") or n.startswith("eb200_selftest") or n.endswith("workspace_bytes")}
    not_for_js |= {"eb200_shutdown", "eb200_device_count", "eb200_last_timing"}
    missing = sorted(n for n in declared - not_for_js if n + "(" not in code)
    assert not missing, missing
    js = open(os.path.join(ROOT, "binding", "index.js")).read()
    for name in re.findall(r'\{"(\w+)", \w+\}', code):
        assert name == "init" or ("native." + name) in js, name       # every addon function is reachable from the JS wrapper
