"""BASELINE.json configs[0] ("config 1"): the 1024 fixed secp256k1 (msgHash, sig, pub) triples of
tests/golden/secp256k1_verify_1024.json.gz (minted by tests/golden/make_config1.py, reference forms:
hex hash, DER / {r,s} signature, SEC1 / {x,y} key) through the oracle, the C port, OpenSSL and -- with
-m gpu -- the CUDA path behind the reference-shaped host API."""
import gzip
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


@pytest.fixture(scope="module")
def fixture():
    d = json.load(gzip.open(os.path.join(HERE, "golden", "secp256k1_verify_1024.json.gz"), "rt"))
    assert len(d["items"]) == 1024
    return d["items"]


def test_fixture_shape(fixture):
    kinds = {}
    for it in fixture:
        kinds[it["kind"]] = kinds.get(it["kind"], 0) + 1
    assert kinds["valid"] == 768 and sum(v for k, v in kinds.items() if k.startswith("valid_")) == 128
    assert sum(v for k, v in kinds.items() if not k.startswith("valid")) == 128 and len(kinds) == 22
    assert sum(it["expected"] is True for it in fixture) == 920
    assert sum(isinstance(it["expected"], str) for it in fixture) == 24
    # both eqXToP candidates, both signature encodings and all four key encodings occur
    assert any(it["kind"] == "r_plus_n_candidate" and it["expected"] is True for it in fixture)
    assert any(isinstance(it["sig"], str) for it in fixture) and any(isinstance(it["sig"], dict) for it in fixture)
    assert {(it["pub"][:2] if isinstance(it["pub"], str) else "xy") for it in fixture} >= {"04", "02", "03", "06", "07", "xy"}


def test_oracle_reproduces_the_fixture(fixture):
    from oracle.ref_py.bn import RefError
    from oracle.ref_py.ec import EC
    ec = EC("secp256k1")
    for it in fixture:
        try:
            got = bool(ec.verify(it["msg"], it["sig"], it["pub"], "hex"))
        except RefError as ex:
            got = "throw:" + ex.args[0]
        assert got == it["expected"], it["i"]


def test_openssl_agrees_where_it_can_express_the_item(fixture):
    spec = importlib.util.spec_from_file_location("make_config1", os.path.join(HERE, "golden", "make_config1.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    checked = 0
    for it in fixture[::3] + fixture[768:]:
        v = mod.openssl(it["msg"], it["sig"], it["pub"])
        assert v == it["openssl"]
        if v is not None:
            assert v == it["expected"], it["i"]
            checked += 1
    assert checked > 450


def _packed(fixture):
    """Items that fit the packed (e, r, s, x, y) form: parsed by the oracle's own importers."""
    from oracle.ref_py.bn import RefError
    from oracle.ref_py.ec import EC
    from oracle.ref_py.signature import Signature
    ec = EC("secp256k1")
    rows, exp = [], []
    for it in fixture:
        if isinstance(it["expected"], str):
            continue
        sig = Signature(it["sig"], "hex")
        if sig.r >> 256 or sig.s >> 256:
            continue
        q = ec.key_from_public(it["pub"], "hex").get_public()
        rows.append((int(it["msg"], 16), sig.r, sig.s, q.x, q.y))
        exp.append(int(it["expected"]))
    col = lambda k: np.frombuffer(b"".join(r[k].to_bytes(32, "big") for r in rows), np.uint8).reshape(-1, 32)
    return col(0), col(1), col(2), np.concatenate([col(3), col(4)], axis=1), np.array(exp, np.uint8)


def test_c_port_agrees(fixture):
    from oracle import c_oracle
    e, r, s, pub, exp = _packed(fixture)
    assert len(exp) >= 990
    assert np.array_equal(c_oracle.verify_batch(e, r, s, pub, 2), exp)


@pytest.mark.gpu
def test_gpu_reference_forms(native, fixture):
    """The reference-shaped call: verify_batch for everything that returns, verify() for everything that throws."""
    from elliptic_b200.ec import EC, EllipticError
    ec = EC("secp256k1")
    ok = [it for it in fixture if not isinstance(it["expected"], str)]
    st = ec.verify_batch([it["msg"] for it in ok], [it["sig"] for it in ok], [it["pub"] for it in ok], "hex")
    assert [bool(v) for v in st] == [it["expected"] for it in ok] and set(int(v) for v in st) <= {0, 1}
    for it in fixture:
        if isinstance(it["expected"], str):
            with pytest.raises(EllipticError) as ei:
                ec.verify(it["msg"], it["sig"], it["pub"], "hex")
            assert str(ei.value) == it["expected"][len("throw:"):], it["i"]


@pytest.mark.gpu
def test_gpu_packed_and_der_entry_points(native, fixture):
    """The same items through the packed C ABI forms: x||y keys, and DER signatures + SEC1 keys parsed on the GPU."""
    from elliptic_b200 import _native as nat
    from elliptic_b200.ec import EC
    ec = EC("secp256k1")
    e, r, s, pub, exp = _packed(fixture)
    assert np.array_equal(ec.verify_batch_packed(e, r, s, pub), exp)
    for tag, fmt in (("04", nat.PUB_SEC1_65), ("02", nat.PUB_SEC1_33), ("03", nat.PUB_SEC1_33)):
        sel = [it for it in fixture if isinstance(it["sig"], str) and isinstance(it["pub"], str) and it["pub"][:2] == tag
               and it["kind"] != "bad_der"]
        if not sel:
            continue
        em = np.frombuffer(b"".join(bytes.fromhex(it["msg"]) for it in sel), np.uint8).reshape(-1, 32)
        pk = np.frombuffer(b"".join(bytes.fromhex(it["pub"]) for it in sel), np.uint8).reshape(len(sel), -1)
        st = ec.verify_batch_der_packed(em, [bytes.fromhex(it["sig"]) for it in sel], pk, fmt)
        want = [int(it["expected"]) if not isinstance(it["expected"], str) else {"throw:invalid point": 2}[it["expected"]] for it in sel]
        assert [int(v) for v in st] == want
