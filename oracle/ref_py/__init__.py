"""ORACLE -- CPU restatement of the reference's hot path (indutny/elliptic
v6.6.1 + vendored bn.js 4.11.9).  TEST INFRASTRUCTURE ONLY: nothing under
oracle/ may be imported, linked or executed by the product package
`elliptic_b200`; only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs use it, and only as the checker.

Parity status: pinned against the reference's own fixtures (ed25519 sign.input
and derivation fixtures, p256/p384 Maxwell vectors, RFC 6979, secp256k1
precomputed-table digest, SEC1 / ladder KATs) -- see tests/test_oracle_golden.py
-- and cross-checked against OpenSSL (`cryptography`) and libsodium (`PyNaCl`).
"""
from .bn import RefError  # noqa: F401
