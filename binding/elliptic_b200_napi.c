/* N-API addon: thin marshalling between JS typed arrays and the C ABI of libelliptic_b200.so
 * (include/elliptic_b200.h).  No arithmetic here.  Build (in a Node.js toolchain):
 *   cc -shared -fPIC -I../include elliptic_b200_napi.c -L../elliptic_b200 -lelliptic_b200 \
 *      -Wl,-rpath,'$ORIGIN/../elliptic_b200' -o elliptic_b200.node
 * This image has no node / node_api.h, so the file is compile-checked against
 * binding/node_api_min.h only (tests/test_capi_load.py) and cannot be loaded here. */
#ifdef EB200_HAVE_NODE_API_H
#include <node_api.h>
#else
#include "node_api_min.h"
#endif
#include "../include/elliptic_b200.h"

static int u8(napi_env env, napi_value v, uint8_t** p, size_t* len) {
  napi_typedarray_type t; napi_value ab; size_t off; void* data;
  if (napi_get_typedarray_info(env, v, &t, len, &data, &ab, &off) != napi_ok || t != napi_uint8_array) return 0;
  *p = (uint8_t*)data;
  return 1;
}
static napi_value fail(napi_env env, int rc) {
  napi_throw_error(env, "EB200", rc == EB200_ERR_CUDA ? eb200_last_error() : eb200_strerror(rc));
  return 0;
}
static napi_value out_u8(napi_env env, size_t n, uint8_t** data) {
  napi_value ab, arr;
  napi_create_arraybuffer(env, n, (void**)data, &ab);
  napi_create_typedarray(env, napi_uint8_array, n, ab, 0, &arr);
  return arr;
}

/* init(device) */
static napi_value Init(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1]; int32_t dev = 0; napi_value undef;
  napi_get_cb_info(env, info, &argc, argv, 0, 0);
  if (argc >= 1) napi_get_value_int32(env, argv[0], &dev);
  int rc = eb200_init(dev);
  if (rc) return fail(env, rc);
  napi_get_undefined(env, &undef);
  return undef;
}

/* ecdsaVerifyBatch(curveId, e, r, s, pub, pubFmt) -> Uint8Array(n) of statuses
 * (EC.prototype.verify semantics per item, lib/elliptic/ec/index.js:188-229) */
static napi_value EcdsaVerifyBatch(napi_env env, napi_callback_info info) {
  size_t argc = 6; napi_value argv[6];
  napi_get_cb_info(env, info, &argc, argv, 0, 0);
  int32_t curve; uint32_t fmt; uint8_t *e, *r, *s, *pub, *st; size_t le, lr, ls, lp;
  if (argc < 6 || napi_get_value_int32(env, argv[0], &curve) != napi_ok || !u8(env, argv[1], &e, &le) ||
      !u8(env, argv[2], &r, &lr) || !u8(env, argv[3], &s, &ls) || !u8(env, argv[4], &pub, &lp) ||
      napi_get_value_uint32(env, argv[5], &fmt) != napi_ok)
    return fail(env, EB200_ERR_ARG);
  size_t len = curve == EB200_CURVE_P384 ? 48 : 32;
  size_t n = le / len;
  size_t pb = fmt == EB200_PUB_XY ? 2 * len : fmt == EB200_PUB_SEC1_65 ? 1 + 2 * len : 1 + len;
  if (le != n * len || lr != le || ls != le || lp != n * pb) return fail(env, EB200_ERR_ARG);
  napi_value arr = out_u8(env, n, &st);
  int rc = eb200_ecdsa_verify_batch(curve, n, e, r, s, pub, fmt, st);
  return rc ? fail(env, rc) : arr;
}

/* eddsaVerifyBatch(R, S, A, h) -> Uint8Array(n)   (EDDSA.prototype.verify, eddsa/index.js:52-63) */
static napi_value EddsaVerifyBatch(napi_env env, napi_callback_info info) {
  size_t argc = 4; napi_value argv[4];
  napi_get_cb_info(env, info, &argc, argv, 0, 0);
  uint8_t *R, *S, *A, *h, *st; size_t a, b, c, d;
  if (argc < 4 || !u8(env, argv[0], &R, &a) || !u8(env, argv[1], &S, &b) || !u8(env, argv[2], &A, &c) ||
      !u8(env, argv[3], &h, &d) || a % 32 || b != a || c != a || d != a)
    return fail(env, EB200_ERR_ARG);
  napi_value arr = out_u8(env, a / 32, &st);
  int rc = eb200_eddsa_verify_batch(a / 32, R, S, A, h, st);
  return rc ? fail(env, rc) : arr;
}

/* x25519DeriveBatch(priv, pubx) -> { out: Uint8Array(32 n), status: Uint8Array(n) }
 * (KeyPair.prototype.derive on curve25519, ec/key.js:102-107) */
static napi_value X25519DeriveBatch(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, 0, 0);
  uint8_t *k, *x, *out, *st; size_t a, b;
  if (argc < 3 || !u8(env, argv[0], &k, &a) || !u8(env, argv[1], &x, &b) || a % 32 || a != b) return fail(env, EB200_ERR_ARG);
  napi_value o = out_u8(env, a, &out), s = out_u8(env, a / 32, &st);
  int rc = eb200_x25519_derive_batch(a / 32, k, x, out, st);
  if (rc) return fail(env, rc);
  napi_set_named_property(env, argv[2], "out", o);      /* argv[2]: result object supplied by the JS wrapper */
  napi_set_named_property(env, argv[2], "status", s);
  return argv[2];
}

static size_t field_len(int curve) { return curve == EB200_CURVE_P384 ? 48 : 32; }

/* ecdsaVerifyBatchDer(curveId, e, sigs, sigOff (n + 1 little-endian u64 offsets, as a Uint8Array view), pub, pubFmt) -> Uint8Array(n)
 * DER signatures as `new Signature(der)` takes them (ec/signature.js:73-134), parsed on the GPU */
static napi_value EcdsaVerifyBatchDer(napi_env env, napi_callback_info info) {
  size_t argc = 6; napi_value argv[6];
  napi_get_cb_info(env, info, &argc, argv, 0, 0);
  int32_t curve; uint32_t fmt; uint8_t *e, *sig, *off, *pub, *st; size_t le, lsg, lo, lp;
  if (argc < 6 || napi_get_value_int32(env, argv[0], &curve) != napi_ok || !u8(env, argv[1], &e, &le) ||
      !u8(env, argv[2], &sig, &lsg) || !u8(env, argv[3], &off, &lo) || !u8(env, argv[4], &pub, &lp) ||
      napi_get_value_uint32(env, argv[5], &fmt) != napi_ok)
    return fail(env, EB200_ERR_ARG);
  size_t n = le / field_len(curve);
  if (lo != 8 * (n + 1)) return fail(env, EB200_ERR_ARG);
  napi_value arr = out_u8(env, n, &st);
  int rc = eb200_ecdsa_verify_batch_der(curve, n, e, sig, (const uint64_t*)off, pub, fmt, st);
  return rc ? fail(env, rc) : arr;
}

/* ecdsaSignBatch(curveId, e, priv, flags, result) -> result {r, s, recid, status}
 * (EC.prototype.sign with RFC 6979 nonces, ec/index.js:110-186) */
static napi_value EcdsaSignBatch(napi_env env, napi_callback_info info) {
  size_t argc = 5; napi_value argv[5];
  napi_get_cb_info(env, info, &argc, argv, 0, 0);
  int32_t curve; uint32_t flags; uint8_t *e, *d, *r, *s, *id, *st; size_t le, ld;
  if (argc < 5 || napi_get_value_int32(env, argv[0], &curve) != napi_ok || !u8(env, argv[1], &e, &le) ||
      !u8(env, argv[2], &d, &ld) || napi_get_value_uint32(env, argv[3], &flags) != napi_ok || le % 32 || ld != le)
    return fail(env, EB200_ERR_ARG);
  size_t n = le / 32;
  napi_value ar = out_u8(env, le, &r), as = out_u8(env, le, &s), ai = out_u8(env, n, &id), ast = out_u8(env, n, &st);
  int rc = eb200_ecdsa_sign_batch(curve, n, e, d, flags, r, s, id, st);
  if (rc) return fail(env, rc);
  napi_set_named_property(env, argv[4], "r", ar);
  napi_set_named_property(env, argv[4], "s", as);
  napi_set_named_property(env, argv[4], "recid", ai);
  napi_set_named_property(env, argv[4], "status", ast);
  return argv[4];
}

/* ecdsaRecoverBatch(curveId, e, r, s, recid, result) -> result {pub, status}   (recoverPubKey, ec/index.js:231-259) */
static napi_value EcdsaRecoverBatch(napi_env env, napi_callback_info info) {
  size_t argc = 6; napi_value argv[6];
  napi_get_cb_info(env, info, &argc, argv, 0, 0);
  int32_t curve; uint8_t *e, *r, *s, *id, *out, *st; size_t le, lr, ls, li;
  if (argc < 6 || napi_get_value_int32(env, argv[0], &curve) != napi_ok || !u8(env, argv[1], &e, &le) ||
      !u8(env, argv[2], &r, &lr) || !u8(env, argv[3], &s, &ls) || !u8(env, argv[4], &id, &li) || le % 32 ||
      lr != le || ls != le || li != le / 32)
    return fail(env, EB200_ERR_ARG);
  size_t n = le / 32;
  napi_value ao = out_u8(env, 64 * n, &out), ast = out_u8(env, n, &st);
  int rc = eb200_ecdsa_recover_batch(curve, n, e, r, s, id, out, st);
  if (rc) return fail(env, rc);
  napi_set_named_property(env, argv[5], "pub", ao);
  napi_set_named_property(env, argv[5], "status", ast);
  return argv[5];
}

/* mulAddBatch(curveId, k1 | null, k2, points | null, result) -> result {points, status}
 * k1 null: Point.mul (short.js:422-432); points null: G.mul; both given: G.mulAdd(k1, P, k2) (short.js:434-441) */
static napi_value MulAddBatch(napi_env env, napi_callback_info info) {
  size_t argc = 5; napi_value argv[5];
  napi_get_cb_info(env, info, &argc, argv, 0, 0);
  int32_t curve; uint8_t *k1 = 0, *k2, *pts = 0, *out, *st; size_t l1 = 0, l2, lp = 0;
  if (argc < 5 || napi_get_value_int32(env, argv[0], &curve) != napi_ok || !u8(env, argv[2], &k2, &l2))
    return fail(env, EB200_ERR_ARG);
  u8(env, argv[1], &k1, &l1);          /* null / undefined leave the pointer at 0 */
  u8(env, argv[3], &pts, &lp);
  size_t len = field_len(curve), n = l2 / len;
  if (l2 != n * len || (k1 && l1 != l2) || (pts && lp != 2 * l2) || (k1 && !pts)) return fail(env, EB200_ERR_ARG);
  napi_value ao = out_u8(env, 2 * len * n, &out), ast = out_u8(env, n, &st);
  int rc = k1 ? eb200_mul_add_batch(curve, n, k1, k2, pts, out, st) : eb200_scalar_mul_batch(curve, n, k2, pts, out, st);
  if (rc) return fail(env, rc);
  napi_set_named_property(env, argv[4], "points", ao);
  napi_set_named_property(env, argv[4], "status", ast);
  return argv[4];
}

/* ecdhDeriveBatch(curveId, priv, pubXY, result) -> result {out, status}   (KeyPair.derive, ec/key.js:102-107) */
static napi_value EcdhDeriveBatch(napi_env env, napi_callback_info info) {
  size_t argc = 4; napi_value argv[4];
  napi_get_cb_info(env, info, &argc, argv, 0, 0);
  int32_t curve; uint8_t *k, *pts, *out, *st; size_t lk, lp;
  if (argc < 4 || napi_get_value_int32(env, argv[0], &curve) != napi_ok || !u8(env, argv[1], &k, &lk) ||
      !u8(env, argv[2], &pts, &lp) || lk % field_len(curve) || lp != 2 * lk)
    return fail(env, EB200_ERR_ARG);
  size_t n = lk / field_len(curve);
  napi_value ao = out_u8(env, lk, &out), ast = out_u8(env, n, &st);
  int rc = eb200_ecdh_derive_batch(curve, n, k, pts, out, st);
  if (rc) return fail(env, rc);
  napi_set_named_property(env, argv[3], "out", ao);
  napi_set_named_property(env, argv[3], "status", ast);
  return argv[3];
}

static napi_value Register(napi_env env, napi_value exports) {
  static const struct { const char* name; napi_callback cb; } fns[] = {
      {"init", Init}, {"ecdsaVerifyBatch", EcdsaVerifyBatch}, {"eddsaVerifyBatch", EddsaVerifyBatch},
      {"x25519DeriveBatch", X25519DeriveBatch}, {"ecdsaVerifyBatchDer", EcdsaVerifyBatchDer},
      {"ecdsaSignBatch", EcdsaSignBatch}, {"ecdsaRecoverBatch", EcdsaRecoverBatch}, {"mulAddBatch", MulAddBatch},
      {"ecdhDeriveBatch", EcdhDeriveBatch}};
  for (unsigned i = 0; i < sizeof fns / sizeof fns[0]; i++) {
    napi_value f;
    napi_create_function(env, fns[i].name, (size_t)-1, fns[i].cb, 0, &f);
    napi_set_named_property(env, exports, fns[i].name, f);
  }
  return exports;
}
static napi_module eb200_module = {1, 0, __FILE__, Register, "elliptic_b200", 0, {0}};
__attribute__((constructor)) static void eb200_register(void) { napi_module_register(&eb200_module); }
