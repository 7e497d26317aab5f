// Host emulation of the kernel bodies -- TEST INFRASTRUCTURE ONLY.
// Compiles the same .cuh bodies the CUDA kernels use with their portable C++
// fallbacks so the kernel *logic* (recoding, tables, exceptional cases, status
// codes) can be checked against the oracle on a CPU-only box.  The product
// library (libelliptic_b200.so) never contains or calls this code.
#include <vector>
#include <cstring>
#include "../../elliptic_b200/csrc/ecdsa_k256_body.cuh"
using namespace eb;

extern "C" {

// op: 0 mul, 1 sqr, 2 add, 3 sub, 4 neg, 5 mul_small(k=b[0]), 6 normalize, 7 inv, 8 sqrt_candidate
void he_fe_op(int op, const u32* a, const u32* b, u32* out) {
  fe A = load_fe(a), B = load_fe(b), R;
  switch (op) {
    case 0: R = fe_mul(A, B); break;
    case 1: R = fe_sqr(A); break;
    case 2: R = fe_add(A, B); break;
    case 3: R = fe_sub(A, B); break;
    case 4: R = fe_neg(A); break;
    case 5: R = fe_mul_small(A, b[0]); break;
    case 6: R = fe_normalize(A); break;
    case 7: R = fe_inv(A); break;
    case 8: R = fe_sqrt_candidate(A); break;
    default: R = fe_zero();
  }
  store_fe(out, R);
}

void he_sc_mont_mul(const u32* a, const u32* b, u32* out) { sc_mont_mul(out, a, b); }

void he_glv(const u32* k, u32* m1, int* n1, u32* m2, int* n2) {
  bool a, b;
  glv_split_odd(k, m1, &a, m2, &b);
  *n1 = a; *n2 = b;
}

void he_gtab(u32* out /* 32*128*16 */, int jlo, int jhi) {
  for (int j = jlo; j < jhi; j++)
    for (int i = 0; i < GTAB_ENTRIES; i++) gtab_entry(j, i, out + ((size_t)j * GTAB_ENTRIES + i) * 16);
}

void he_gtab_dims(int* windows, int* entries, int* wbits) { *windows = GTAB_WINDOWS; *entries = GTAB_ENTRIES; *wbits = GTAB_W; }

// Fast gtab build for tests: incremental instead of per-entry scalar mults.
void he_gtab_fast(u32* out) {
  ge_jac base = jac_from_aff(k256_G());
  for (int j = 0; j < GTAB_WINDOWS; j++) {
    ge_aff b = jac_to_aff(base);
    ge_jac d = jac_dbl(jac_from_aff(b));
    ge_jac acc = jac_from_aff(b);
    for (int i = 0; i < GTAB_ENTRIES; i++) {
      ge_aff r = jac_to_aff(acc);
      r.x = fe_normalize(r.x); r.y = fe_normalize(r.y);
      store_fe(out + ((size_t)j * GTAB_ENTRIES + i) * 16, r.x);
      store_fe(out + ((size_t)j * GTAB_ENTRIES + i) * 16 + 8, r.y);
      acc = jac_add_inl(acc, d);
    }
    for (int k = 0; k < GTAB_W; k++) base = jac_dbl(base);
  }
}

void he_verify(size_t N, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* pub,
               const u32* gtab, uint8_t* status) {
  std::vector<u32> ws((size_t)PREP_WORDS * N), scratch((size_t)8 * N), qtab((size_t)QTAB_WORDS * N);
  size_t T = (N + PREP_BATCH - 1) / PREP_BATCH;
  for (size_t t = 0; t < T; t++) prep_thread(t, T, N, e, r, s, ws.data(), scratch.data());
  for (size_t i = 0; i < N; i++) status[i] = verify_item(i, N, pub, r, ws.data(), gtab, qtab.data());
}
}

// ---------------------------------------------------------------------------
// generic short-Weierstrass path (p256 / p384) through the same bodies
#include "../../elliptic_b200/csrc/ecdsa_sw_body.cuh"
template <class C>
static void sw_verify_host(size_t N, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* pub, uint8_t* status) {
  typedef SW<C> W;
  static std::vector<u32> gtab;
  if (gtab.empty()) {
    gtab.resize((size_t)W::GWINDOWS * W::GENTRIES * 2 * W::N);
    typename W::aff g; g.x = C::gx(); g.y = C::gy();
    typename W::jac base = W::from_aff(g);
    for (int j = 0; j < W::GWINDOWS; j++) {
      typename W::aff b = W::to_aff(base);
      typename W::jac d = W::dbl(W::from_aff(b));
      typename W::jac acc = W::from_aff(b);
      for (int i = 0; i < W::GENTRIES; i++) {
        typename W::aff a = W::to_aff(acc);
        store_fe_n<W::N>(&gtab[((size_t)j * W::GENTRIES + i) * 2 * W::N], a.x);
        store_fe_n<W::N>(&gtab[((size_t)j * W::GENTRIES + i) * 2 * W::N + W::N], a.y);
        acc = W::add(acc, d);
      }
      for (int k = 0; k < W::GW; k++) base = W::dbl(base);
    }
  }
  std::vector<u32> ws((size_t)W::PREP_WORDS * N), scratch((size_t)W::N * N), qtab((size_t)W::QTAB_WORDS * N);
  size_t T = (N + W::BATCH - 1) / W::BATCH;
  for (size_t t = 0; t < T; t++) W::prep_thread(t, T, N, e, r, s, ws.data(), scratch.data());
  for (size_t i = 0; i < N; i++) status[i] = W::verify_item(i, N, pub, r, ws.data(), gtab.data(), qtab.data());
}
extern "C" void he_sw_verify(int curve, size_t N, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* pub, uint8_t* status) {
  if (curve == 2) sw_verify_host<P256>(N, e, r, s, pub, status);
  else if (curve == 6) sw_verify_host<P521>(N, e, r, s, pub, status);
  else if (curve == 7) sw_verify_host<P192>(N, e, r, s, pub, status);
  else if (curve == 8) sw_verify_host<P224>(N, e, r, s, pub, status);
  else sw_verify_host<P384>(N, e, r, s, pub, status);
}

// ---------------------------------------------------------------------------
// ed25519 verify / curve25519 derive through the same bodies
#include "../../elliptic_b200/csrc/ed25519_body.cuh"
static const std::vector<u32>& ed_host_gtab() {
  static std::vector<u32> gtab;
  if (gtab.empty()) {
    gtab.resize((size_t)ED_GWINDOWS * ED_GENTRIES * 24);
    ed_ext base = ed_identity();
    ed_G(&base.x, &base.y);
    base.t = f25_mul(base.x, base.y);
    for (int j = 0; j < ED_GWINDOWS; j++) {
      ed_cached cb = ed_to_cached(base);
      ed_ext acc = ed_identity();
      for (int i = 0; i < ED_GENTRIES; i++) {
        f25 zi = f25_inv(acc.z);
        f25 x = f25_mul(acc.x, zi), y = f25_mul(acc.y, zi);
        u32* o = &gtab[((size_t)j * ED_GENTRIES + i) * 24];
        f25_store(o, f25_normalize(f25_add(y, x)));
        f25_store(o + 8, f25_normalize(f25_sub(y, x)));
        f25_store(o + 16, f25_normalize(f25_mul(f25_mul(x, y), f25_2d())));
        acc = ed_add_cached(acc, cb);
      }
      for (int k = 0; k < ED_GW; k++) base = ed_dbl(base);
    }
  }
  return gtab;
}
extern "C" void he_ed25519_verify(size_t N, const uint8_t* R, const uint8_t* S, const uint8_t* A, const uint8_t* h, uint8_t* status) {
  const std::vector<u32>& gtab = ed_host_gtab();
  std::vector<u32> atab((size_t)ED_ATAB_WORDS * N);
  for (size_t i = 0; i < N; i++) status[i] = ed25519_verify_item(i, R, S, A, h, gtab.data(), atab.data());
}
extern "C" void he_ed25519_sign(size_t N, const uint8_t* secrets, const uint8_t* msgs, const u64* off, uint8_t* sig, uint8_t* pub) {
  const std::vector<u32>& gtab = ed_host_gtab();
  for (size_t i = 0; i < N; i++) ed25519_sign_item(i, secrets, msgs, off, gtab.data(), sig, pub);
}
extern "C" void he_x25519_derive(size_t N, const uint8_t* priv, const uint8_t* pubx, uint8_t* out, uint8_t* status) {
  for (size_t i = 0; i < N; i++) status[i] = x25519_derive_item(i, priv, pubx, out);
}
extern "C" void he_f25_op(int op, const u32* a, const u32* b, u32* out) {
  f25 A = f25_load(a), B = f25_load(b), R;
  switch (op) {
    case 0: R = f25_mul(A, B); break;
    case 1: R = f25_sqr(A); break;
    case 2: R = f25_add(A, B); break;
    case 3: R = f25_sub(A, B); break;
    case 4: R = f25_neg(A); break;
    case 5: R = f25_mul_small(A, b[0]); break;
    case 6: R = f25_normalize(A); break;
    case 7: R = f25_inv(A); break;
    case 8: R = f25_pow_p58(A); break;
    case 9: R = f25_legendre(A); break;
    default: R = f25_zero();
  }
  f25_store(out, R);
}

// ---------------------------------------------------------------------------
// exact replay of the reference schedule (off-curve keys)
#include "../../elliptic_b200/csrc/ecdsa_k256_replay.cuh"
static std::vector<u32>& replay_tab() {
  static std::vector<u32> tab;
  if (tab.empty()) {
    tab.resize(REPLAY_TAB_WORDS);
    for (int t = 0; t < 2 * REPLAY_NAF_PTS; t++) rp_tab_entry(t, &tab[16 * t]);
  }
  return tab;
}
extern "C" void he_replay_jmuladd(const u32* u1, const u32* u2, const u32* qx, const u32* qy, u32* xyz) {
  ge_jac r = rp_jmul_add(u1, u2, load_fe(qx), load_fe(qy), replay_tab().data());
  store_fe(xyz, fe_normalize(r.x)); store_fe(xyz + 8, fe_normalize(r.y)); store_fe(xyz + 16, fe_normalize(r.z));
}
extern "C" void he_replay_verify(size_t N, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* pub, uint8_t* status) {
  for (size_t i = 0; i < N; i++) status[i] = rp_verify_item(i, e, r, s, pub, replay_tab().data());
}

extern "C" void he_recover(size_t N, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* recid,
                           const u32* gtab, uint8_t* out, uint8_t* status) {
  std::vector<u32> ws((size_t)PREP_WORDS * N), scratch((size_t)8 * N), qtab((size_t)QTAB_WORDS * N);
  size_t T = (N + PREP_BATCH - 1) / PREP_BATCH;
  for (size_t t = 0; t < T; t++) prep_thread(t, T, N, e, r, s, ws.data(), scratch.data(), 1);
  for (size_t i = 0; i < N; i++) status[i] = recover_item(i, N, r, recid, ws.data(), gtab, qtab.data(), out);
}

// ---------------------------------------------------------------------------
#include "../../elliptic_b200/csrc/sha2.cuh"
extern "C" void he_sha512(const uint8_t* p, size_t n, uint8_t* out) { sha512_ctx c; sha512_init(&c); sha512_update(&c, p, n); sha512_final(&c, out); }
extern "C" void he_sha256(const uint8_t* p, size_t n, uint8_t* out) { sha256_ctx c; sha256_init(&c); sha256_update(&c, p, n); sha256_final(&c, out); }
extern "C" void he_hmac256(const uint8_t* k, const uint8_t* a, size_t na, const uint8_t* b, size_t nb, uint8_t* out) { hmac_sha256(k, a, na, b, nb, 0, 0, out); }
extern "C" void he_ed25519_hash(size_t N, const uint8_t* R, const uint8_t* A, const uint8_t* msgs, const u64* off, uint8_t* h) {
  for (size_t i = 0; i < N; i++) ed25519_hash_item(i, R, A, msgs, off, h);
}

// ---------------------------------------------------------------------------
#include "../../elliptic_b200/csrc/ecdsa_k256_sign.cuh"
extern "C" void he_fe_inv_chain(const u32* a, u32* out) { store_fe(out, fe_inv_chain(load_fe(a))); }
extern "C" void he_sign(size_t N, const uint8_t* e, const uint8_t* priv, u32 canonical, const u32* gtab,
                        uint8_t* r, uint8_t* s, uint8_t* recid, uint8_t* status) {
  for (size_t i = 0; i < N; i++) status[i] = k256_sign_item(i, e, priv, canonical, gtab, r, s, recid);
}

// ---------------------------------------------------------------------------
// exact replay of the reference's non-GLV wNAF schedule (p256 / p384)
#include "../../elliptic_b200/csrc/ecdsa_sw_replay.cuh"
template <class C>
static std::vector<u32>& sw_replay_tab() {
  static std::vector<u32> tab;
  if (tab.empty()) {
    tab.resize(SWReplay<C>::TAB_WORDS);
    for (int t = 0; t < SWReplay<C>::NAF_PTS; t++) SWReplay<C>::tab_entry(t, &tab[2 * C::N * t]);
  }
  return tab;
}
template <class C>
static void sw_replay_jmuladd_host(const u32* u1, const u32* u2, const u32* qx, const u32* qy, u32* xyz) {
  typedef SWReplay<C> R;
  typename R::aff Q;
  typename R::fe t;
  copy_n<C::N>(t.v, qx); Q.x = R::F::to_mont(t);
  copy_n<C::N>(t.v, qy); Q.y = R::F::to_mont(t);
  typename R::jac a = R::jmul_add(u1, u2, Q, sw_replay_tab<C>().data());
  copy_n<C::N>(xyz, R::F::from_mont(a.x).v);
  copy_n<C::N>(xyz + C::N, R::F::from_mont(a.y).v);
  copy_n<C::N>(xyz + 2 * C::N, R::F::from_mont(a.z).v);
}
extern "C" void he_sw_replay_jmuladd(int curve, const u32* u1, const u32* u2, const u32* qx, const u32* qy, u32* xyz) {
  if (curve == 2) sw_replay_jmuladd_host<P256>(u1, u2, qx, qy, xyz);
  else if (curve == 6) sw_replay_jmuladd_host<P521>(u1, u2, qx, qy, xyz);
  else if (curve == 7) sw_replay_jmuladd_host<P192>(u1, u2, qx, qy, xyz);
  else if (curve == 8) sw_replay_jmuladd_host<P224>(u1, u2, qx, qy, xyz);
  else sw_replay_jmuladd_host<P384>(u1, u2, qx, qy, xyz);
}
extern "C" void he_sw_replay_verify(int curve, size_t N, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* pub, uint8_t* status) {
  for (size_t i = 0; i < N; i++)
    status[i] = curve == 2 ? SWReplay<P256>::verify_item(i, e, r, s, pub, sw_replay_tab<P256>().data())
              : curve == 6 ? SWReplay<P521>::verify_item(i, e, r, s, pub, sw_replay_tab<P521>().data())
              : curve == 7 ? SWReplay<P192>::verify_item(i, e, r, s, pub, sw_replay_tab<P192>().data())
              : curve == 8 ? SWReplay<P224>::verify_item(i, e, r, s, pub, sw_replay_tab<P224>().data())
                           : SWReplay<P384>::verify_item(i, e, r, s, pub, sw_replay_tab<P384>().data());
}

// ---------------------------------------------------------------------------
// Point.mul / mulAdd bodies (secp256k1): fast path + exact replay
extern "C" void he_mul_add(size_t N, const uint8_t* k1, const uint8_t* k2, const uint8_t* pts, const u32* gtab,
                           uint8_t* out, uint8_t* status) {
  std::vector<u32> ws((size_t)PREP_WORDS * N), qtab((size_t)QTAB_WORDS * N);
  for (size_t i = 0; i < N; i++) prep_scalars_item(i, N, k1, k2, ws.data());
  for (size_t i = 0; i < N; i++) {
    status[i] = mul_add_item(i, N, pts, ws.data(), gtab, qtab.data(), out);
    if (status[i] == ST_NEEDS_HOST) status[i] = rp_mul_add_item(i, k1, k2, pts, replay_tab().data(), out);
  }
}
extern "C" void he_mul_g(size_t N, const uint8_t* k, const u32* gtab, uint8_t* out, uint8_t* status) {
  for (size_t i = 0; i < N; i++) status[i] = k256_mul_g_item(i, k, gtab, out);
}

// Point.mul / mulAdd on p256 / p384: fast path + exact replay
template <class C>
static void sw_mul_add_host(size_t N, const uint8_t* k1, const uint8_t* k2, const uint8_t* pts, uint8_t* out, uint8_t* status) {
  typedef SW<C> W;
  // reuse the verify helper's table by running it once on zero items
  static std::vector<u32> gtab;
  if (gtab.empty()) {
    gtab.resize((size_t)W::GWINDOWS * W::GENTRIES * 2 * W::N);
    for (int j = 0; j < W::GWINDOWS; j++)
      for (int i = 0; i < W::GENTRIES; i++) W::gtab_entry(j, i, &gtab[((size_t)j * W::GENTRIES + i) * 2 * W::N]);
  }
  if (!pts) {
    for (size_t i = 0; i < N; i++) status[i] = W::mul_g_item(i, k2, gtab.data(), out);
    return;
  }
  std::vector<u32> ws((size_t)W::PREP_WORDS * N), qtab((size_t)W::QTAB_WORDS * N);
  for (size_t i = 0; i < N; i++) W::prep_scalars_item(i, N, k1, k2, ws.data());
  for (size_t i = 0; i < N; i++) {
    status[i] = W::mul_add_item(i, N, pts, ws.data(), gtab.data(), qtab.data(), out);
    if (status[i] == 4) status[i] = SWReplay<C>::mul_add_item(i, k1, k2, pts, sw_replay_tab<C>().data(), out);
  }
}
extern "C" void he_sw_mul_add(int curve, size_t N, const uint8_t* k1, const uint8_t* k2, const uint8_t* pts, uint8_t* out, uint8_t* status) {
  if (curve == 2) sw_mul_add_host<P256>(N, k1, k2, pts, out, status);
  else if (curve == 6) sw_mul_add_host<P521>(N, k1, k2, pts, out, status);
  else if (curve == 7) sw_mul_add_host<P192>(N, k1, k2, pts, out, status);
  else if (curve == 8) sw_mul_add_host<P224>(N, k1, k2, pts, out, status);
  else sw_mul_add_host<P384>(N, k1, k2, pts, out, status);
}

// ---------------------------------------------------------------------------
// two-kernel signing pipeline (nonce -> finish -> flagged items through the literal loop)
#include "../../elliptic_b200/csrc/ecdsa_k256_sign_fast.cuh"
extern "C" void he_sign_fast(size_t N, const uint8_t* e, const uint8_t* priv, u32 canonical, const u32* gtab,
                             uint8_t* r, uint8_t* s, uint8_t* recid, uint8_t* status, int force_slow_every) {
  std::vector<u32> ws((size_t)SIGN_WS_WORDS * N), scratch((size_t)SIGN_SCRATCH_WORDS * N);
  for (size_t i = 0; i < N; i++) k256_sign_nonce_item(i, N, e, priv, gtab, ws.data(), status);
  if (force_slow_every)      // exercise the fallback: pretend the nonce kernel flagged every k-th item
    for (size_t i = 0; i < N; i += force_slow_every) {
      status[i] = ST_NEEDS_HOST;
      for (int w = 0; w < 8; w++) { ws[(size_t)(16 + w) * N + i] = w == 0; ws[(size_t)(24 + w) * N + i] = w == 0; }
    }
  size_t T = (N + PREP_BATCH - 1) / PREP_BATCH;
  for (size_t t = 0; t < T; t++)
    k256_sign_finish_thread(t, T, N, e, priv, canonical, ws.data(), scratch.data(), r, s, recid, status);
  for (size_t i = 0; i < N; i++)
    if (status[i] == ST_NEEDS_HOST) status[i] = k256_sign_item(i, e, priv, canonical, gtab, r, s, recid);
}
extern "C" void he_drbg_first_k(const uint8_t* priv, const uint8_t* msg, uint8_t* out) {
  u32 d[8], m[8], k[8];
  for (int w = 0; w < 8; w++) {
    d[w] = ((u32)priv[4 * w] << 24) | ((u32)priv[4 * w + 1] << 16) | ((u32)priv[4 * w + 2] << 8) | priv[4 * w + 3];
    m[w] = ((u32)msg[4 * w] << 24) | ((u32)msg[4 * w + 1] << 16) | ((u32)msg[4 * w + 2] << 8) | msg[4 * w + 3];
  }
  drbg_first_k(d, m, k);
  for (int w = 0; w < 8; w++) for (int b = 0; b < 4; b++) out[4 * w + b] = (uint8_t)(k[w] >> (24 - 8 * b));
}

// ---------------------------------------------------------------------------
// DER signature import (ec/signature.js:73-134)
#include "../../elliptic_b200/csrc/der_sig.cuh"
extern "C" int he_der_import(const uint8_t* data, size_t n, size_t len, uint8_t* r, uint8_t* s) {
  return der_import(data, n, len, r, s) ? 1 : 0;
}

// ---------------------------------------------------------------------------
// carry-free 9 x 29-bit field (fq_pm.cuh): op 0 mul, 1 sqr, 2 add, 3 sub, 4 neg, 6 canon, 9 weak(7a), 10 is_zero;
// operands/results are packed 8 x 32 words.  which: 0 = secp256k1 prime, 1 = 2^255 - 19.
template <class P>
static void fq_op_t(int op, const u32* a, const u32* b, u32* out) {
  fq<P, 1> A = fq_from_words<P>(a), B = fq_from_words<P>(b), R;
  switch (op) {
    case 0: R = fq_mul(A, B); break;
    case 1: R = fq_sqr(A); break;
    case 2: R = fq_weak(fq_add(A, B)); break;
    case 3: R = fq_weak(fq_sub(A, B)); break;
    case 4: R = fq_weak(fq_neg(A)); break;
    case 6: R = A; break;
    case 9: R = fq_weak(fq_sub(fq_mul_int<3>(A), fq_mul_int<2>(B))); break;       // magnitude 3 + 2 + 1 + ... = 6
    case 10: { bool z = fq_is_zero(fq_sub(A, B)); for (int i = 0; i < 8; i++) out[i] = 0; out[0] = z; return; }
    case 11: R = fq_mul(fq_add(fq_add(A, A), A), fq_add(B, B)); break;            // magnitudes 3 x 2
    case 12: R = fq_sqr(fq_add(A, B)); break;                                     // magnitude 2
    case 13: R = fq_weak(fq_cneg(A, b[0] & 1)); break;
    default: R = A;
  }
  fq_to_words<P>(out, fq_canon(R));
}
extern "C" void he_fq_op(int which, int op, const u32* a, const u32* b, u32* out) {
  if (which == 0) fq_op_t<PmK256>(op, a, b, out); else fq_op_t<Pm25519>(op, a, b, out);
}
// group law on the carry-free field against the packed-field code: op 0 dbl, 1 madd
extern "C" void he_gq_op(int op, const u32* jac24, const u32* aff16, u32* out24) {
  ge_jac p; ge_aff q;
  for (int i = 0; i < 8; i++) { p.x.v[i] = jac24[i]; p.y.v[i] = jac24[8 + i]; p.z.v[i] = jac24[16 + i]; q.x.v[i] = aff16[i]; q.y.v[i] = aff16[8 + i]; }
  gq_jac P = gq_from_jac(p), R;
  gq_aff Qa; Qa.x = fqk_from_fe(q.x); Qa.y = fqk_from_fe(q.y);
  R = op == 0 ? gq_dbl(P) : gq_madd(P, Qa);
  ge_jac r = gq_to_jac(R);
  for (int i = 0; i < 8; i++) { out24[i] = r.x.v[i]; out24[8 + i] = r.y.v[i]; out24[16 + i] = r.z.v[i]; }
}

// ---------------------------------------------------------------------------
// EC.sign on p256 / p384: nonce -> finish -> flagged items through the literal loop
#include "../../elliptic_b200/csrc/ecdsa_sw_sign.cuh"
template <class SG, class C>
static void sw_sign_host(size_t N, const uint8_t* e, const uint8_t* priv, u32 canonical, uint8_t* r, uint8_t* s,
                         uint8_t* recid, uint8_t* status, int force_slow_every) {
  typedef SW<C> W;
  static std::vector<u32> gtab;
  if (gtab.empty()) {
    gtab.resize((size_t)W::GWINDOWS * W::GENTRIES * 2 * W::N);
    for (int j = 0; j < W::GWINDOWS; j++)
      for (int i = 0; i < W::GENTRIES; i++) W::gtab_entry(j, i, &gtab[((size_t)j * W::GENTRIES + i) * 2 * W::N]);
  }
  std::vector<u32> ws((size_t)SG::WS_WORDS * N), scratch((size_t)SG::SCRATCH_WORDS * N);
  for (size_t i = 0; i < N; i++) SG::nonce_item(i, N, e, priv, gtab.data(), ws.data(), status);
  if (force_slow_every)
    for (size_t i = 0; i < N; i += force_slow_every) {
      status[i] = 4;
      typename W::fe one = W::F::one();
      for (int w = 0; w < W::N; w++) { ws[(size_t)(2 * W::N + w) * N + i] = one.v[w]; ws[(size_t)(3 * W::N + w) * N + i] = w == 0; }
    }
  size_t T = (N + SG::BATCH - 1) / SG::BATCH;
  for (size_t t = 0; t < T; t++) SG::finish_thread(t, T, N, e, priv, canonical, ws.data(), scratch.data(), r, s, recid, status);
  for (size_t i = 0; i < N; i++)
    if (status[i] == 4) status[i] = SG::slow_item(i, e, priv, canonical, gtab.data(), r, s, recid);
}
extern "C" void he_sw_sign(int curve, size_t N, const uint8_t* e, const uint8_t* priv, u32 canonical, uint8_t* r, uint8_t* s,
                           uint8_t* recid, uint8_t* status, int force_slow_every) {
  if (curve == 2) sw_sign_host<SWSign<P256, Sha256W>, P256>(N, e, priv, canonical, r, s, recid, status, force_slow_every);
  else if (curve == 6) sw_sign_host<SWSign<P521, Sha512W>, P521>(N, e, priv, canonical, r, s, recid, status, force_slow_every);
  else if (curve == 7) sw_sign_host<SWSign<P192, Sha256W>, P192>(N, e, priv, canonical, r, s, recid, status, force_slow_every);
  else if (curve == 8) sw_sign_host<SWSign<P224, Sha256W>, P224>(N, e, priv, canonical, r, s, recid, status, force_slow_every);
  else sw_sign_host<SWSign<P384, Sha384W>, P384>(N, e, priv, canonical, r, s, recid, status, force_slow_every);
}

// EC.recoverPubKey on p256 / p384 / p521
template <class C>
static void sw_recover_host(size_t N, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* recid,
                            uint8_t* out, uint8_t* status) {
  typedef SW<C> W;
  static std::vector<u32> gtab;
  if (gtab.empty()) {
    gtab.resize((size_t)W::GWINDOWS * W::GENTRIES * 2 * W::N);
    for (int j = 0; j < W::GWINDOWS; j++)
      for (int i = 0; i < W::GENTRIES; i++) W::gtab_entry(j, i, &gtab[((size_t)j * W::GENTRIES + i) * 2 * W::N]);
  }
  std::vector<u32> ws((size_t)W::PREP_WORDS * N), qtab((size_t)W::QTAB_WORDS * N);
  for (size_t i = 0; i < N; i++) W::prep_recover_item(i, N, e, r, s, ws.data());
  for (size_t i = 0; i < N; i++) status[i] = W::recover_item(i, N, r, recid, ws.data(), gtab.data(), qtab.data(), out);
}
extern "C" void he_sw_recover(int curve, size_t N, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* recid,
                              uint8_t* out, uint8_t* status) {
  if (curve == 2) sw_recover_host<P256>(N, e, r, s, recid, out, status);
  else if (curve == 6) sw_recover_host<P521>(N, e, r, s, recid, out, status);
  else if (curve == 7) sw_recover_host<P192>(N, e, r, s, recid, out, status);
  else if (curve == 8) sw_recover_host<P224>(N, e, r, s, recid, out, status);
  else sw_recover_host<P384>(N, e, r, s, recid, out, status);
}

// Red.sqrt restatement (SW<C>::sqrt_ref): plain a in, plain root out; returns the status (0 ok, 5 assertion)
template <class C>
static int sw_sqrt_host(const u32* a, u32* out) {
  typedef SW<C> W;
  typename W::fe t, y = W::F::zero();
  copy_n<C::N>(t.v, a);
  int st = W::sqrt_ref(W::F::to_mont(t), &y);
  copy_n<C::N>(out, W::F::from_mont(y).v);
  return st;
}
extern "C" int he_sw_sqrt(int curve, const u32* a, u32* out) {
  if (curve == 2) return sw_sqrt_host<P256>(a, out);
  if (curve == 6) return sw_sqrt_host<P521>(a, out);
  if (curve == 7) return sw_sqrt_host<P192>(a, out);
  if (curve == 8) return sw_sqrt_host<P224>(a, out);
  return sw_sqrt_host<P384>(a, out);
}

// Coordinate-field operations of a short curve on plain integers (to_mont / op / from_mont):
// op 0 mul, 1 sqr, 2 add, 3 sub, 4 neg, 6 reduce (toRed), 7 inv.  a, b, out: C::N words.
template <class C>
static void sw_fe_op_t(int op, const u32* a, const u32* b, u32* out) {
  typedef typename SW<C>::F F;
  constexpr int NL = C::N;
  typename F::fe A = F::to_mont(load_fe_n<NL>(a)), B = F::to_mont(load_fe_n<NL>(b)), R;
  switch (op) {
    case 0: R = F::mul(A, B); break;
    case 1: R = F::sqr(A); break;
    case 2: R = F::add(A, B); break;
    case 3: R = F::sub(A, B); break;
    case 4: R = F::neg(A); break;
    case 7: R = F::inv(A); break;
    case 8: R = F::template mul_k<3>(A, B); break;
    case 9: R = F::template mul_k<4>(A, B); break;
    case 10: R = F::template sqr_k<8>(A); break;
    default: R = A;
  }
  store_fe_n<NL>(out, F::from_mont(R));
}
// the word-level reduction alone: c = lo + hi 2^(32 NL) (any words) -> canonical residue
template <class RED, class PF, int NL>
static void solinas_reduce_t(const u32* c, int scale, u32* out) {
  u32 t[2 * NL + 2], p[NL];
  for (int i = 0; i < 2 * NL + 2; i++) t[i] = i < 2 * NL ? c[i] : 0;
  PF::mod(p);
  if (scale != 1) RED::reduce_scaled(out, t, p, scale);
  else RED::reduce(out, t, p);
}
extern "C" void he_solinas_reduce(int curve, const u32* c, int scale, u32* out) {
  if (curve == 2) solinas_reduce_t<RedP256, P256_FP, 8>(c, scale, out);
  else solinas_reduce_t<RedP384, P384_FP, 12>(c, scale, out);
}
extern "C" void he_sw_fe_op(int curve, int op, const u32* a, const u32* b, u32* out) {
  switch (curve) {
    case 2: sw_fe_op_t<P256>(op, a, b, out); break;
    case 3: sw_fe_op_t<P384>(op, a, b, out); break;
    case 6: sw_fe_op_t<P521>(op, a, b, out); break;
    case 7: sw_fe_op_t<P192>(op, a, b, out); break;
    default: sw_fe_op_t<P224>(op, a, b, out);
  }
}

// ---------------------------------------------------------------------------
// EC.sign options (k, pers) and EC.genKeyPair({entropy, pers}) through the kernel bodies.
// mode 0: caller nonces kgiven (one attempt; 10 = the reference would ask for k(iter + 1)), mode 1: pers
template <class SG, class C>
static void sw_sign_opt_host(int mode, size_t N, const uint8_t* e, const uint8_t* priv, const uint8_t* kgiven, const uint8_t* pers, int np,
                             u32 canonical, uint8_t* r, uint8_t* s, uint8_t* recid, uint8_t* status) {
  typedef SW<C> W;
  static std::vector<u32> gtab;
  if (gtab.empty()) {
    gtab.resize((size_t)W::GWINDOWS * W::GENTRIES * 2 * W::N);
    for (int j = 0; j < W::GWINDOWS; j++)
      for (int i = 0; i < W::GENTRIES; i++) W::gtab_entry(j, i, &gtab[((size_t)j * W::GENTRIES + i) * 2 * W::N]);
  }
  if (mode == 1) {
    for (size_t i = 0; i < N; i++) status[i] = SG::slow_item_pers(i, e, priv, pers, np, canonical, gtab.data(), r, s, recid);
    return;
  }
  std::vector<u32> ws((size_t)SG::WS_WORDS * N), scratch((size_t)SG::SCRATCH_WORDS * N);
  for (size_t i = 0; i < N; i++) SG::nonce_item(i, N, e, priv, gtab.data(), ws.data(), status, kgiven);
  size_t T = (N + SG::BATCH - 1) / SG::BATCH;
  for (size_t t = 0; t < T; t++) SG::finish_thread(t, T, N, e, priv, canonical, ws.data(), scratch.data(), r, s, recid, status);
  for (size_t i = 0; i < N; i++) if (status[i] == 4) status[i] = 10;
}
extern "C" void he_sign_opt(int curve, int mode, size_t N, const uint8_t* e, const uint8_t* priv, const uint8_t* kgiven, const uint8_t* pers,
                            int np, u32 canonical, const u32* k256_gtab, uint8_t* r, uint8_t* s, uint8_t* recid, uint8_t* status) {
  if (curve == 1) {
    if (mode == 1) { for (size_t i = 0; i < N; i++) status[i] = k256_sign_item_pers(i, e, priv, pers, np, canonical, k256_gtab, r, s, recid); return; }
    std::vector<u32> ws((size_t)SIGN_WS_WORDS * N), scratch((size_t)SIGN_SCRATCH_WORDS * N);
    for (size_t i = 0; i < N; i++) k256_sign_nonce_item(i, N, e, priv, k256_gtab, ws.data(), status, kgiven);
    size_t T = (N + PREP_BATCH - 1) / PREP_BATCH;
    for (size_t t = 0; t < T; t++) k256_sign_finish_thread(t, T, N, e, priv, canonical, ws.data(), scratch.data(), r, s, recid, status);
    for (size_t i = 0; i < N; i++) if (status[i] == ST_NEEDS_HOST) status[i] = 10;
  }
  else if (curve == 2) sw_sign_opt_host<SWSign<P256, Sha256W>, P256>(mode, N, e, priv, kgiven, pers, np, canonical, r, s, recid, status);
  else if (curve == 3) sw_sign_opt_host<SWSign<P384, Sha384W>, P384>(mode, N, e, priv, kgiven, pers, np, canonical, r, s, recid, status);
  else if (curve == 6) sw_sign_opt_host<SWSign<P521, Sha512W>, P521>(mode, N, e, priv, kgiven, pers, np, canonical, r, s, recid, status);
  else if (curve == 7) sw_sign_opt_host<SWSign<P192, Sha256W>, P192>(mode, N, e, priv, kgiven, pers, np, canonical, r, s, recid, status);
  else sw_sign_opt_host<SWSign<P224, Sha256W>, P224>(mode, N, e, priv, kgiven, pers, np, canonical, r, s, recid, status);
}
extern "C" void he_keygen(int curve, size_t N, const uint8_t* entropy, int ne, const uint8_t* pers, int np, uint8_t* out_priv, uint8_t* status) {
  for (size_t i = 0; i < N; i++) {
    if (curve == 1) status[i] = k256_keygen_item(i, entropy, ne, pers, np, out_priv);
    else if (curve == 2) status[i] = SWSign<P256, Sha256W>::keygen_item(i, entropy, ne, pers, np, out_priv);
    else if (curve == 3) status[i] = SWSign<P384, Sha384W>::keygen_item(i, entropy, ne, pers, np, out_priv);
    else if (curve == 6) status[i] = SWSign<P521, Sha512W>::keygen_item(i, entropy, ne, pers, np, out_priv);
    else if (curve == 7) status[i] = SWSign<P192, Sha256W>::keygen_item(i, entropy, ne, pers, np, out_priv);
    else status[i] = SWSign<P224, Sha256W>::keygen_item(i, entropy, ne, pers, np, out_priv);
  }
}

// ---------------------------------------------------------------------------
// the `ec` API over ed25519 (ed25519_ec.cuh)
#include "../../elliptic_b200/csrc/ed25519_ec.cuh"
extern "C" void he_ed_ec_verify(size_t N, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* pub, int fmt, uint8_t* status) {
  const std::vector<u32>& gtab = ed_host_gtab();
  std::vector<u32> atab((size_t)ED_ATAB_WORDS * N);
  std::vector<uint8_t> xy(64 * N), pre(N, 0);
  const uint8_t* pxy = pub;
  if (fmt) {
    for (size_t i = 0; i < N; i++) pre[i] = ed_ec_decode_pub(pub + (fmt == 1 ? 65 : 33) * i, (u32)fmt, xy.data() + 64 * i);
    pxy = xy.data();
  }
  for (size_t i = 0; i < N; i++) status[i] = ed_ec_verify_item(i, e, r, s, pxy, fmt ? pre.data() : nullptr, gtab.data(), atab.data());
}
extern "C" void he_ed_ec_sign(size_t N, const uint8_t* e, const uint8_t* priv, const uint8_t* kgiven, const uint8_t* pers, int np,
                              u32 canonical, uint8_t* r, uint8_t* s, uint8_t* recid, uint8_t* status) {
  const std::vector<u32>& gtab = ed_host_gtab();
  for (size_t i = 0; i < N; i++) status[i] = ed_ec_sign_item(i, e, priv, kgiven, pers, np, canonical, gtab.data(), r, s, recid);
}
extern "C" void he_ed_ec_keygen(size_t N, const uint8_t* entropy, int ne, const uint8_t* pers, int np, uint8_t* out_priv, uint8_t* status) {
  for (size_t i = 0; i < N; i++) status[i] = ed_ec_keygen_item(i, entropy, ne, pers, np, out_priv);
}
extern "C" void he_ed_ec_mul_add(size_t N, const uint8_t* k1, const uint8_t* k2, const uint8_t* pts, int derive, uint8_t* out, uint8_t* status) {
  const std::vector<u32>& gtab = ed_host_gtab();
  std::vector<u32> atab((size_t)ED_ATAB_WORDS * N);
  for (size_t i = 0; i < N; i++) status[i] = ed_ec_mul_add_item(i, k1, k2, pts, derive != 0, gtab.data(), atab.data(), out);
}
extern "C" void he_x25519_mul(size_t N, const uint8_t* k, const uint8_t* px, uint8_t* out, uint8_t* status) {
  for (size_t i = 0; i < N; i++) status[i] = x25519_mul_item(i, k, px, out);
}

// ---------------------------------------------------------------------------
// run-time short curves (sw_runtime.cuh): the host fills the parameter block exactly like rt_make in eb200.cu
#include "../../elliptic_b200/csrc/sw_runtime.cuh"
extern "C" void he_rt_item(int op, const u32* p8, const u32* r1, const u32* r2, const u32* a_m, const u32* b_m, u32 n0inv, u32 len,
                           const uint8_t* k1, const uint8_t* p1, const uint8_t* k2, const uint8_t* p2, u32 klen, uint8_t* out, uint8_t* status) {
  RtCurve<8> C;
  for (int i = 0; i < 8; i++) { C.p[i] = p8[i]; C.r1[i] = r1[i]; C.r2[i] = r2[i]; C.a[i] = a_m[i]; C.b[i] = b_m[i]; }
  C.n0inv = n0inv; C.len = len; C.a_is_zero = 0;
  status[0] = RtG<8>::item(op, 0, k1, p1, k2, p2, klen, out, C);
}

// ---------------------------------------------------------------------------
// the chunk plan of the pipelined host calls (pure host logic of the product library)
#include "../../elliptic_b200/csrc/chunk_plan.h"
extern "C" int he_chunk_plan(size_t n, unsigned long long* lo /* EB_MAX_CHUNKS + 2 */, unsigned long long* max_m) {
  ChunkPlan P = make_plan(n);
  for (int i = 0; i <= P.chunks; i++) lo[i] = P.lo[i];
  *max_m = P.max_m;
  return P.chunks;
}
