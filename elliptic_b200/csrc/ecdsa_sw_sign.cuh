// ecdsa_sw_sign.cuh -- EC.prototype.sign (lib/elliptic/ec/index.js:110-186) on p256 (SHA-256) and p384 (SHA-384),
// the curves' default hashes (curves.js:73-107), as the same two-kernel pipeline as secp256k1
// (ecdsa_k256_sign_fast.cuh): a nonce kernel (RFC 6979 HMAC-DRBG in registers + k*G from the fixed table, left
// Jacobian) and a finish kernel (Montgomery's trick over 16 signatures for JPoint.toP and k^-1, then r, s,
// recoveryParam, `canonical`).  Items that need the reference's retry loop (k outside [2, n-2], r = 0, s = 0;
// ec/index.js:158-170) are flagged and redone by slow_item, which runs that loop literally.
#pragma once
#include "ecdsa_sw_body.cuh"
#include "hmac_drbg_w.cuh"

namespace eb {

// Nonce source: HMAC-DRBG seeded with (private key, message), yielding candidate k's as little-endian limbs.
// WORDS = true: key, message and digest have the same length (p256 / SHA-256, p384 / SHA-384) -> the register-only
// generator; false: any lengths (p192, p224, p521) -> the byte-stream generator, plus _truncateToN(k, true).
template <class C, class H, bool WORDS>
struct SignDrbg;

template <class C, class H>
struct SignDrbg<C, H, true> {
  typedef typename H::W HW;
  static constexpr int N = C::N;
  HmacDrbgW<H> g;
  EB_HD void init(const uint8_t* priv, const uint8_t* msg) {
    constexpr int B = H::WB / 8;
    HW dw[H::D], ew[H::D];
    for (int i = 0; i < H::D; i++) {
      HW a = 0, b = 0;
      for (int k = 0; k < B; k++) { a = (a << 8) | priv[B * i + k]; b = (b << 8) | msg[B * i + k]; }
      dw[i] = a; ew[i] = b;
    }
    g.init(dw, ew);
  }
  EB_HD void next_k(u32* k) {
    constexpr int B = H::WB / 32;                // 32-bit limbs per hash word
    HW kw[H::D];
    g.generate(kw);
    for (int i = 0; i < H::D; i++)
      for (int j = 0; j < B; j++) k[N - 1 - (B * i + j)] = (u32)(kw[i] >> (H::WB - 32 * (j + 1)));
  }
};

template <class C, class H>
struct SignDrbg<C, H, false> {
  static constexpr int N = C::N;
  HmacDrbgB<H> g;
  EB_HD void init(const uint8_t* priv, const uint8_t* msg, const uint8_t* pers = nullptr, int np = 0) {
    g.init(priv, C::LEN, msg, C::LEN, pers, np);
  }
  // _truncateToN(k, true) (ec/index.js:81-108, BN input) of a LEN-byte big-endian value: the shift is taken from
  // the VALUE's byte length
  static EB_HD void truncate_k(u32* k, const uint8_t* kb) {
    load_be_len<N>(k, kb, C::LEN);
    int top = 0;
    while (top < C::LEN && kb[top] == 0) top++;
    int delta = 8 * (C::LEN - top) - C::BITS;
    if (delta > 0) {
      for (int w = 0; w < N; w++) k[w] = (k[w] >> delta) | ((w + 1 < N ? k[w + 1] : 0u) << (32 - delta));
    }
  }
  EB_HD void next_k(u32* k) {
    uint8_t kb[C::LEN];
    g.generate(kb, C::LEN);                                  // drbg.generate(n.byteLength())
    truncate_k(k, kb);
  }
};

template <class C, class H>
struct SWSign {
  typedef SW<C> W;
  typedef typename W::F F;
  typedef typename W::S S;
  typedef typename F::fe fe;
  typedef typename S::fe sc;
  typedef typename W::jac jac;
  typedef typename W::aff aff;
  static constexpr bool WORDS = (C::LEN == 4 * C::N) && (C::LEN == H::D * H::WB / 8);
  typedef SignDrbg<C, H, WORDS> Drbg;
  static constexpr int N = C::N;
  static constexpr int WS_WORDS = 4 * N;        // X, Y, Z, k  (word-major SoA)
  static constexpr int SCRATCH_WORDS = 2 * N;   // prefix products of Z and k
  static constexpr int BATCH = 16;

  // ec/index.js:158-159
  static EB_HD bool k_in_range(const u32* k) {
    u32 nmod[N], ns1[N], one[N];
    W::n_limbs(nmod);
    for (int w = 0; w < N; w++) one[w] = 0;
    one[0] = 1;
    sub_n<N>(ns1, nmod, one);
    bool le1 = k[0] <= 1;
    for (int w = 1; w < N; w++) le1 = le1 && k[w] == 0;
    return !le1 && !geq_n<N>(k, ns1);
  }
  // k*G for 0 < k < n from the fixed table, Jacobian
  static EB_HD jac mul_g_jac(const u32* k, const u32* gtab) {
    u32 nmod[N], kv[N];
    W::n_limbs(nmod);
    copy_n<N>(kv, k);
    bool negg = (kv[0] & 1) == 0;
    if (negg) sub_n<N>(kv, nmod, kv);
    u32 m[N];
    for (int w = 0; w < N; w++) m[w] = (kv[w] >> 1) | ((w < N - 1 ? kv[w + 1] : 0u) << 31);
    jac acc = W::infinity();
    for (int j = 0; j < W::GWINDOWS; j++) {
      u32 chunk = W::extract(m, 1, 0, 0, W::GW * j, W::GW);
      const u32 half = 1u << (W::GW - 1);
      bool dneg = (j != W::GWINDOWS - 1) && (chunk < half);
      u32 idx = (j == W::GWINDOWS - 1) ? (chunk & (half - 1)) : (dneg ? half - 1 - chunk : chunk - half);
      const u32* ent = gtab + ((size_t)j * W::GENTRIES + idx) * 2 * N;
      aff P;
      P.x = load_fe_n<N>(ent);
      P.y = load_fe_n<N>(ent + N);
      P.y = F::cmov(P.y, F::neg(P.y), dneg != negg);
      acc = W::madd(acc, P);
    }
    return acc;
  }

  // kgiven != NULL: the caller's own nonce for this attempt (options.k, ec/index.js:154-157) instead of the DRBG
  static EB_HD void nonce_item(size_t i, size_t cnt, const uint8_t* e, const uint8_t* priv, const u32* gtab, u32* ws,
                               uint8_t* status, const uint8_t* kgiven = nullptr) {
    u32 k[N];
    if (kgiven) {
      SignDrbg<C, H, false>::truncate_k(k, kgiven + C::LEN * i);
    } else {
      Drbg g;
      g.init(priv + C::LEN * i, e + C::LEN * i);
      g.next_k(k);
    }
    bool ok = k_in_range(k);
    jac acc = W::infinity();
    if (ok) acc = mul_g_jac(k, gtab);
    else {
      acc.z = F::one();
      for (int w = 0; w < N; w++) k[w] = w == 0;
    }
    for (int w = 0; w < N; w++) {
      ws[(size_t)w * cnt + i] = acc.x.v[w];
      ws[(size_t)(N + w) * cnt + i] = acc.y.v[w];
      ws[(size_t)(2 * N + w) * cnt + i] = acc.z.v[w];
      ws[(size_t)(3 * N + w) * cnt + i] = k[w];
    }
    status[i] = ok ? 1 : 4;
  }

  // r, s, recid from an affine k*G (Montgomery form) and k^-1 (Montgomery form); false = retry needed
  static EB_HD bool finish_one(size_t i, const fe& ax_m, const fe& ay_m, const sc& kinv_m, const uint8_t* e,
                               const uint8_t* priv, u32 canonical, uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid) {
    u32 nmod[N];
    W::n_limbs(nmod);
    fe ax = F::from_mont(ax_m), ay = F::from_mont(ay_m);
    sc r;
    copy_n<N>(r.v, ax.v);
    bool xr_differ = geq_n<N>(r.v, nmod);
    if (xr_differ) sub_n<N>(r.v, r.v, nmod);            // kpX.umod(n): p < 2n on these curves
    if (is_zero_n<N>(r.v)) return false;
    sc ev, dv;
    W::ldb(ev.v, e + C::LEN * i);
    W::ldb(dv.v, priv + C::LEN * i);
    sc rd = S::mul(r, S::to_mont(dv));                  // r d mod n (plain)
    sc t = S::add(rd, ev);                              // e < n: _truncateToN already subtracted n once
    sc s = S::mul(t, kinv_m);                           // k^-1 (r d + e) mod n
    if (is_zero_n<N>(s.v)) return false;
    u32 rec = (ay.v[0] & 1) | (xr_differ ? 2u : 0u);
    if (canonical) {
      u32 nh[N], d2[N];
      for (int w = 0; w < N; w++) nh[w] = (nmod[w] >> 1) | ((w < N - 1 ? nmod[w + 1] : 0u) << 31);
      if (sub_n<N>(d2, nh, s.v) != 0) { sub_n<N>(s.v, nmod, s.v); rec ^= 1; }
    }
    W::stb(out_r + C::LEN * i, r.v);
    W::stb(out_s + C::LEN * i, s.v);
    out_recid[i] = (uint8_t)rec;
    return true;
  }

  static EB_HD void finish_thread(size_t tid, size_t T, size_t cnt, const uint8_t* e, const uint8_t* priv, u32 canonical,
                                  const u32* ws, u32* scratch, uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid,
                                  uint8_t* status) {
    fe zprod = F::one();
    sc kprod = S::one();
    int n_items = 0;
    for (int j = 0; j < BATCH; j++) {
      size_t i = tid + (size_t)j * T;
      if (i >= cnt) break;
      n_items = j + 1;
      fe z; sc k;
      for (int w = 0; w < N; w++) { z.v[w] = ws[(size_t)(2 * N + w) * cnt + i]; k.v[w] = ws[(size_t)(3 * N + w) * cnt + i]; }
      for (int w = 0; w < N; w++) { scratch[(size_t)w * cnt + i] = zprod.v[w]; scratch[(size_t)(N + w) * cnt + i] = kprod.v[w]; }
      zprod = F::mul(zprod, z);
      kprod = S::mul(kprod, S::to_mont(k));
    }
    if (n_items == 0) return;
    fe zinv_all = F::inv(zprod);
    sc kinv_all = S::inv(kprod);
    for (int j = n_items - 1; j >= 0; j--) {
      size_t i = tid + (size_t)j * T;
      fe x, y, z, zpre; sc k, kpre;
      for (int w = 0; w < N; w++) {
        x.v[w] = ws[(size_t)w * cnt + i]; y.v[w] = ws[(size_t)(N + w) * cnt + i]; z.v[w] = ws[(size_t)(2 * N + w) * cnt + i];
        k.v[w] = ws[(size_t)(3 * N + w) * cnt + i];
        zpre.v[w] = scratch[(size_t)w * cnt + i]; kpre.v[w] = scratch[(size_t)(N + w) * cnt + i];
      }
      fe zi = F::mul(zinv_all, zpre);
      zinv_all = F::mul(zinv_all, z);
      sc kinv = S::mul(kinv_all, kpre);
      kinv_all = S::mul(kinv_all, S::to_mont(k));
      if (status[i] != 1) continue;
      fe zi2 = F::sqr(zi);
      fe ax = F::mul(x, zi2), ay = F::mul(F::mul(y, zi2), zi);
      if (!finish_one(i, ax, ay, kinv, e, priv, canonical, out_r, out_s, out_recid)) status[i] = 4;
    }
  }

  // the same loop with the `pers` option (ec/index.js:143-151): seed = key || msg || pers, byte-stream generator
  static EB_HD uint8_t slow_item_pers(size_t i, const uint8_t* e, const uint8_t* priv, const uint8_t* pers, int np,
                                      u32 canonical, const u32* gtab, uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid) {
    SignDrbg<C, H, false> g;
    g.init(priv + C::LEN * i, e + C::LEN * i, pers, np);
    for (int iter = 0; iter < 128; iter++) {
      u32 k[N];
      g.next_k(k);
      if (!k_in_range(k)) continue;
      aff kp = W::to_aff(mul_g_jac(k, gtab));
      sc km;
      copy_n<N>(km.v, k);
      sc kinv = S::inv(S::to_mont(km));
      if (finish_one(i, kp.x, kp.y, kinv, e, priv, canonical, out_r, out_s, out_recid)) return 1;
    }
    return 0;
  }

  // EC.genKeyPair (ec/index.js:55-79): HmacDRBG(hash, entropy, nonce = n.toArray(), pers); the first candidate
  // priv = BN(generate(n.byteLength())) with priv <= n - 2, plus one.  Writes LEN bytes big-endian.
  static EB_HD uint8_t keygen_item(size_t i, const uint8_t* entropy, int ne, const uint8_t* pers, int np, uint8_t* out_priv) {
    HmacDrbgB<H> g;
    u32 nmod[N], ns2[N], two[N];
    W::n_limbs(nmod);
    for (int w = 0; w < N; w++) two[w] = w == 0 ? 2u : 0u;
    sub_n<N>(ns2, nmod, two);
    uint8_t nb[C::LEN];
    store_be_len<N>(nb, nmod, C::LEN);
    g.init(entropy + (size_t)ne * i, ne, nb, C::LEN, pers, np);
    for (int iter = 0; iter < 65536; iter++) {
      uint8_t kb[C::LEN];
      g.generate(kb, C::LEN);
      bool fits = true;                                     // the raw value may be wider than the limb array holds (p521: 528 bits)
      for (int b = 0; b < C::LEN - 4 * N; b++) fits = fits && kb[b] == 0;
      u32 k[N];
      load_be_len<N>(k, kb, C::LEN);
      if (!fits || (geq_n<N>(k, ns2) && !eq_n<N>(k, ns2))) continue;     // priv.cmp(ns2) > 0
      u32 one[N];
      for (int w = 0; w < N; w++) one[w] = w == 0;
      add_n<N>(k, k, one);
      store_be_len<N>(out_priv + C::LEN * i, k, C::LEN);
      return 1;
    }
    return 0;
  }

  // the literal loop of ec/index.js:153-185 for one flagged item
  static EB_HD uint8_t slow_item(size_t i, const uint8_t* e, const uint8_t* priv, u32 canonical, const u32* gtab,
                                 uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid) {
    Drbg g;
    g.init(priv + C::LEN * i, e + C::LEN * i);
    for (int iter = 0; iter < 128; iter++) {
      u32 k[N];
      g.next_k(k);
      if (!k_in_range(k)) continue;
      aff kp = W::to_aff(mul_g_jac(k, gtab));
      sc km;
      copy_n<N>(km.v, k);
      sc kinv = S::inv(S::to_mont(km));
      if (finish_one(i, kp.x, kp.y, kinv, e, priv, canonical, out_r, out_s, out_recid)) return 1;
    }
    return 0;
  }
};

// secp256k1: EC.sign with the `pers` option and EC.genKeyPair on the byte-stream generator (SHA-256)
EB_HD uint8_t k256_sign_item_pers(size_t i, const uint8_t* e, const uint8_t* priv, const uint8_t* pers, int np, u32 canonical,
                                  const u32* gtab, uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid) {
  u32 ev[8], dv[8];
  load_be<8>(ev, e + 32 * i);
  load_be<8>(dv, priv + 32 * i);
  HmacDrbgB<Sha256W> g;
  g.init(priv + 32 * i, 32, e + 32 * i, 32, pers, np);
  for (int iter = 0; iter < 128; iter++) {
    uint8_t kb[32];
    g.generate(kb, 32);
    u32 k[8];
    load_be<8>(k, kb);
    if (k256_sign_try(i, k, ev, dv, canonical, gtab, out_r, out_s, out_recid)) return ST_TRUE;
  }
  return ST_FALSE;
}
EB_HD uint8_t k256_keygen_item(size_t i, const uint8_t* entropy, int ne, const uint8_t* pers, int np, uint8_t* out_priv) {
  u32 nn[8], ns2[8], two[8] = {2, 0, 0, 0, 0, 0, 0, 0}, one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  K256N::n(nn);
  sub_n<8>(ns2, nn, two);
  uint8_t nb[32];
  store_be<8>(nb, nn);
  HmacDrbgB<Sha256W> g;
  g.init(entropy + (size_t)ne * i, ne, nb, 32, pers, np);
  for (int iter = 0; iter < 65536; iter++) {
    uint8_t kb[32];
    g.generate(kb, 32);
    u32 k[8];
    load_be<8>(k, kb);
    if (geq_n<8>(k, ns2) && !eq_n<8>(k, ns2)) continue;       // priv.cmp(ns2) > 0
    add_n<8>(k, k, one);
    store_be<8>(out_priv + 32 * i, k);
    return ST_TRUE;
  }
  return ST_FALSE;
}

}  // namespace eb
