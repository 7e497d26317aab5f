// ecdsa_k256_sign_fast.cuh -- the common case of EC.prototype.sign (lib/elliptic/ec/index.js:110-186) as a
// two-kernel pipeline; ecdsa_k256_sign.cuh keeps the literal per-item loop and serves the items flagged here.
//
//   nonce kernel  (thread per signature): first HMAC-DRBG output k (hmac-drbg, dist/elliptic.js:8686-8800,
//                 seeded with key || msg as ec/index.js:135-148 does), word-oriented SHA-256 in registers with
//                 the HMAC pad states reused between calls (16 compressions), then R = k*G from the fixed
//                 table, left in Jacobian form.  k outside [2, n-2] (ec/index.js:158-159) -> flagged.
//   finish kernel (thread per 16 signatures): Montgomery's trick inverts the 16 Z's (JPoint.toP,
//                 short.js:516-526) and the 16 k's with one field and one scalar inversion; r = x mod n,
//                 s = k^-1 (e + r d) mod n, recovery param, `canonical`.  r = 0 or s = 0 -> flagged.
// Flagged items (probability ~2^-127) are redone by k256_sign_item, which runs the reference's retry loop.
#pragma once
#include "ecdsa_k256_sign.cuh"

namespace eb {

constexpr int SIGN_WS_WORDS = 32;      // per item: X, Y, Z, k (8 words each), word-major SoA
constexpr int SIGN_SCRATCH_WORDS = 16; // per item: prefix products of Z and of k

EB_HD void eb_prefetch_l2(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

EB_HD void sha256_compress_w(u32* st, const u32* win) {
  u32 w[16];
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = win[i];
  u32 a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
  for (int i = 0; i < 64; i++) {
    u32 wi;
    if (i < 16) wi = w[i];
    else {
      u32 w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      u32 s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      u32 s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
      w[i & 15] = wi;
    }
    u32 S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    u32 ch = (e & f) ^ (~e & g);
    u32 t1 = h + S1 + ch + sha256_k(i) + wi;
    u32 S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    u32 mj = (a & b) ^ (a & c) ^ (b & c);
    u32 t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

EB_HD void sha256_iv(u32* st) {
  const u32 iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
#pragma unroll
  for (int i = 0; i < 8; i++) st[i] = iv[i];
}

struct hmac_pads { u32 in[8], out[8]; };   // SHA-256 states after the ipad / opad block of a 32-byte key

EB_HD void hmac_key(hmac_pads* p, const u32* key) {
  u32 w[16];
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = (i < 8 ? key[i] : 0u) ^ 0x36363636u;
  sha256_iv(p->in);
  sha256_compress_w(p->in, w);
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = (i < 8 ? key[i] : 0u) ^ 0x5c5c5c5cu;
  sha256_iv(p->out);
  sha256_compress_w(p->out, w);
}

// outer hash: opad state + 32-byte inner digest
EB_HD void hmac_outer(const hmac_pads* p, const u32* inner, u32* out) {
  u32 w[16];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = inner[i];
  w[8] = 0x80000000u;
#pragma unroll
  for (int i = 9; i < 15; i++) w[i] = 0;
  w[15] = (64 + 32) * 8;
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = p->out[i];
  sha256_compress_w(out, w);
}

// HMAC(key, V) for a 32-byte V
EB_HD void hmac_v(const hmac_pads* p, const u32* V, u32* out) {
  u32 w[16], st[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { w[i] = V[i]; st[i] = p->in[i]; }
  w[8] = 0x80000000u;
#pragma unroll
  for (int i = 9; i < 15; i++) w[i] = 0;
  w[15] = (64 + 32) * 8;
  sha256_compress_w(st, w);
  hmac_outer(p, st, out);
}

// HMAC(key, V || sep || a || b) for 32-byte V, a, b and a one-byte separator (HmacDRBG._update with a seed)
EB_HD void hmac_v_sep_seed(const hmac_pads* p, const u32* V, u32 sep, const u32* a, const u32* b, u32* out) {
  u32 w[16], st[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { w[i] = V[i]; st[i] = p->in[i]; }
  // the 65 bytes sep || a || b start at byte 32 of this block, shifted by one byte against the word grid
  w[8] = (sep << 24) | (a[0] >> 8);
#pragma unroll
  for (int i = 1; i < 8; i++) w[8 + i] = (a[i - 1] << 24) | (a[i] >> 8);
  sha256_compress_w(st, w);
  w[0] = (a[7] << 24) | (b[0] >> 8);
#pragma unroll
  for (int i = 1; i < 8; i++) w[i] = (b[i - 1] << 24) | (b[i] >> 8);
  w[8] = (b[7] << 24) | 0x00800000u;
#pragma unroll
  for (int i = 9; i < 15; i++) w[i] = 0;
  w[15] = (64 + 32 + 1 + 64) * 8;
  sha256_compress_w(st, w);
  hmac_outer(p, st, out);
}

// First output of HmacDRBG({entropy: priv, nonce: msg}).generate(32): big-endian words of k.
EB_HD void drbg_first_k(const u32* priv, const u32* msg, u32* k) {
  u32 K[8], V[8], T[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { K[i] = 0; V[i] = 0x01010101u; }
  hmac_pads p;
  hmac_key(&p, K);
  hmac_v_sep_seed(&p, V, 0x00, priv, msg, T);       // K = HMAC(K, V || 00 || seed)
  hmac_key(&p, T);
  hmac_v(&p, V, V);                                 // V = HMAC(K, V)
  hmac_v_sep_seed(&p, V, 0x01, priv, msg, T);       // K = HMAC(K, V || 01 || seed)
  hmac_key(&p, T);
  hmac_v(&p, V, V);                                 // V = HMAC(K, V)
  hmac_v(&p, V, k);                                 // generate: V = HMAC(K, V), output V
}

// ---- nonce kernel body
// kgiven != NULL: the caller's own nonce for this attempt (options.k, ec/index.js:154-157) instead of the DRBG
EB_HD void k256_sign_nonce_item(size_t i, size_t N, const uint8_t* e, const uint8_t* priv, const u32* gtab, u32* ws,
                                uint8_t* status, const uint8_t* kgiven = nullptr) {
  u32 ew[8], dw[8], kw[8], k[8];
  if (kgiven) {
#pragma unroll
    for (int w = 0; w < 8; w++) {
      const uint8_t* pk = kgiven + 32 * i + 4 * w;
      kw[w] = ((u32)pk[0] << 24) | ((u32)pk[1] << 16) | ((u32)pk[2] << 8) | pk[3];
    }
  } else {
#pragma unroll
    for (int w = 0; w < 8; w++) {
      const uint8_t* pe = e + 32 * i + 4 * w;
      const uint8_t* pd = priv + 32 * i + 4 * w;
      ew[w] = ((u32)pe[0] << 24) | ((u32)pe[1] << 16) | ((u32)pe[2] << 8) | pe[3];
      dw[w] = ((u32)pd[0] << 24) | ((u32)pd[1] << 16) | ((u32)pd[2] << 8) | pd[3];
    }
    drbg_first_k(dw, ew, kw);
  }
#pragma unroll
  for (int w = 0; w < 8; w++) k[w] = kw[7 - w];           // little-endian limbs
  u32 nn[8], ns1[8], one8[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  K256N::n(nn);
  sub_n<8>(ns1, nn, one8);
  bool le1 = (k[0] <= 1) && ((k[1] | k[2] | k[3] | k[4] | k[5] | k[6] | k[7]) == 0);
  bool slow = le1 || geq_n<8>(k, ns1);                     // ec/index.js:158-159 -> retry loop, slow path
  ge_jac acc = jac_infinity();
  if (slow) {
    acc.z = fe_one();
    copy_n<8>(k, one8);
  } else {
    u32 kk[8];
    copy_n<8>(kk, k);
    bool negg = (kk[0] & 1) == 0;
    if (negg) sub_n<8>(kk, nn, kk);
    u32 m[8];
    for (int w = 0; w < 8; w++) m[w] = (kk[w] >> 1) | ((w < 7 ? kk[w + 1] : 0u) << 31);
    // all table addresses are known up front: compute them, prefetch the (HBM-resident) entries, then add
    const u32* ent[GTAB_WINDOWS];
    u32 negmask = 0;
    for (int j = 0; j < GTAB_WINDOWS; j++) {
      const int pos = GTAB_W * j;
      u32 lo = 0, hi = 0;
      for (int w = 0; w < 8; w++) { lo = (w == (pos >> 5)) ? m[w] : lo; hi = (w == (pos >> 5) + 1) ? m[w] : hi; }
      u64 both = ((u64)hi << 32) | lo;
      u32 chunk = (u32)(both >> (pos & 31)) & ((1u << GTAB_W) - 1);
      const u32 half = 1u << (GTAB_W - 1);
      bool dneg = (j != GTAB_WINDOWS - 1) && (chunk < half);
      u32 idx = (j == GTAB_WINDOWS - 1) ? (chunk & (half - 1)) : (dneg ? half - 1 - chunk : chunk - half);
      ent[j] = gtab + ((size_t)j * GTAB_ENTRIES + idx) * 16;
      eb_prefetch_l2(ent[j]);
      if (dneg != negg) negmask |= 1u << j;
    }
    for (int j = 0; j < GTAB_WINDOWS; j++) {
      ge_aff P;
      P.x = load_fe(ent[j]);
      P.y = load_fe(ent[j] + 8);
      acc = jac_madd(acc, aff_neg_if(P, (negmask >> j) & 1));
    }
  }
  for (int w = 0; w < 8; w++) {
    ws[(size_t)w * N + i] = acc.x.v[w];
    ws[(size_t)(8 + w) * N + i] = acc.y.v[w];
    ws[(size_t)(16 + w) * N + i] = acc.z.v[w];
    ws[(size_t)(24 + w) * N + i] = k[w];
  }
  status[i] = slow ? ST_NEEDS_HOST : ST_TRUE;
}

// ---- finish kernel body: thread `tid` of `T` handles items tid, tid+T, ...
EB_HD void k256_sign_finish_thread(size_t tid, size_t T, size_t N, const uint8_t* e, const uint8_t* priv, u32 canonical,
                                   const u32* ws, u32* scratch, uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid,
                                   uint8_t* status) {
  u32 R2[8], one[8], nn[8];
  K256N::r2(R2); K256N::r1(one); K256N::n(nn);
  fe zprod = fe_one();
  u32 kprod[8];
  copy_n<8>(kprod, one);
  int cnt = 0;
  for (int j = 0; j < PREP_BATCH; j++) {
    size_t i = tid + (size_t)j * T;
    if (i >= N) break;
    cnt = j + 1;
    fe z; u32 k[8], km[8], t[8];
    for (int w = 0; w < 8; w++) { z.v[w] = ws[(size_t)(16 + w) * N + i]; k[w] = ws[(size_t)(24 + w) * N + i]; }
    for (int w = 0; w < 8; w++) { scratch[(size_t)w * N + i] = zprod.v[w]; scratch[(size_t)(8 + w) * N + i] = kprod[w]; }
    zprod = fe_mul(zprod, z);
    sc_mont_mul(km, k, R2);
    sc_mont_mul(t, kprod, km);
    copy_n<8>(kprod, t);
  }
  if (cnt == 0) return;
  fe zinv_all = fe_inv_chain(zprod);
  u32 kinv_all[8];
  sc_mont_inv(kinv_all, kprod);
  for (int j = cnt - 1; j >= 0; j--) {
    size_t i = tid + (size_t)j * T;
    fe x, y, z, zpre; u32 k[8], km[8], kpre[8], kinv[8], t[8];
    for (int w = 0; w < 8; w++) {
      x.v[w] = ws[(size_t)w * N + i]; y.v[w] = ws[(size_t)(8 + w) * N + i]; z.v[w] = ws[(size_t)(16 + w) * N + i];
      k[w] = ws[(size_t)(24 + w) * N + i];
      zpre.v[w] = scratch[(size_t)w * N + i]; kpre[w] = scratch[(size_t)(8 + w) * N + i];
    }
    fe zi = fe_mul(zinv_all, zpre);
    zinv_all = fe_mul(zinv_all, z);
    sc_mont_mul(km, k, R2);
    sc_mont_mul(kinv, kinv_all, kpre);                 // k^-1, Montgomery form
    sc_mont_mul(t, kinv_all, km);
    copy_n<8>(kinv_all, t);
    if (status[i] != ST_TRUE) continue;                // flagged by the nonce kernel
    fe zi2 = fe_sqr(zi);
    fe ax = fe_normalize(fe_mul(x, zi2));
    fe ay = fe_normalize(fe_mul(fe_mul(y, zi2), zi));
    u32 r[8];
    copy_n<8>(r, ax.v);
    bool xr_differ = geq_n<8>(r, nn);
    if (xr_differ) sub_n<8>(r, r, nn);                 // kpX.umod(n)
    if (is_zero_n<8>(r)) { status[i] = ST_NEEDS_HOST; continue; }
    u32 ev[8], dv[8], dm[8], rd[8], s[8];
    load_be<8>(ev, e + 32 * i);
    load_be<8>(dv, priv + 32 * i);
    sc_mont_mul(dm, dv, R2);
    sc_mont_mul(rd, r, dm);                            // r d mod n
    u32 cy = add_n<8>(t, rd, ev);
    if (cy || geq_n<8>(t, nn)) sub_n<8>(t, t, nn);
    sc_mont_mul(s, t, kinv);                           // k^-1 (r d + e) mod n
    if (is_zero_n<8>(s)) { status[i] = ST_NEEDS_HOST; continue; }
    u32 rec = (ay.v[0] & 1) | (xr_differ ? 2u : 0u);
    if (canonical) {
      u32 nh[8], d2[8];
      for (int w = 0; w < 8; w++) nh[w] = (nn[w] >> 1) | ((w < 7 ? nn[w + 1] : 0u) << 31);
      if (sub_n<8>(d2, nh, s) != 0) { sub_n<8>(s, nn, s); rec ^= 1; }
    }
    store_be<8>(out_r + 32 * i, r);
    store_be<8>(out_s + 32 * i, s);
    out_recid[i] = (uint8_t)rec;
  }
}

}  // namespace eb
