// ecdsa_k256_replay.cuh -- exact replay of the reference's secp256k1 double-scalar schedule.
//
// For a public key that is NOT on the curve the reference still "verifies"
// (lib/elliptic/ec/key.js:95 and curve/short.js:251-271 never validate {x,y} / uncompressed keys).
// Its answer is then whatever its own sequence of chord-and-tangent operations produces, so the only
// way to return the identical boolean is to run the identical sequence:
//   ShortCurve._endoWnafMulAdd   short.js:218-249  (exact _endoSplit / divRound, sign fix-ups)
//   BaseCurve._wnafMulAdd        base.js:128-253   (getNAF w=7 for +-G, +-beta*G; JSF comb for Q, beta*Q)
//   utils.getNAF / getJSF        utils.js:15-101
//   Point.add / dbl, JPoint.add / mixedAdd / dbl, eqXToP   short.js:365-412, 532-603, 668-737, 908-925
// This is the slow, divergent path: one thread per flagged item, run only for the (rare) items the
// fast kernel marked ST_NEEDS_HOST.  Same file compiles for the host-emulation tests.
#pragma once
#include "ecdsa_k256_body.cuh"

namespace eb {

constexpr int REPLAY_NAF_W = 7, REPLAY_NAF_PTS = 128;   // precomputed/secp256k1.js: naf wnd 7
constexpr int REPLAY_TAB_WORDS = REPLAY_NAF_PTS * 16 * 2;   // (2i+1)G then beta images, x||y limbs

struct rp_aff { fe x, y; bool inf; };          // canonical coordinates

EB_HD rp_aff rp_inf() { rp_aff r; r.x = fe_zero(); r.y = fe_zero(); r.inf = true; return r; }
EB_HD rp_aff rp_neg(const rp_aff& p) { rp_aff r = p; if (!p.inf) r.y = fe_normalize(fe_neg(p.y)); return r; }
EB_HD bool rp_eq(const rp_aff& a, const rp_aff& b) {
  if (a.inf != b.inf) return false;
  return a.inf || (eq_n<8>(a.x.v, b.x.v) && eq_n<8>(a.y.v, b.y.v));
}
// Point.dbl, short.js:394-412
EB_HD rp_aff rp_dbl(const rp_aff& p) {
  if (p.inf) return p;
  fe ys1 = fe_add(p.y, p.y);
  if (fe_is_zero(ys1)) return rp_inf();
  fe x2 = fe_sqr(p.x);
  fe c = fe_mul(fe_mul_small(x2, 3), fe_inv(ys1));
  rp_aff r;
  r.x = fe_sub(fe_sqr(c), fe_add(p.x, p.x));
  r.y = fe_normalize(fe_sub(fe_mul(c, fe_sub(p.x, r.x)), p.y));
  r.x = fe_normalize(r.x);
  r.inf = false;
  return r;
}
// Point.add, short.js:365-392
EB_HD rp_aff rp_add(const rp_aff& a, const rp_aff& b) {
  if (a.inf) return b;
  if (b.inf) return a;
  if (rp_eq(a, b)) return rp_dbl(a);
  if (rp_eq(rp_neg(a), b) || eq_n<8>(a.x.v, b.x.v)) return rp_inf();
  fe c = fe_sub(a.y, b.y);
  if (!fe_is_zero(c)) c = fe_mul(c, fe_inv(fe_sub(a.x, b.x)));
  rp_aff r;
  r.x = fe_sub(fe_sub(fe_sqr(c), a.x), b.x);
  r.y = fe_normalize(fe_sub(fe_mul(c, fe_sub(a.x, r.x)), a.y));
  r.x = fe_normalize(r.x);
  r.inf = false;
  return r;
}

// JPoint ops with the reference's early-outs (infinity operands) -- group-law bodies from ge_k256.cuh
EB_HD ge_jac rp_jdbl(const ge_jac& p) { return fe_is_zero(p.z) ? p : jac_dbl_inl(p); }
EB_HD ge_jac rp_jmadd(const ge_jac& a, const rp_aff& p) {
  if (fe_is_zero(a.z)) {                                   // short.js:571-572: p.toJ()
    if (p.inf) return jac_infinity();
    ge_aff q; q.x = p.x; q.y = p.y;
    return jac_from_aff(q);
  }
  if (p.inf) return a;                                     // short.js:575-576
  ge_aff q; q.x = p.x; q.y = p.y;
  return jac_madd_inl(a, q);
}

// 9-limb helpers for the exact divRound
template <int N>
EB_HD bool geq_big(const u32* a, const u32* b) { u32 t[N]; return sub_n<N>(t, a, b) == 0; }

// exact round(a*k / n) for a < 2^128 (4 limbs), k < n: BN.divRound, dist/elliptic.js:6387-6404
EB_HD void rp_mul_div_round_n(u32* c4, const u32* a4, const u32* g8, const u32* k8) {
  u32 nn[8]; K256N::n(nn);
  // estimate c' = (k*g + 2^383) >> 384  (|c' - c| <= 1)
  u32 t[16];
  mul_rect<8, 8>(t, k8, g8);
  const u32 half[4] = {0, 0, 0, 0x80000000u};
  u32 cy = add_n<4>(t + 8, t + 8, half);
  u32 one[4] = {cy, 0, 0, 0};
  u32 c[5];
  add_n<4>(c, t + 12, one);
  c[4] = 0;
  // v = a*k + (n-1)/2 - c'*n   (signed, 13 limbs two's complement)
  u32 ak[13], cn[13], v[13], h[13];
  { u32 p[12]; mul_rect<8, 4>(p, k8, a4); for (int i = 0; i < 12; i++) ak[i] = p[i]; ak[12] = 0; }
  { u32 p[12]; mul_rect<8, 4>(p, nn, c); for (int i = 0; i < 12; i++) cn[i] = p[i]; cn[12] = 0; }
  for (int i = 0; i < 13; i++) h[i] = 0;
  for (int i = 0; i < 8; i++) h[i] = (nn[i] >> 1) | ((i < 7 ? nn[i + 1] : 0u) << 31);   // (n-1)/2 = n >> 1
  add_n<13>(v, ak, h);
  sub_n<13>(v, v, cn);
  u32 n13[13];
  for (int i = 0; i < 13; i++) n13[i] = i < 8 ? nn[i] : 0;
  u32 onec[4] = {1, 0, 0, 0};
  if (v[12] >> 31) {                                   // v < 0  -> c = c' - 1
    sub_n<4>(c, c, onec);
  } else if (geq_big<13>(v, n13)) {                    // v >= n -> c = c' + 1
    add_n<4>(c, c, onec);
  }
  for (int i = 0; i < 4; i++) c4[i] = c[i];
}

// ShortCurve._endoSplit, short.js:168-185 -- exact; k1, k2 as sign + magnitude (8 limbs)
EB_HD void rp_endo_split(const u32* k, u32* k1, bool* neg1, u32* k2, bool* neg2) {
  const u32 g1[8] = {0x45dbb031u, 0xe893209au, 0x71e8ca7fu, 0x3daa8a14u, 0x9284eb15u, 0xe86c90e4u, 0xa7d46bcdu, 0x3086d221u};
  const u32 g2[8] = {0x8ac47f71u, 0x1571b4aeu, 0x9df506c6u, 0x221208acu, 0x0abfe4c4u, 0x6f547fa9u, 0x010e8828u, 0xe4437ed6u};
  const u32 a1[4] = {0x9284eb15u, 0xe86c90e4u, 0xa7d46bcdu, 0x3086d221u};
  const u32 mb1[4] = {0x0abfe4c3u, 0x6f547fa9u, 0x010e8828u, 0xe4437ed6u};
  const u32 a2[5] = {0x9d44cfd8u, 0x57c1108du, 0xa8e2f3f6u, 0x14ca50f7u, 0x00000001u};
  const u32 b2[4] = {0x9284eb15u, 0xe86c90e4u, 0xa7d46bcdu, 0x3086d221u};
  u32 c1[4], c2[4];
  rp_mul_div_round_n(c1, b2, g1, k);       // c1 = round(b2 * k / n)
  rp_mul_div_round_n(c2, mb1, g2, k);      // c2 = round(-b1 * k / n)
  u32 p[9], q[8], t1[8], t2[8];
  mul_rect<4, 4>(p, c1, a1);
  sub_n<8>(t1, k, p);
  mul_rect<5, 4>(p, a2, c2);
  sub_n<8>(t1, t1, p);                     // k1 = k - c1 a1 - c2 a2  (two's complement, |k1| small)
  mul_rect<4, 4>(t2, c1, mb1);
  mul_rect<4, 4>(q, c2, b2);
  sub_n<8>(t2, t2, q);                     // k2 = -(c1 b1 + c2 b2)
  *neg1 = (t1[7] >> 31) != 0;
  *neg2 = (t2[7] >> 31) != 0;
  u32 a[8];
  neg256(a, t1); cmov_n<8>(t1, a, *neg1);
  neg256(a, t2); cmov_n<8>(t2, a, *neg2);
  copy_n<8>(k1, t1); copy_n<8>(k2, t2);
}

EB_HD int rp_bitlen(const u32* a) {
  for (int i = 7; i >= 0; i--)
    if (a[i]) { int b = 32; while (!((a[i] >> (b - 1)) & 1)) b--; return 32 * i + b; }
  return 0;
}
EB_HD void rp_shr1(u32* a) {
  for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
  a[7] >>= 1;
}
// utils.getNAF, utils.js:15-44
EB_HD int rp_get_naf(int8_t* naf, const u32* k, int w, int bits) {
  int len = rp_bitlen(k); if (bits > len) len = bits; len += 1;
  int ws = 1 << (w + 1);
  u32 t[8]; copy_n<8>(t, k);
  for (int i = 0; i < len; i++) {
    int z = 0;
    int mod = (int)(t[0] & (u32)(ws - 1));
    if (t[0] & 1) {
      z = (mod > (ws >> 1) - 1) ? (ws >> 1) - mod : mod;
      u32 zz[8] = {(u32)(z < 0 ? -z : z), 0, 0, 0, 0, 0, 0, 0};
      if (z >= 0) sub_n<8>(t, t, zz); else add_n<8>(t, t, zz);
    }
    naf[i] = (int8_t)z;
    rp_shr1(t);
  }
  return len;
}
// utils.getJSF, utils.js:47-101
EB_HD int rp_get_jsf(int8_t* j1, int8_t* j2, const u32* a, const u32* b) {
  u32 k1[8], k2[8];
  copy_n<8>(k1, a); copy_n<8>(k2, b);
  int d1 = 0, d2 = 0, len = 0;
  while ((d1 ? true : !is_zero_n<8>(k1)) || (d2 ? true : !is_zero_n<8>(k2))) {
    int m14 = (int)((k1[0] & 3) + d1) & 3;
    int m24 = (int)((k2[0] & 3) + d2) & 3;
    if (m14 == 3) m14 = -1;
    if (m24 == 3) m24 = -1;
    int u1, u2;
    if ((m14 & 1) == 0) u1 = 0;
    else { int m8 = (int)((k1[0] & 7) + d1) & 7; u1 = ((m8 == 3 || m8 == 5) && m24 == 2) ? -m14 : m14; }
    j1[len] = (int8_t)u1;
    if ((m24 & 1) == 0) u2 = 0;
    else { int m8 = (int)((k2[0] & 7) + d2) & 7; u2 = ((m8 == 3 || m8 == 5) && m14 == 2) ? -m24 : m24; }
    j2[len] = (int8_t)u2;
    len++;
    if (2 * d1 == u1 + 1) d1 = 1 - d1;
    if (2 * d2 == u2 + 1) d2 = 1 - d2;
    rp_shr1(k1); rp_shr1(k2);
  }
  return len;
}

struct rp_any { bool is_j; rp_aff a; ge_jac j; };

// g.jmulAdd(u1, Q, u2) exactly as the reference computes it; returns the Jacobian result.
// tab: REPLAY_TAB_WORDS words: (2i+1)G, i < 128, then their beta images (canonical x||y limbs).
// u1 == nullptr: Q.mul(u2) = _endoWnafMulAdd([Q], [u2]) (short.js:428-429), i.e. the JSF comb alone.
EB_HD ge_jac rp_jmul_add(const u32* u1, const u32* u2, const fe& qx, const fe& qy, const u32* tab) {
  u32 k1g[8], k2g[8], k1q[8], k2q[8];
  bool ng = false, nbg = false, nq, nbq;
  if (u1) rp_endo_split(u1, k1g, &ng, k2g, &nbg);
  rp_endo_split(u2, k1q, &nq, k2q, &nbq);
  rp_aff Qp; Qp.x = fe_normalize(qx); Qp.y = fe_normalize(qy); Qp.inf = false;
  rp_aff Qb = Qp;
  Qb.x = fe_normalize(fe_mul(Qp.x, fe_beta()));          // _getBeta, short.js:290
  if (nq) Qp = rp_neg(Qp);
  if (nbq) Qb = rp_neg(Qb);
  // comb, base.js:161-178
  rp_any comb[4];
  comb[0].is_j = false; comb[0].a = Qp;
  comb[3].is_j = false; comb[3].a = Qb;
  fe nyb = fe_normalize(fe_neg(Qb.y));
  ge_jac jq = rp_jmadd(jac_infinity(), Qp);              // Qp.toJ()
  if (eq_n<8>(Qp.y.v, Qb.y.v)) {
    comb[1].is_j = false; comb[1].a = rp_add(Qp, Qb);
    comb[2].is_j = true;  comb[2].j = rp_jmadd(jq, rp_neg(Qb));
  } else if (eq_n<8>(Qp.y.v, nyb.v)) {
    comb[1].is_j = true;  comb[1].j = rp_jmadd(jq, Qb);
    comb[2].is_j = false; comb[2].a = rp_add(Qp, rp_neg(Qb));
  } else {
    comb[1].is_j = true;  comb[1].j = rp_jmadd(jq, Qb);
    comb[2].is_j = true;  comb[2].j = rp_jmadd(jq, rp_neg(Qb));
  }
  const int8_t INDEX[9] = {-3, -1, -5, -7, 0, 7, 5, 1, 3};
  int8_t j1[264], j2[264], nqd[264], ngd[264], nbgd[264];
  int max = rp_get_jsf(j1, j2, k1q, k2q);
  for (int j = 0; j < max; j++) nqd[j] = INDEX[(j1[j] + 1) * 3 + (j2[j] + 1)];
  int lq = max;
  int lg = u1 ? rp_get_naf(ngd, k1g, REPLAY_NAF_W, 256) : 0;
  int lbg = u1 ? rp_get_naf(nbgd, k2g, REPLAY_NAF_W, 256) : 0;
  if (lg > max) max = lg;
  if (lbg > max) max = lbg;
  ge_jac acc = jac_infinity();
  for (int i = max; i >= 0; i--) {
    int k = 0, zg = 0, zbg = 0, zq = 0;
    while (i >= 0) {
      zg = i < lg ? ngd[i] : 0; zbg = i < lbg ? nbgd[i] : 0; zq = i < lq ? nqd[i] : 0;
      if (zg || zbg || zq) break;
      k++; i--;
    }
    if (i >= 0) k++;
    for (int d = 0; d < k; d++) acc = rp_jdbl(acc);     // dblp, short.js:605-619
    if (i < 0) break;
    if (zg) {
      int az = zg < 0 ? -zg : zg;
      rp_aff p; p.inf = false;
      p.x = load_fe(tab + 16 * ((az - 1) >> 1)); p.y = load_fe(tab + 16 * ((az - 1) >> 1) + 8);
      if ((zg < 0) != ng) p = rp_neg(p);
      acc = rp_jmadd(acc, p);
    }
    if (zbg) {
      int az = zbg < 0 ? -zbg : zbg;
      rp_aff p; p.inf = false;
      p.x = load_fe(tab + 16 * (REPLAY_NAF_PTS + ((az - 1) >> 1))); p.y = load_fe(tab + 16 * (REPLAY_NAF_PTS + ((az - 1) >> 1)) + 8);
      if ((zbg < 0) != nbg) p = rp_neg(p);
      acc = rp_jmadd(acc, p);
    }
    if (zq) {
      int az = zq < 0 ? -zq : zq;
      const rp_any& c = comb[(az - 1) >> 1];
      if (!c.is_j) acc = rp_jmadd(acc, zq < 0 ? rp_neg(c.a) : c.a);
      else {
        ge_jac q = c.j;
        if (zq < 0) q.y = fe_neg(q.y);
        acc = jac_add_inl(acc, q);
      }
    }
  }
  return acc;
}

// full verify for one flagged item (inputs as in verify_item; e reduced mod n here)
EB_HD uint8_t rp_verify_item(size_t i, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* pub,
                             const u32* tab) {
  u32 ev[8], rv[8], sv[8], R2[8], one_m[8];
  load_be<8>(ev, e + 32 * i); load_be<8>(rv, r + 32 * i); load_be<8>(sv, s + 32 * i);
  if (!sc_in_range(rv) || !sc_in_range(sv)) return 0;
  K256N::r2(R2); K256N::r1(one_m);
  u32 sm[8], sinv[8], u1[8], u2[8];
  sc_mont_mul(sm, sv, R2);
  sc_mont_inv(sinv, sm);
  sc_mont_mul(u1, ev, sinv);
  sc_mont_mul(u2, rv, sinv);
  fe qx = fe_from_be(pub + 64 * i), qy = fe_from_be(pub + 64 * i + 32);
  ge_jac acc = rp_jmul_add(u1, u2, qx, qy, tab);
  if (fe_is_zero(acc.z)) return 0;
  fe z2 = fe_sqr(acc.z);
  fe rf = fe_from_be(r + 32 * i);
  if (fe_eq(acc.x, fe_mul(rf, z2))) return 1;
  const u32 pmn[8] = {0x2fc9baeeu, 0x402da172u, 0x50b75fc4u, 0x45512319u, 0x00000001u, 0, 0, 0};
  if (!geq_n<8>(rf.v, pmn)) {
    u32 nn[8]; K256N::n(nn);
    fe rn;
    add_n<8>(rn.v, rf.v, nn);
    if (fe_eq(acc.x, fe_mul(rn, z2))) return 1;
  }
  return 0;
}

// Point.mulAdd / Point.mul for an off-curve point (the constructor never validates, short.js:251-271):
// same schedule, then JPoint.toP (short.js:516-526).  Scalars are used as given (no reduction mod n), as
// the reference does.  k1 == nullptr: P.mul(k2).
EB_HD uint8_t rp_mul_add_item(size_t i, const uint8_t* k1, const uint8_t* k2, const uint8_t* pts, const u32* tab,
                              uint8_t* out) {
  u32 u1[8], u2[8];
  if (k1) load_be<8>(u1, k1 + 32 * i);
  load_be<8>(u2, k2 + 32 * i);
  fe px = fe_from_be(pts + 64 * i), py = fe_from_be(pts + 64 * i + 32);
  ge_jac acc = rp_jmul_add(k1 ? u1 : nullptr, u2, px, py, tab);
  for (int b = 0; b < 64; b++) out[64 * i + b] = 0;
  if (fe_is_zero(acc.z)) return ST_INFINITY;
  ge_aff q = jac_to_aff(acc);
  fe qx = fe_normalize(q.x), qy = fe_normalize(q.y);
  store_be<8>(out + 64 * i, qx.v);
  store_be<8>(out + 64 * i + 32, qy.v);
  return ST_TRUE;
}

// table builder: entry t < 128: (2t+1)G ; t >= 128: (beta*x, y) of entry t-128
EB_HD void rp_tab_entry(int t, u32* out16) {
  int idx = t & (REPLAY_NAF_PTS - 1);
  ge_jac acc = jac_infinity();
  ge_aff g = k256_G();
  u32 s = 2 * idx + 1;
  for (int k = 8; k >= 0; k--) {
    acc = jac_dbl_inl(acc);
    if ((s >> k) & 1) acc = jac_madd_inl(acc, g);
  }
  ge_aff a = jac_to_aff(acc);
  fe x = fe_normalize(t >= REPLAY_NAF_PTS ? fe_mul(a.x, fe_beta()) : a.x);
  fe y = fe_normalize(a.y);
  store_fe(out16, x); store_fe(out16 + 8, y);
}

}  // namespace eb
