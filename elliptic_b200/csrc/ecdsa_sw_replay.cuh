// ecdsa_sw_replay.cuh -- exact replay of the reference's double-scalar schedule on the non-GLV short
// curves (p256, p384) for un-validated off-curve public keys (lib/elliptic/ec/key.js:95).
//
// Reference: EC.verify (ec/index.js:188-229) -> Point.jmulAdd (short.js:443-450) ->
// BaseCurve._wnafMulAdd(1, [G, Q], [u1, u2], 2, true) (base.js:128-253): G carries the wnd-8 table built
// by g.precompute() in the EC constructor (ec/index.js:36, base.js:312-327, 357-370: (2i+1)G), Q gets
// wnd 1; both scalars go through utils.getNAF (utils.js:15-44); the loop doubles by runs of zero columns
// and adds with JPoint.mixedAdd (short.js:569-603); the result is compared by eqXToP (short.js:908-925).
// For a point off the curve the outcome depends on this exact sequence, so it is reproduced step by step.
#pragma once
#include "ecdsa_sw_body.cuh"

namespace eb {

template <class C>
struct SWReplay {
  typedef SW<C> W;
  typedef typename W::F F;
  typedef typename W::S S;
  typedef typename W::fe fe;
  typedef typename W::jac jac;
  typedef typename W::aff aff;
  static constexpr int N = C::N;
  static constexpr int NAF_PTS = 128;                 // digits |z| <= 255 only reach the first 128 entries
  static constexpr int TAB_WORDS = NAF_PTS * 2 * N;
  static constexpr int MAXLEN = 32 * N + 2;

  // utils.getNAF, utils.js:15-44.  One spare limb: k - z can exceed 2^(32N) for a raw scalar near the top.
  static EB_HD int get_naf(int16_t* naf, const u32* k, int w, int bits) {
    constexpr int M = N + 1;
    u32 t[M];
    for (int q = 0; q < N; q++) t[q] = k[q];
    t[N] = 0;
    int len = 0;
    for (int q = N - 1; q >= 0; q--)
      if (k[q]) { int b = 32; while (!((k[q] >> (b - 1)) & 1)) b--; len = 32 * q + b; break; }
    if (bits > len) len = bits;
    len += 1;
    int ws = 1 << (w + 1);
    for (int i = 0; i < len; i++) {
      int z = 0;
      int mod = (int)(t[0] & (u32)(ws - 1));
      if (t[0] & 1) {
        z = (mod > (ws >> 1) - 1) ? (ws >> 1) - mod : mod;
        u32 zz[M];
        for (int q = 0; q < M; q++) zz[q] = 0;
        zz[0] = (u32)(z < 0 ? -z : z);
        if (z >= 0) sub_n<M>(t, t, zz); else add_n<M>(t, t, zz);
      }
      naf[i] = (int16_t)z;
      for (int q = 0; q < M - 1; q++) t[q] = (t[q] >> 1) | (t[q + 1] << 31);
      t[M - 1] >>= 1;
    }
    return len;
  }

  // table entry i: (2i+1) G, affine, Montgomery form
  static EB_HD void tab_entry(int idx, u32* out) {
    aff g; g.x = C::gx(); g.y = C::gy();
    u32 s = 2 * idx + 1;
    jac acc = W::infinity();
    for (int k = 8; k >= 0; k--) {
      acc = W::dbl(acc);
      if ((s >> k) & 1) acc = W::madd(acc, g);
    }
    aff r = W::to_aff(acc);
    store_fe_n<N>(out, F::canon(r.x));
    store_fe_n<N>(out + N, F::canon(r.y));
  }

  // Point.jmulAdd(u1, Q, u2) on G, as scheduled by _wnafMulAdd (base.js:128-253)
  static EB_HD jac jmul_add(const u32* u1, const u32* u2, const aff& Q, const u32* tab) {
    int16_t n0[MAXLEN], n1[MAXLEN];
    int l0 = get_naf(n0, u1, 8, C::BITS);
    int l1 = get_naf(n1, u2, 1, C::BITS);
    int max = l0 > l1 ? l0 : l1;
    jac acc = W::infinity();
    for (int k = max; k >= 0; k--) {
      int run = 0, z0 = 0, z1 = 0;
      while (k >= 0) {
        z0 = k < l0 ? n0[k] : 0; z1 = k < l1 ? n1[k] : 0;
        if (z0 || z1) break;
        run++; k--;
      }
      if (k >= 0) run++;
      if (!F::is_zero(acc.z))
        for (int d = 0; d < run; d++) acc = W::dbl(acc);           // dblp: no-op on infinity (short.js:608-609)
      if (k < 0) break;
      if (z0) {
        int az = z0 < 0 ? -z0 : z0;
        aff p;
        p.x = load_fe_n<N>(tab + 2 * N * ((az - 1) >> 1));
        p.y = load_fe_n<N>(tab + 2 * N * ((az - 1) >> 1) + N);
        if (z0 < 0) p.y = F::neg(p.y);
        acc = W::madd(acc, p);
      }
      if (z1) {
        aff p = Q;
        if (z1 < 0) p.y = F::neg(p.y);
        acc = W::madd(acc, p);
      }
    }
    return acc;
  }

  // ---- affine Point.add / Point.dbl with the reference's early-outs (short.js:365-412) ----------------
  struct raff { fe x, y; bool inf; };
  static EB_HD raff r_inf() { raff r; r.x = F::zero(); r.y = F::zero(); r.inf = true; return r; }
  static EB_HD raff r_neg(const raff& p) { raff r = p; if (!p.inf) r.y = F::neg(p.y); return r; }
  static EB_HD bool r_eq(const raff& a, const raff& b) {
    return a.inf == b.inf && (a.inf || (F::eq(a.x, b.x) && F::eq(a.y, b.y)));
  }
  static EB_HD raff r_dbl(const raff& p) {
    if (p.inf) return p;
    fe ys1 = F::add(p.y, p.y);
    if (F::is_zero(ys1)) return r_inf();
    fe x2 = F::sqr(p.x);
    fe dyinv = F::inv(ys1);
    fe c = F::mul(F::sub(F::add(F::add(x2, x2), x2), C::three()), dyinv);      // (3x^2 + a) / 2y, a = -3
    raff r; r.inf = false;
    r.x = F::sub(F::sqr(c), F::add(p.x, p.x));
    r.y = F::sub(F::mul(c, F::sub(p.x, r.x)), p.y);
    return r;
  }
  static EB_HD raff r_add(const raff& a, const raff& b) {
    if (a.inf) return b;
    if (b.inf) return a;
    if (r_eq(a, b)) return r_dbl(a);
    if (r_eq(r_neg(a), b)) return r_inf();
    if (F::eq(a.x, b.x)) return r_inf();
    fe c = F::sub(a.y, b.y);
    if (!F::is_zero(c)) c = F::mul(c, F::inv(F::sub(a.x, b.x)));
    raff r; r.inf = false;
    r.x = F::sub(F::sub(F::sqr(c), a.x), b.x);
    r.y = F::sub(F::mul(c, F::sub(a.x, r.x)), a.y);
    return r;
  }

  // BaseCurve._wnafMul(P, k) (base.js:86-126): w = 4 table P, 3P, .. by affine adds (_getNAFPoints,
  // base.js:357-370; only the first 8 of its 15 entries can be indexed by a w = 4 NAF), Jacobian
  // accumulator, runs of doublings, mixedAdd of +-table entries.
  static EB_HD jac wnaf_mul(const u32* k, const aff& P) {
    raff tab[8];
    tab[0].x = P.x; tab[0].y = P.y; tab[0].inf = false;
    raff d = r_dbl(tab[0]);
    for (int t = 1; t < 8; t++) tab[t] = r_add(tab[t - 1], d);
    int16_t naf[MAXLEN];
    int len = get_naf(naf, k, 4, C::BITS);
    jac acc = W::infinity();
    for (int i = len - 1; i >= 0; i--) {
      int l = 0;
      for (; i >= 0 && naf[i] == 0; i--) l++;
      if (i >= 0) l++;
      if (!F::is_zero(acc.z))
        for (int t = 0; t < l; t++) acc = W::dbl(acc);
      if (i < 0) break;
      int z = naf[i];
      raff p = tab[((z < 0 ? -z : z) - 1) >> 1];
      if (z < 0) p = r_neg(p);
      if (p.inf) continue;                                   // mixedAdd: p.isInfinity() -> this
      aff q; q.x = p.x; q.y = p.y;
      acc = W::madd(acc, q);
    }
    return acc;
  }

  // Point.mul (k1 == nullptr) / G.mulAdd(k1, P, k2) for an off-curve P: the reference's schedule, then toP.
  static EB_HD uint8_t mul_add_item(size_t i, const uint8_t* k1, const uint8_t* k2, const uint8_t* pts, const u32* tab,
                                    uint8_t* out) {
    const size_t LEN = C::LEN;
    u32 u1[N], u2[N];
    if (k1) W::ldb(u1, k1 + LEN * i);
    W::ldb(u2, k2 + LEN * i);
    aff P = W::load_point(pts, i);
    jac acc = k1 ? jmul_add(u1, u2, P, tab) : wnaf_mul(u2, P);
    for (size_t b = 0; b < 2 * LEN; b++) out[2 * LEN * i + b] = 0;
    if (F::is_zero(acc.z)) return 7;
    W::store_point(out, i, W::to_aff(acc));
    return 1;
  }

  static EB_HD uint8_t verify_item(size_t i, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* pub,
                                   const u32* tab) {
    typedef typename S::fe sc;
    const size_t LEN = C::LEN;
    u32 nmod[N]; S::Params::mod(nmod);
    sc ev, rv, sv;
    W::ldb(ev.v, e + LEN * i); W::ldb(rv.v, r + LEN * i); W::ldb(sv.v, s + LEN * i);
    if (is_zero_n<N>(rv.v) || geq_n<N>(rv.v, nmod) || is_zero_n<N>(sv.v) || geq_n<N>(sv.v, nmod)) return 0;
    sc sinv = S::inv(S::to_mont(sv));
    sc u1 = S::mul(ev, sinv), u2 = S::mul(rv, sinv);
    aff Q;
    {
      fe t;
      W::ldb(t.v, pub + 2 * LEN * i);       Q.x = F::to_mont(t);
      W::ldb(t.v, pub + 2 * LEN * i + LEN); Q.y = F::to_mont(t);
    }
    jac acc = jmul_add(u1.v, u2.v, Q, tab);
    if (F::is_zero(acc.z)) return 0;
    fe z2 = F::sqr(acc.z);
    fe rp;
    W::ldb(rp.v, r + LEN * i);
    if (F::eq(acc.x, F::mul(F::to_mont(rp), z2))) return 1;
    u32 pmn[N]; C::p_minus_n(pmn);
    if (!geq_n<N>(rp.v, pmn)) {
      fe rn;
      add_n<N>(rn.v, rp.v, nmod);
      if (F::eq(acc.x, F::mul(F::to_mont(rn), z2))) return 1;
    }
    return 0;
  }
};

}  // namespace eb
