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

  static EB_HD void shr1(u32* a) {
    for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[N - 1] >>= 1;
  }
  static EB_HD int bitlen(const u32* a) {
    for (int i = N - 1; i >= 0; i--)
      if (a[i]) { int b = 32; while (!((a[i] >> (b - 1)) & 1)) b--; return 32 * i + b; }
    return 0;
  }
  // utils.getNAF, utils.js:15-44
  static EB_HD int get_naf(int16_t* naf, const u32* k, int w, int bits) {
    int len = bitlen(k); if (bits > len) len = bits; len += 1;
    int ws = 1 << (w + 1);
    u32 t[N]; copy_n<N>(t, k);
    for (int i = 0; i < len; i++) {
      int z = 0;
      int mod = (int)(t[0] & (u32)(ws - 1));
      if (t[0] & 1) {
        z = (mod > (ws >> 1) - 1) ? (ws >> 1) - mod : mod;
        u32 zz[N];
        for (int q = 0; q < N; q++) zz[q] = 0;
        zz[0] = (u32)(z < 0 ? -z : z);
        if (z >= 0) sub_n<N>(t, t, zz); else add_n<N>(t, t, zz);
      }
      naf[i] = (int16_t)z;
      shr1(t);
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
    store_fe_n<N>(out, r.x);
    store_fe_n<N>(out + N, r.y);
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

  static EB_HD uint8_t verify_item(size_t i, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* pub,
                                   const u32* tab) {
    typedef typename S::fe sc;
    const size_t LEN = 4 * N;
    u32 nmod[N]; S::Params::mod(nmod);
    sc ev, rv, sv;
    load_be<N>(ev.v, e + LEN * i); load_be<N>(rv.v, r + LEN * i); load_be<N>(sv.v, s + LEN * i);
    if (is_zero_n<N>(rv.v) || geq_n<N>(rv.v, nmod) || is_zero_n<N>(sv.v) || geq_n<N>(sv.v, nmod)) return 0;
    sc sinv = S::inv(S::to_mont(sv));
    sc u1 = S::mul(ev, sinv), u2 = S::mul(rv, sinv);
    aff Q;
    {
      fe t;
      load_be<N>(t.v, pub + 2 * LEN * i);       Q.x = F::to_mont(t);
      load_be<N>(t.v, pub + 2 * LEN * i + LEN); Q.y = F::to_mont(t);
    }
    jac acc = jmul_add(u1.v, u2.v, Q, tab);
    if (F::is_zero(acc.z)) return 0;
    fe z2 = F::sqr(acc.z);
    fe rp;
    load_be<N>(rp.v, r + LEN * i);
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
