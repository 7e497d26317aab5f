// fp_mont.cuh -- generic prime-field arithmetic in Montgomery form (CIOS), N x 32-bit limbs.
//
// Replaces bn.js `Mont` (reference dist/elliptic.js:7312-7381: mul/imul/convertTo/
// convertFrom/invm) for the curves the reference runs on BN.mont(p) -- p256, p384
// (lib/elliptic/curves.js:75,90; `prime: null` -> curve/base.js:14) -- and the
// scalar fields mod n of every short curve.  bn.js uses R = 2^(26*ceil(bits/26));
// here R = 2^(32 N).  Either way only canonical residues are observable.
//
// P (parameter struct) provides: N, mod(u32*), r1(u32*) = R mod p, r2(u32*) = R^2 mod p,
// n0inv = -p^-1 mod 2^32.
#pragma once
#include "limbs.cuh"

namespace eb {

template <int N>
struct fe_n { u32 v[N]; };

template <class P>
struct Fp {
  static constexpr int N = P::N;
  typedef P Params;
  typedef fe_n<N> fe;

  static EB_HD fe zero() { fe r; for (int i = 0; i < N; i++) r.v[i] = 0; return r; }
  static EB_HD fe one() { fe r; P::r1(r.v); return r; }   // Montgomery form of 1

  // a*b*R^-1 mod p.  Requires a < R, b < p (or a < p, b < R); result in [0, p).
  static EB_HD fe mul(const fe& a, const fe& b) {
    u32 p[N]; P::mod(p);
    u32 t[N + 2];
#pragma unroll
    for (int i = 0; i < N + 2; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      u64 c = 0;
#pragma unroll
      for (int j = 0; j < N; j++) {
        c += (u64)a.v[j] * b.v[i] + t[j];
        t[j] = (u32)c; c >>= 32;
      }
      c += t[N];
      t[N] = (u32)c; t[N + 1] = (u32)(c >> 32);
      u32 m = t[0] * P::n0inv;
      c = (u64)m * p[0] + t[0];
      c >>= 32;
#pragma unroll
      for (int j = 1; j < N; j++) {
        c += (u64)m * p[j] + t[j];
        t[j - 1] = (u32)c; c >>= 32;
      }
      c += t[N];
      t[N - 1] = (u32)c;
      t[N] = t[N + 1] + (u32)(c >> 32);
      t[N + 1] = 0;
    }
    fe r, d;
    u32 bw = sub_n<N>(d.v, t, p);
    bool ge = t[N] != 0 || bw == 0;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = ge ? d.v[i] : t[i];
    return r;
  }
  static EB_HD fe sqr(const fe& a) { return mul(a, a); }

  static EB_HD fe add(const fe& a, const fe& b) {
    u32 p[N]; P::mod(p);
    fe r, d;
    u32 cy = add_n<N>(r.v, a.v, b.v);
    u32 bw = sub_n<N>(d.v, r.v, p);
    bool ge = cy != 0 || bw == 0;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = ge ? d.v[i] : r.v[i];
    return r;
  }
  static EB_HD fe sub(const fe& a, const fe& b) {
    u32 p[N]; P::mod(p);
    fe r, d;
    u32 bw = sub_n<N>(r.v, a.v, b.v);
    add_n<N>(d.v, r.v, p);
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = bw ? d.v[i] : r.v[i];
    return r;
  }
  static EB_HD fe neg(const fe& a) { return sub(zero(), a); }
  static EB_HD fe dbl(const fe& a) { return add(a, a); }
  static EB_HD bool is_zero(const fe& a) { return is_zero_n<N>(a.v); }
  static EB_HD bool eq(const fe& a, const fe& b) { return eq_n<N>(a.v, b.v); }
  static EB_HD fe cmov(const fe& a, const fe& b, bool c) {
    fe r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = c ? b.v[i] : a.v[i];
    return r;
  }

  // plain integer (< 2^(32N)) -> Montgomery form (reduces mod p: `toRed`, dist:7292-7296)
  static EB_HD fe to_mont(const fe& a) { fe r2; P::r2(r2.v); return mul(a, r2); }
  // Montgomery form -> canonical residue (`fromRed`)
  static EB_HD fe from_mont(const fe& a) { fe o = zero(); o.v[0] = 1; return mul(a, o); }

  // plain a >= p ?
  static EB_HD bool geq_mod(const u32* a) { u32 p[N]; P::mod(p); return geq_n<N>(a, p); }

  // a^e, e = N limbs (square-and-multiply, MSB first); a in Montgomery form
  static EB_HD fe pow(const fe& a, const u32* e) {
    fe r = one();
    bool started = false;
    for (int i = 32 * N - 1; i >= 0; i--) {
      if (started) r = sqr(r);
      if ((e[i >> 5] >> (i & 31)) & 1) {
        r = started ? mul(r, a) : a;
        started = true;
      }
    }
    return r;
  }
  // a^(p-2): inverse (0 -> 0, like bn.js _invmp(0), dist:6568-6579)
  static EB_HD fe inv(const fe& a) {
    u32 e[N]; P::mod(e);
    e[0] -= 2;  // p is odd and > 2: no borrow
    return pow(a, e);
  }
};

template <int N>
EB_HD fe_n<N> load_fe_n(const u32* src) { fe_n<N> a; for (int i = 0; i < N; i++) a.v[i] = src[i]; return a; }
template <int N>
EB_HD void store_fe_n(u32* dst, const fe_n<N>& a) { for (int i = 0; i < N; i++) dst[i] = a.v[i]; }

}  // namespace eb
