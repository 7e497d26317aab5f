// sw_runtime.cuh -- short Weierstrass curves given at RUN time: y^2 = x^3 + a x + b over any odd prime p of up
// to 32 N bits (N = 8, 12, 18 limbs), for the batch form of the reference's generic `.curve` API:
//   new elliptic.curve.short({p, a, b, ...})  (lib/elliptic/curve/short.js:10-24), curve.point(x, y),
//   Point.add / dbl / mul / mulAdd / validate (short.js:365-450, 206-216; JPoint :532-603, :739-800).
// The six presets keep their specialised kernels; this path exists so that ANY parameter set the reference accepts
// (its own test uses the toy curve p = 0x1d, a = 4, b = 0x14, test/curve-test.js:9-22) has a batch entry point.
// Field: CIOS Montgomery with the modulus in a by-value parameter block; group law: Jacobian add / double for a
// general `a` with every exceptional case resolved (tiny fields hit them all the time).  For on-curve inputs the
// affine results are the group-law answers, i.e. exactly what the reference returns; points that do not satisfy the
// curve equation are reported (status 4), never guessed.  Not tuned: this is the low-traffic end of the API.
#pragma once
#include "limbs.cuh"

namespace eb {

template <int N>
struct RtCurve {
  u32 p[N], r1[N], r2[N];      // modulus, R mod p, R^2 mod p  (R = 2^(32 N))
  u32 a[N], b[N];              // Montgomery form
  u32 n0inv;
  u32 len;                     // bytes of a coordinate on the wire
  u32 a_is_zero;
};

template <int N> struct rfe { u32 v[N]; };

template <int N>
struct RtF {
  typedef rfe<N> fe;
  static EB_HD fe zero() { fe r; for (int i = 0; i < N; i++) r.v[i] = 0; return r; }
  static EB_HD fe mul(const fe& a, const fe& b, const RtCurve<N>& C) {
    u32 t[N + 2];
    for (int i = 0; i < N + 2; i++) t[i] = 0;
    for (int i = 0; i < N; i++) {
      u64 c = 0;
      for (int j = 0; j < N; j++) {
        c += (u64)a.v[j] * b.v[i] + t[j];
        t[j] = (u32)c; c >>= 32;
      }
      c += t[N];
      t[N] = (u32)c; t[N + 1] = (u32)(c >> 32);
      u32 m = t[0] * C.n0inv;
      c = (u64)m * C.p[0] + t[0];
      c >>= 32;
      for (int j = 1; j < N; j++) {
        c += (u64)m * C.p[j] + t[j];
        t[j - 1] = (u32)c; c >>= 32;
      }
      c += t[N];
      t[N - 1] = (u32)c;
      t[N] = t[N + 1] + (u32)(c >> 32);
      t[N + 1] = 0;
    }
    fe r, d;
    u32 bw = sub_n<N>(d.v, t, C.p);
    bool ge = t[N] != 0 || bw == 0;
    for (int i = 0; i < N; i++) r.v[i] = ge ? d.v[i] : t[i];
    return r;
  }
  static EB_HD fe sqr(const fe& a, const RtCurve<N>& C) { return mul(a, a, C); }
  static EB_HD fe add(const fe& a, const fe& b, const RtCurve<N>& C) {
    fe r, d;
    u32 cy = add_n<N>(r.v, a.v, b.v);
    u32 bw = sub_n<N>(d.v, r.v, C.p);
    bool ge = cy != 0 || bw == 0;
    for (int i = 0; i < N; i++) r.v[i] = ge ? d.v[i] : r.v[i];
    return r;
  }
  static EB_HD fe sub(const fe& a, const fe& b, const RtCurve<N>& C) {
    fe r, d;
    u32 bw = sub_n<N>(r.v, a.v, b.v);
    add_n<N>(d.v, r.v, C.p);
    for (int i = 0; i < N; i++) r.v[i] = bw ? d.v[i] : r.v[i];
    return r;
  }
  static EB_HD fe dbl(const fe& a, const RtCurve<N>& C) { return add(a, a, C); }
  static EB_HD bool is_zero(const fe& a) { return is_zero_n<N>(a.v); }
  static EB_HD bool eq(const fe& a, const fe& b) { return eq_n<N>(a.v, b.v); }
  static EB_HD fe one(const RtCurve<N>& C) { fe r; for (int i = 0; i < N; i++) r.v[i] = C.r1[i]; return r; }
  static EB_HD fe to_mont(const fe& a, const RtCurve<N>& C) { fe r2; for (int i = 0; i < N; i++) r2.v[i] = C.r2[i]; return mul(a, r2, C); }
  static EB_HD fe from_mont(const fe& a, const RtCurve<N>& C) { fe o = zero(); o.v[0] = 1; return mul(a, o, C); }
  // a^(p-2)
  static EB_HD fe inv(const fe& a, const RtCurve<N>& C) {
    u32 e[N], two[N];
    for (int i = 0; i < N; i++) { e[i] = C.p[i]; two[i] = i == 0 ? 2u : 0u; }
    sub_n<N>(e, e, two);
    fe r = one(C);
    bool started = false;
    for (int i = 32 * N - 1; i >= 0; i--) {
      if (started) r = sqr(r, C);
      if ((e[i >> 5] >> (i & 31)) & 1) { r = started ? mul(r, a, C) : a; started = true; }
    }
    return r;
  }
};

template <int N>
struct RtG {
  typedef RtF<N> F;
  typedef rfe<N> fe;
  struct jac { fe x, y, z; };       // infinity <=> z == 0

  static EB_HD jac infinity(const RtCurve<N>& C) { jac r; r.x = F::one(C); r.y = F::one(C); r.z = F::zero(); return r; }
  // wire x || y (len bytes each, big-endian; values >= p are reduced like toRed) -> Montgomery affine
  static EB_HD void load_xy(fe* x, fe* y, const uint8_t* p, const RtCurve<N>& C) {
    fe rx, ry;
    load_be_len<N>(rx.v, p, (int)C.len);
    load_be_len<N>(ry.v, p + C.len, (int)C.len);
    *x = F::to_mont(rx, C);
    *y = F::to_mont(ry, C);
  }
  // y^2 == x^3 + a x + b  (ShortCurve.validate, short.js:206-216)
  static EB_HD bool on_curve(const fe& x, const fe& y, const RtCurve<N>& C) {
    fe a, b;
    for (int i = 0; i < N; i++) { a.v[i] = C.a[i]; b.v[i] = C.b[i]; }
    fe rhs = F::add(F::add(F::mul(F::sqr(x, C), x, C), F::mul(a, x, C), C), b, C);
    return F::eq(F::sqr(y, C), rhs);
  }
  // general-a doubling (dbl-2007-bl shape, JPoint._dbl short.js:739-800): infinity and 2-torsion give infinity
  static EB_HD jac dbl(const jac& p, const RtCurve<N>& C) {
    if (F::is_zero(p.z) || F::is_zero(p.y)) return infinity(C);
    fe a;
    for (int i = 0; i < N; i++) a.v[i] = C.a[i];
    fe xx = F::sqr(p.x, C), yy = F::sqr(p.y, C), yyyy = F::sqr(yy, C), zz = F::sqr(p.z, C);
    fe s = F::dbl(F::sub(F::sub(F::sqr(F::add(p.x, yy, C), C), xx, C), yyyy, C), C);
    fe m = F::add(F::add(F::dbl(xx, C), xx, C), F::mul(a, F::sqr(zz, C), C), C);
    jac r;
    r.x = F::sub(F::sqr(m, C), F::dbl(s, C), C);
    fe y8 = F::dbl(F::dbl(F::dbl(yyyy, C), C), C);
    r.y = F::sub(F::mul(m, F::sub(s, r.x, C), C), y8, C);
    r.z = F::sub(F::sub(F::sqr(F::add(p.y, p.z, C), C), yy, C), zz, C);
    return r;
  }
  // general addition with every exceptional case (JPoint.add, short.js:532-567)
  static EB_HD jac add(const jac& a, const jac& b, const RtCurve<N>& C) {
    if (F::is_zero(a.z)) return b;
    if (F::is_zero(b.z)) return a;
    fe bz2 = F::sqr(b.z, C), az2 = F::sqr(a.z, C);
    fe u1 = F::mul(a.x, bz2, C), u2 = F::mul(b.x, az2, C);
    fe s1 = F::mul(a.y, F::mul(bz2, b.z, C), C), s2 = F::mul(b.y, F::mul(az2, a.z, C), C);
    fe h = F::sub(u1, u2, C), rr = F::sub(s1, s2, C);
    if (F::is_zero(h)) return F::is_zero(rr) ? dbl(a, C) : infinity(C);
    fe h2 = F::sqr(h, C), h3 = F::mul(h2, h, C), v = F::mul(u1, h2, C);
    jac r;
    r.x = F::sub(F::sub(F::add(F::sqr(rr, C), h3, C), v, C), v, C);
    r.y = F::sub(F::mul(rr, F::sub(v, r.x, C), C), F::mul(s1, h3, C), C);
    r.z = F::mul(F::mul(a.z, b.z, C), h, C);
    return r;
  }
  // k * P, k: klen bytes big-endian (any value), left-to-right double-and-add
  static EB_HD jac mul(const uint8_t* k, u32 klen, const jac& P, const RtCurve<N>& C) {
    jac acc = infinity(C);
    for (u32 i = 0; i < klen; i++) {
      u32 byte = k[i];
      for (int b = 7; b >= 0; b--) {
        acc = dbl(acc, C);
        if ((byte >> b) & 1) acc = add(acc, P, C);
      }
    }
    return acc;
  }
  // Jacobian -> wire affine; returns ST_INFINITY (7) for the point at infinity, else ST_TRUE (1)
  static EB_HD uint8_t store(uint8_t* out, const jac& a, const RtCurve<N>& C) {
    for (u32 b = 0; b < 2 * C.len; b++) out[b] = 0;
    if (F::is_zero(a.z)) return 7;
    fe zi = F::inv(a.z, C), zi2 = F::sqr(zi, C);
    fe x = F::from_mont(F::mul(a.x, zi2, C), C), y = F::from_mont(F::mul(F::mul(a.y, zi2, C), zi, C), C);
    store_be_len<N>(out, x.v, (int)C.len);
    store_be_len<N>(out + C.len, y.v, (int)C.len);
    return 1;
  }
  // op 0: k1*P1 (+ k2*P2 when k2 != NULL);  1: P1 + P2;  2: 2*P1;  3: validate(P1) (status 1 / 0, no output)
  static EB_HD uint8_t item(int op, size_t i, const uint8_t* k1, const uint8_t* p1, const uint8_t* k2, const uint8_t* p2, u32 klen,
                            uint8_t* out, const RtCurve<N>& C) {
    const size_t pl = 2 * (size_t)C.len;
    fe x1, y1;
    load_xy(&x1, &y1, p1 + pl * i, C);
    bool ok1 = on_curve(x1, y1, C);
    if (op == 3) return ok1 ? 1 : 0;
    if (!ok1) { for (size_t b = 0; b < pl; b++) out[pl * i + b] = 0; return 4; }
    jac P1; P1.x = x1; P1.y = y1; P1.z = F::one(C);
    jac R;
    if (op == 2) R = dbl(P1, C);
    else {
      jac P2 = infinity(C);
      if (p2) {
        fe x2, y2;
        load_xy(&x2, &y2, p2 + pl * i, C);
        if (!on_curve(x2, y2, C)) { for (size_t b = 0; b < pl; b++) out[pl * i + b] = 0; return 4; }
        P2.x = x2; P2.y = y2; P2.z = F::one(C);
      }
      if (op == 1) R = add(P1, P2, C);
      else {
        R = mul(k1 + (size_t)klen * i, klen, P1, C);
        if (k2) R = add(R, mul(k2 + (size_t)klen * i, klen, P2, C), C);
      }
    }
    return store(out + pl * i, R, C);
  }
};

}  // namespace eb
