// gq_k256.cuh -- secp256k1 group law and the double-scalar core u1*G + u2*Q on the carry-free
// 9 x 29-bit field (fq_pm.cuh).  Same EFD formulas as ge_k256.cuh / the reference's JPoint
// (lib/elliptic/curve/short.js: mixedAdd :569-603, _zeroDbl :668-737), re-associated so that every
// product sees operands within the magnitude budget (checked at compile time by the fq types):
//   dbl   3M + 4S: A = X^2, B2 = 2 Y^2, C4 = B2^2 (= 4 Y^4), XB2 = X * B2 (= 2 X Y^2), E = 3A,
//         X3 = E^2 - 4 XB2, Y3 = E (2 XB2 - X3) - 2 C4, Z3 = 2 Y Z           (dbl-2009-l, D = 2 XB2)
//   madd  8M + 3S (madd-2007-bl shape used by the reference), exceptional cases in a cold path that
//         falls back to the packed-field code so they get exactly the reference's answers.
#pragma once
#include "fq_pm.cuh"
#include "ge_k256.cuh"

namespace eb {

template <int M> using fqk = fq<PmK256, M>;

struct gq_jac { fqk<1> x, y; fqk<2> z; };     // infinity <=> z == 0 (mod p)
struct gq_aff { fqk<1> x; fqk<2> y; };        // y may be a lazily negated table value

EB_HD fqk<1> fqk_from_fe(const fe& a) { return fq_from_words<PmK256>(a.v); }
EB_HD fe fqk_to_fe(const fqk<1>& a) { fe r; fq_to_words<PmK256>(r.v, fq_canon(a)); return r; }
template <int M> EB_HD fe fqk_to_fe_m(const fqk<M>& a) { fe r; fq_to_words<PmK256>(r.v, fq_canon(a)); return r; }
EB_HD fqk<1> fqk_one() { fqk<1> r; for (int i = 0; i < FQ_L; i++) r.v[i] = 0; r.v[0] = 1; return r; }
EB_HD fqk<1> fqk_zero() { fqk<1> r; for (int i = 0; i < FQ_L; i++) r.v[i] = 0; return r; }

EB_HD gq_jac gq_from_jac(const ge_jac& p) {
  gq_jac r; r.x = fqk_from_fe(p.x); r.y = fqk_from_fe(p.y); r.z = fqk_from_fe(p.z); return r;
}
EB_HD ge_jac gq_to_jac(const gq_jac& p) {
  ge_jac r; r.x = fqk_to_fe(p.x); r.y = fqk_to_fe(p.y); r.z = fqk_to_fe_m(p.z); return r;
}
EB_HD gq_jac gq_from_aff(const gq_aff& p) {
  gq_jac r; r.x = p.x; r.y = fq_weak(p.y); r.z = fqk_one(); return r;
}

// 2P (a = 0).  Infinity in -> infinity out (Z3 = 2 Y Z).
EB_HD gq_jac gq_dbl_inl(const gq_jac& p) {
  fqk<1> A = fq_sqr(p.x);
  fqk<2> B2 = fq_mul_int<2>(fq_sqr(p.y));
  fqk<1> C4 = fq_sqr(B2);
  fqk<1> XB2 = fq_mul(p.x, B2);
  fqk<1> E = fq_weak(fq_mul_int<3>(A));
  fqk<1> F = fq_sqr(E);
  gq_jac r;
  r.x = fq_weak(fq_sub(F, fq_mul_int<4>(XB2)));
  fqk<1> m = fq_mul(E, fq_sub(fq_mul_int<2>(XB2), r.x));
  r.y = fq_weak(fq_sub(m, fq_mul_int<2>(C4)));
  r.z = fq_mul_int<2>(fq_mul(p.y, p.z));
  return r;
}

// acc + P, P affine; all cases exact.
EB_HD gq_jac gq_madd_inl(const gq_jac& a, const gq_aff& p) {
  fqk<1> z2 = fq_sqr(a.z);
  fqk<1> u2 = fq_mul(p.x, z2);
  fqk<1> s2 = fq_mul(fq_mul(p.y, z2), a.z);
  fqk<1> h = fq_weak(fq_sub(a.x, u2));
  fqk<1> rr = fq_weak(fq_sub(a.y, s2));
  fqk<1> h2 = fq_sqr(h);
  fqk<1> h3 = fq_mul(h2, h);
  fqk<1> v = fq_mul(a.x, h2);
  gq_jac r;
  r.x = fq_weak(fq_sub(fq_add(fq_sqr(rr), h3), fq_mul_int<2>(v)));
  r.y = fq_weak(fq_sub(fq_mul(rr, fq_sub(v, r.x)), fq_mul(a.y, h3)));
  fqk<1> z3 = fq_mul(a.z, h);
  r.z = z3;
  if (fq_is_zero(z3)) {                          // cold: acc == O, or h == 0 (P + P, P + (-P))
    ge_aff pa; pa.x = fqk_to_fe(p.x); pa.y = fqk_to_fe_m(p.y);
    return gq_from_jac(jac_madd_inl(gq_to_jac(a), pa));
  }
  return r;
}

#if defined(__CUDACC__)
#define EB_GQ_FN __host__ __device__ __noinline__
#else
#define EB_GQ_FN
#endif
#ifndef EB_GQ_OUTLINE
#define EB_GQ_OUTLINE 1
#endif
#if EB_GQ_OUTLINE
EB_GQ_FN gq_jac gq_dbl(gq_jac p) { return gq_dbl_inl(p); }
EB_GQ_FN gq_jac gq_madd(gq_jac a, gq_aff p) { return gq_madd_inl(a, p); }
#else
EB_HD gq_jac gq_dbl(const gq_jac& p) { return gq_dbl_inl(p); }
EB_HD gq_jac gq_madd(const gq_jac& a, const gq_aff& p) { return gq_madd_inl(a, p); }
#endif

// madd that also returns h with Z3 = Z1 * h (table build; inputs never exceptional: odd multiples of an
// on-curve point of prime order plus 2Q).
struct gq_madd_out { gq_jac r; fqk<1> h; };
EB_HD gq_madd_out gq_madd_h(const gq_jac& a, const fqk<1>& px, const fqk<1>& py) {
  gq_madd_out o;
  fqk<1> z2 = fq_sqr(a.z);
  fqk<1> u2 = fq_mul(px, z2);
  fqk<1> s2 = fq_mul(fq_mul(py, z2), a.z);
  fqk<1> h = fq_weak(fq_sub(a.x, u2));
  fqk<1> rr = fq_weak(fq_sub(a.y, s2));
  fqk<1> h2 = fq_sqr(h);
  fqk<1> h3 = fq_mul(h2, h);
  fqk<1> v = fq_mul(a.x, h2);
  o.r.x = fq_weak(fq_sub(fq_add(fq_sqr(rr), h3), fq_mul_int<2>(v)));
  o.r.y = fq_weak(fq_sub(fq_mul(rr, fq_sub(v, o.r.x)), fq_mul(a.y, h3)));
  o.r.z = fq_mul(a.z, h);
  o.h = h;
  return o;
}

// per-item table entry: x (9 limbs, padded to 12 words so that 128-bit loads stay aligned), y, beta*x
constexpr int FQ_PAD = 12;
constexpr int QTABQ_ENTRY_WORDS = 3 * FQ_PAD;
constexpr int QTABQ_WORDS = 8 * QTABQ_ENTRY_WORDS;     // 288 words = 1152 B per item

template <int M>
EB_HD void fq_store12(u32* dst, const fqk<M>& a) {
#if defined(__CUDA_ARCH__)
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  d4[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  d4[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
  d4[2] = make_uint4(a.v[8], 0u, 0u, 0u);
#else
  for (int i = 0; i < FQ_L; i++) dst[i] = a.v[i];
  dst[9] = dst[10] = dst[11] = 0;
#endif
}
EB_HD fqk<1> fq_load12(const u32* src) {
  fqk<1> a;
#if defined(__CUDA_ARCH__)
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4 q0 = s4[0], q1 = s4[1];
  a.v[0] = q0.x; a.v[1] = q0.y; a.v[2] = q0.z; a.v[3] = q0.w;
  a.v[4] = q1.x; a.v[5] = q1.y; a.v[6] = q1.z; a.v[7] = q1.w;
  a.v[8] = src[8];
#else
  for (int i = 0; i < FQ_L; i++) a.v[i] = src[i];
#endif
  return a;
}

}  // namespace eb
