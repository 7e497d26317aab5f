// ge_k256.cuh -- secp256k1 group law on Jacobian / affine points.
//
// Replaces the reference's JPoint arithmetic for the a = 0 curve
// (lib/elliptic/curve/short.js: add :532-567, mixedAdd :569-603,
// _zeroDbl :668-737, isInfinity :935-938, eqXToP :908-925).  The formulas are
// the same EFD ones (dbl-2009-l, 8M+3S mixed add); what differs is the
// handling of the exceptional cases: the common path is branch-free and the
// rare cases (accumulator at infinity, P + P, P + (-P)) are detected from
// Z3 == 0 and resolved exactly in a cold path, so that every lane of a warp
// runs the same instruction stream for well-formed inputs while adversarial
// inputs still get the exact group-law answer the reference computes.
#pragma once
#include "fe_k256.cuh"

namespace eb {

struct ge_aff { fe x, y; };          // affine, never infinity
struct ge_jac { fe x, y, z; };       // Jacobian; infinity <=> z == 0 (mod p)

#if defined(__CUDACC__)
#define EB_NOINLINE __noinline__
#else
#define EB_NOINLINE
#endif

EB_HD ge_jac jac_infinity() {
  ge_jac r; r.x = fe_one(); r.y = fe_one(); r.z = fe_zero(); return r;
}
EB_HD bool jac_is_infinity(const ge_jac& a) { return fe_is_zero(a.z); }

EB_HD ge_jac jac_from_aff(const ge_aff& p) {
  ge_jac r; r.x = p.x; r.y = p.y; r.z = fe_one(); return r;
}

// 2*P, a = 0: dbl-2009-l (2M + 5S).  Infinity in -> infinity out (Z3 = 2*Y*Z).
// (short.js:697-733.)
EB_HD ge_jac jac_dbl_inl(const ge_jac& p) {
  fe A = fe_sqr_hot(p.x);
  fe B = fe_sqr_hot(p.y);
  fe C = fe_sqr_hot(B);
  fe t = fe_add(p.x, B);
  t = fe_sqr_hot(t);
  t = fe_sub(t, A);
  t = fe_sub(t, C);
  fe D = fe_dbl(t);
  fe E = fe_mul_small(A, 3);
  fe F = fe_sqr_hot(E);
  ge_jac r;
  r.x = fe_sub(F, fe_dbl(D));
  fe C8 = fe_mul_small(C, 8);
#if EB_FE_MUL2
  fe2 m = fe_mul2(E, fe_sub(D, r.x), p.y, p.z);
  r.y = fe_sub(m.a, C8);
  r.z = fe_dbl(m.b);
#else
  r.y = fe_sub(fe_mul(E, fe_sub(D, r.x)), C8);
  r.z = fe_dbl(fe_mul(p.y, p.z));
#endif
  return r;
}

// Exact affine doubling result as a Jacobian point (used by the cold path).
EB_HD ge_jac jac_dbl_aff(const ge_aff& p) { return jac_dbl_inl(jac_from_aff(p)); }

// acc + P for Jacobian acc and affine P (8M + 3S), all cases exact.
// (short.js:569-603.)
EB_HD ge_jac jac_madd_inl(const ge_jac& a, const ge_aff& p) {
  fe z2 = fe_sqr_hot(a.z);
#if EB_FE_MUL2
  fe2 m1 = fe_mul2(p.x, z2, p.y, z2);            // u2, y2 z1^2
  fe h = fe_sub(a.x, m1.a);
  fe2 m2 = fe_mul2(m1.b, a.z, a.z, h);           // s2, z3
  fe rr = fe_sub(a.y, m2.a);
  fe h2 = fe_sqr_hot(h);
  fe2 m3 = fe_mul2(h2, h, a.x, h2);              // h^3, v
  fe h3 = m3.a, v = m3.b;
  ge_jac r;
  r.x = fe_sub(fe_sub(fe_add(fe_sqr_hot(rr), h3), v), v);
  fe2 m4 = fe_mul2(rr, fe_sub(v, r.x), a.y, h3);
  r.y = fe_sub(m4.a, m4.b);
  r.z = m2.b;
#else
  fe u2 = fe_mul(p.x, z2);
  fe s2 = fe_mul(fe_mul(p.y, z2), a.z);
  fe h = fe_sub(a.x, u2);
  fe rr = fe_sub(a.y, s2);
  fe h2 = fe_sqr_hot(h);
  fe h3 = fe_mul(h2, h);
  fe v = fe_mul(a.x, h2);
  ge_jac r;
  r.x = fe_sub(fe_sub(fe_add(fe_sqr_hot(rr), h3), v), v);
  r.y = fe_sub(fe_mul(rr, fe_sub(v, r.x)), fe_mul(a.y, h3));
  r.z = fe_mul(a.z, h);
#endif
  if (fe_is_zero(r.z)) {                       // cold: a == inf, or h == 0
    if (fe_is_zero(a.z)) return jac_from_aff(p);   // O + P = P      (short.js:571-572)
    if (fe_is_zero(rr)) return jac_dbl_inl(a);     // P + P = 2P     (short.js:591)
    return jac_infinity();                         // P + (-P) = O   (short.js:588-589)
  }
  return r;
}

// acc + Q for two Jacobian points (12M + 4S), all cases exact (short.js:532-567).
EB_HD ge_jac jac_add_inl(const ge_jac& a, const ge_jac& b) {
  if (fe_is_zero(a.z)) return b;
  if (fe_is_zero(b.z)) return a;
  fe bz2 = fe_sqr(b.z);
  fe az2 = fe_sqr(a.z);
  fe u1 = fe_mul(a.x, bz2);
  fe u2 = fe_mul(b.x, az2);
  fe s1 = fe_mul(a.y, fe_mul(bz2, b.z));
  fe s2 = fe_mul(b.y, fe_mul(az2, a.z));
  fe h = fe_sub(u1, u2);
  fe rr = fe_sub(s1, s2);
  if (fe_is_zero(h)) {
    if (fe_is_zero(rr)) return jac_dbl_inl(a);
    return jac_infinity();
  }
  fe h2 = fe_sqr(h);
  fe h3 = fe_mul(h2, h);
  fe v = fe_mul(u1, h2);
  ge_jac r;
  r.x = fe_sub(fe_sub(fe_add(fe_sqr(rr), h3), v), v);
  r.y = fe_sub(fe_mul(rr, fe_sub(v, r.x)), fe_mul(s1, h3));
  r.z = fe_mul(fe_mul(a.z, b.z), h);
  return r;
}

EB_HD ge_aff aff_neg_if(const ge_aff& p, bool neg) {
  ge_aff r; r.x = p.x; r.y = fe_cmov(p.y, fe_neg(p.y), neg); return r;
}

// y^2 == x^3 + 7 ?   (ShortCurve.validate, short.js:206-216)
EB_HD bool aff_on_curve(const ge_aff& p) {
  fe x3 = fe_mul(fe_sqr(p.x), p.x);
  fe seven = fe_zero(); seven.v[0] = 7;
  return fe_eq(fe_sqr(p.y), fe_add(x3, seven));
}

// Jacobian -> affine (JPoint.toP, short.js:516-526).  Caller handles infinity.
EB_HD ge_aff jac_to_aff(const ge_jac& a) {
  fe zi = fe_inv(a.z);
  fe zi2 = fe_sqr(zi);
  ge_aff r;
  r.x = fe_mul(a.x, zi2);
  r.y = fe_mul(fe_mul(a.y, zi2), zi);
  return r;
}

}  // namespace eb
