// fe_25519.cuh -- arithmetic in GF(2^255 - 19) for ed25519 / curve25519.
//
// Replaces bn.js `Red` over the P25519 pseudo-Mersenne prime (reference
// dist/elliptic.js:7027-7051 P25519.imulK, 6904-6934 MPrime.ireduce, 7106-7175 Red ops,
// 7177-7232 Red.sqrt (Tonelli-Shanks), 7234-7242 Red.invm) on the EdDSA-verify and
// X25519-style ECDH paths.  8 x 32-bit limbs in registers, values weakly reduced to
// [0, 2^256) (2^256 = 38 mod p); canonical residues only where the reference exposes them.
#pragma once
#include "limbs.cuh"
#if defined(__CUDACC__)
#ifndef EB_SQR8_INCLUDED
#define EB_SQR8_INCLUDED
namespace eb {
#include "sqr_gen.inc"
}
#endif
#endif

namespace eb {

struct f25 { u32 v[8]; };

EB_HD f25 f25_zero() { f25 r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
EB_HD f25 f25_one() { f25 r = f25_zero(); r.v[0] = 1; return r; }
EB_HD f25 f25_small(u32 k) { f25 r = f25_zero(); r.v[0] = k; return r; }

#if defined(__CUDA_ARCH__)
// Device fold as two 3-address carry chains + one merge (see fe_k256.cuh::fe_reduce512_ptx).
EB_D void f25_reduce512_ptx(u32* r, const u32* t) {
  const u32 K = 38u, Z = 0;
  u32 E[9], O[9], A[9];
#define F25_MADLO_CC(d, a, b, c) asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define F25_MADCLO_CC(d, a, b, c) asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define F25_MADCHI_CC(d, a, b, c) asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define F25_ADD_CC(d, a, b) asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define F25_ADDC_CC(d, a, b) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define F25_ADDC(d, a, b) asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
  F25_MADLO_CC(E[0], t[8], K, t[0]);   F25_MADCHI_CC(E[1], t[8], K, t[1]);
  F25_MADCLO_CC(E[2], t[10], K, t[2]); F25_MADCHI_CC(E[3], t[10], K, t[3]);
  F25_MADCLO_CC(E[4], t[12], K, t[4]); F25_MADCHI_CC(E[5], t[12], K, t[5]);
  F25_MADCLO_CC(E[6], t[14], K, t[6]); F25_MADCHI_CC(E[7], t[14], K, t[7]);
  F25_ADDC(E[8], Z, Z);
  F25_MADLO_CC(O[1], t[9], K, Z);   F25_MADCHI_CC(O[2], t[9], K, Z);
  F25_MADCLO_CC(O[3], t[11], K, Z); F25_MADCHI_CC(O[4], t[11], K, Z);
  F25_MADCLO_CC(O[5], t[13], K, Z); F25_MADCHI_CC(O[6], t[13], K, Z);
  F25_MADCLO_CC(O[7], t[15], K, Z); F25_MADCHI_CC(O[8], t[15], K, Z);
  A[0] = E[0];
  F25_ADD_CC(A[1], E[1], O[1]);
#pragma unroll
  for (int k = 2; k < 8; k++) F25_ADDC_CC(A[k], E[k], O[k]);
  F25_ADDC(A[8], E[8], O[8]);                      // A < 39 * 2^256: A[8] <= 38
  u32 c1;
  F25_MADLO_CC(r[0], A[8], K, A[0]);               // A[8]*38 < 2^11: no high part
#pragma unroll
  for (int k = 1; k < 8; k++) F25_ADDC_CC(r[k], A[k], Z);
  F25_ADDC(c1, Z, Z);
  u32 kK = c1 * K;                                  // wrapped value < 2^11: adding 38 stays in limb 0..1
  F25_ADD_CC(r[0], r[0], kK);
  F25_ADDC(r[1], r[1], Z);
}
#endif

// fold a 512-bit value to [0, 2^256): lo + 38*hi, twice
EB_HD void f25_reduce512(u32* r, const u32* t) {
#if defined(__CUDA_ARCH__) && !defined(EB_REDUCE_C)
  f25_reduce512_ptx(r, t);
  return;
#endif
  u32 A[9];
  u64 c = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    c += (u64)t[8 + j] * 38u + t[j];
    A[j] = (u32)c; c >>= 32;
  }
  A[8] = (u32)c;                       // < 38 + 1
  c = (u64)A[8] * 38u + A[0];
  r[0] = (u32)c; c >>= 32;
#pragma unroll
  for (int j = 1; j < 8; j++) {
    c += A[j];
    r[j] = (u32)c; c >>= 32;
  }
  // c in {0,1}: wrapped value < 38*39, adding 38 cannot carry out again
  u32 k = (u32)c;
  c = (u64)r[0] + (k ? 38u : 0u);
  r[0] = (u32)c; c >>= 32;
  r[1] += (u32)c;
}

EB_HD f25 f25_mul_inl(const f25& a, const f25& b) {
  u32 t[16];
  mul_wide<8>(t, a.v, b.v);
  f25 r;
  f25_reduce512(r.v, t);
  return r;
}
EB_HD f25 f25_sqr_inl(const f25& a) {
  u32 t[16];
#if defined(__CUDA_ARCH__)
  sqr_wide8_ptx(t, a.v);
#else
  mul_wide<8>(t, a.v, a.v);
#endif
  f25 r;
  f25_reduce512(r.v, t);
  return r;
}
#ifndef EB_F25_SQR_INLINE
#define EB_F25_SQR_INLINE 1    // squarer inlined in the group-law bodies, multiplier out of line.  Its own switch: the
#endif                         // secp256k1 kernel now prefers its squarer out of line (EB_FE_SQR_INLINE = 0), these
                               // kernels do not (r02: ed25519 verify 33.4 vs 34.3 ms, X25519 24.0 vs 24.7 ms)
#if defined(__CUDACC__)
__host__ __device__ __noinline__ f25 f25_mul(f25 a, f25 b) { return f25_mul_inl(a, b); }
__host__ __device__ __noinline__ f25 f25_sqr(f25 a) { return f25_sqr_inl(a); }
#if EB_F25_SQR_INLINE
EB_HD f25 f25_sqr_hot(const f25& a) { return f25_sqr_inl(a); }     // doubling / ladder step only
#else
EB_HD f25 f25_sqr_hot(const f25& a) { return f25_sqr(a); }
#endif
#else
EB_HD f25 f25_mul(const f25& a, const f25& b) { return f25_mul_inl(a, b); }
EB_HD f25 f25_sqr(const f25& a) { return f25_sqr_inl(a); }
EB_HD f25 f25_sqr_hot(const f25& a) { return f25_sqr_inl(a); }
#endif

#if defined(__CUDA_ARCH__)
// Device add / sub: the 2^256 = 38 wrap touches limb 0 only unless it carries out of it (38 / 2^32 of
// the time); that propagation and the second wrap behind it sit in a cold branch (as fe_k256.cuh).
EB_D f25 f25_add_ptx(const f25& a, const f25& b) {
  f25 r;
  const u32 Z = 0;
  u32 cy, c2;
  F25_ADD_CC(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) F25_ADDC_CC(r.v[i], a.v[i], b.v[i]);
  F25_ADDC(cy, Z, Z);
  u32 kK = (0u - cy) & 38u;
  F25_ADD_CC(r.v[0], r.v[0], kK);
  F25_ADDC(c2, Z, Z);
  if (c2) {
    u32 c3;
    F25_ADD_CC(r.v[1], r.v[1], c2);
#pragma unroll
    for (int i = 2; i < 8; i++) F25_ADDC_CC(r.v[i], r.v[i], Z);
    F25_ADDC(c3, Z, Z);
    r.v[0] += (0u - c3) & 38u;            // wrapped twice: the value is tiny now
  }
  return r;
}
EB_D f25 f25_sub_ptx(const f25& a, const f25& b) {
  f25 r;
  const u32 Z = 0;
  u32 bw, b2;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r.v[0]) : "r"(a.v[0]), "r"(b.v[0]));
#pragma unroll
  for (int i = 1; i < 8; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r.v[i]) : "r"(a.v[i]), "r"(b.v[i]));
  asm volatile("subc.u32 %0, %1, %1;" : "=r"(bw) : "r"(Z));      // 0 or 0xFFFFFFFF
  u32 kK = bw & 38u;
  asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(r.v[0]) : "r"(kK));
  asm volatile("subc.u32 %0, %1, %1;" : "=r"(b2) : "r"(Z));
  if (b2) {
    u32 b3;
    asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(r.v[1]) : "r"(1u));
#pragma unroll
    for (int i = 2; i < 8; i++) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(r.v[i]) : "r"(Z));
    asm volatile("subc.u32 %0, %1, %1;" : "=r"(b3) : "r"(Z));
    r.v[0] -= b3 & 38u;                   // wrapped below zero twice: the low limb is >= 2^32 - 38
  }
  return r;
}
#endif

EB_HD f25 f25_add(const f25& a, const f25& b) {
#if defined(__CUDA_ARCH__)
  return f25_add_ptx(a, b);
#endif
  f25 r;
  u32 cy = add_n<8>(r.v, a.v, b.v);
  u32 t[8] = {cy ? 38u : 0u, 0, 0, 0, 0, 0, 0, 0};
  cy = add_n<8>(r.v, r.v, t);
  r.v[0] += cy ? 38u : 0u;             // second wrap: value is < 38 then
  return r;
}
EB_HD f25 f25_sub(const f25& a, const f25& b) {
#if defined(__CUDA_ARCH__)
  return f25_sub_ptx(a, b);
#endif
  f25 r;
  u32 bw = sub_n<8>(r.v, a.v, b.v);
  u32 t[8] = {bw ? 38u : 0u, 0, 0, 0, 0, 0, 0, 0};
  bw = sub_n<8>(r.v, r.v, t);
  r.v[0] -= bw ? 38u : 0u;             // second borrow: value is >= 2^256 - 38, low limb >= 2^32 - 38
  return r;
}
EB_HD f25 f25_neg(const f25& a) { return f25_sub(f25_zero(), a); }
EB_HD f25 f25_dbl(const f25& a) { return f25_add(a, a); }

// a * k, k < 2^26
EB_HD f25 f25_mul_small(const f25& a, u32 k) {
  f25 r;
  u64 c = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    c += (u64)a.v[j] * k;
    r.v[j] = (u32)c; c >>= 32;
  }
  u64 m = c * 38u;                      // < 2^32
  u32 t[8] = {(u32)m, (u32)(m >> 32), 0, 0, 0, 0, 0, 0};
  u32 cy = add_n<8>(r.v, r.v, t);
  r.v[0] += cy ? 38u : 0u;
  return r;
}

// canonical residue in [0, p), p = 2^255 - 19
EB_HD f25 f25_normalize(const f25& a) {
  f25 r = a;
  // fold bit 255: v = (v mod 2^255) + 19 * (v >> 255)   -> < 2^255 + 19
  u32 top = r.v[7] >> 31;
  r.v[7] &= 0x7FFFFFFFu;
  u32 t[8] = {top ? 19u : 0u, 0, 0, 0, 0, 0, 0, 0};
  add_n<8>(r.v, r.v, t);
  // conditional subtract p (twice is never needed: value < 2^255 + 19 < 2p)
  const u32 p[8] = {0xFFFFFFEDu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x7FFFFFFFu};
  f25 d;
  u32 bw = sub_n<8>(d.v, r.v, p);
  cmov_n<8>(r.v, d.v, bw == 0);
  return r;
}
EB_HD bool f25_is_zero(const f25& a) { f25 n = f25_normalize(a); return is_zero_n<8>(n.v); }
EB_HD bool f25_eq(const f25& a, const f25& b) { return f25_is_zero(f25_sub(a, b)); }
EB_HD bool f25_is_odd(const f25& a) { return f25_normalize(a).v[0] & 1; }
EB_HD f25 f25_cmov(const f25& a, const f25& b, bool c) {
  f25 r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c ? b.v[i] : a.v[i];
  return r;
}

EB_HD f25 f25_sqr_n(f25 a, int n) { for (int i = 0; i < n; i++) a = f25_sqr(a); return a; }

// a^(2^250 - 1) and a^11, the shared stem of the classic curve25519 addition chains
EB_HD f25 f25_pow_2_250_1(const f25& a, f25* a11) {
  f25 z2 = f25_sqr(a);
  f25 z8 = f25_sqr_n(z2, 2);
  f25 z9 = f25_mul(z8, a);
  f25 z11 = f25_mul(z9, z2);
  f25 z22 = f25_sqr(z11);
  f25 z_5_0 = f25_mul(z22, z9);                         // 2^5 - 1
  f25 z_10_0 = f25_mul(f25_sqr_n(z_5_0, 5), z_5_0);
  f25 z_20_0 = f25_mul(f25_sqr_n(z_10_0, 10), z_10_0);
  f25 z_40_0 = f25_mul(f25_sqr_n(z_20_0, 20), z_20_0);
  f25 z_50_0 = f25_mul(f25_sqr_n(z_40_0, 10), z_10_0);
  f25 z_100_0 = f25_mul(f25_sqr_n(z_50_0, 50), z_50_0);
  f25 z_200_0 = f25_mul(f25_sqr_n(z_100_0, 100), z_100_0);
  f25 z_250_0 = f25_mul(f25_sqr_n(z_200_0, 50), z_50_0);
  *a11 = z11;
  return z_250_0;
}
// a^(p-2) = a^(2^255 - 21): inverse; 0 -> 0 (bn.js _invmp(0) == 0, dist:6568-6579)
EB_HD f25 f25_inv(const f25& a) {
  f25 a11;
  f25 t = f25_pow_2_250_1(a, &a11);
  return f25_mul(f25_sqr_n(t, 5), a11);
}
// a^((p-5)/8) = a^(2^252 - 3)
EB_HD f25 f25_pow_p58(const f25& a) {
  f25 a11;
  f25 t = f25_pow_2_250_1(a, &a11);
  return f25_mul(f25_sqr_n(t, 2), a);
}
// a^((p-1)/2) = a^(2^254 - 10): Legendre symbol (0, 1 or p-1)
EB_HD f25 f25_legendre(const f25& a) {
  // 2^254 - 10 = (2^250 - 1) * 2^4 + 6
  f25 a11;
  f25 t = f25_pow_2_250_1(a, &a11);
  f25 a2 = f25_sqr(a);
  f25 a6 = f25_mul(f25_sqr(a2), a2);
  return f25_mul(f25_sqr_n(t, 4), a6);
}

// a^((p-3)/2) = a^(2^254 - 11) = (a^(2^250 - 1))^(2^4) * a^5.  For a != 0: a * this = chi(a) (Legendre symbol) and
// chi(a) * this = 1 / a -- one exponentiation answers "is a a square" AND inverts it.
EB_HD f25 f25_pow_p32(const f25& a) {
  f25 a11;
  f25 t = f25_pow_2_250_1(a, &a11);
  f25 a5 = f25_mul(f25_sqr_n(a, 2), a);
  return f25_mul(f25_sqr_n(t, 4), a5);
}

EB_HD f25 f25_sqrt_m1() {   // 2^((p-1)/4)
  f25 r;
  const u32 v[8] = {0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u, 0x3dfbd7a7u, 0x2b4d0099u, 0x4fc1df0bu, 0x2b832480u};
  for (int i = 0; i < 8; i++) r.v[i] = v[i];
  return r;
}
EB_HD f25 f25_d() {         // ed25519 d (curves.js:157)
  f25 r;
  const u32 v[8] = {0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du, 0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu};
  for (int i = 0; i < 8; i++) r.v[i] = v[i];
  return r;
}
EB_HD f25 f25_2d() {
  f25 r;
  const u32 v[8] = {0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au, 0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu};
  for (int i = 0; i < 8; i++) r.v[i] = v[i];
  return r;
}

EB_HD f25 f25_load(const u32* src) { f25 a; for (int i = 0; i < 8; i++) a.v[i] = src[i]; return a; }
EB_HD void f25_store(u32* dst, const f25& a) { for (int i = 0; i < 8; i++) dst[i] = a.v[i]; }

}  // namespace eb
