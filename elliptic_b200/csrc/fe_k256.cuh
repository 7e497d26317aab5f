// fe_k256.cuh -- arithmetic in GF(p), p = 2^256 - 2^32 - 977 (secp256k1).
//
// Replaces bn.js `Red` over the K256 pseudo-Mersenne prime for the hot path:
//   Red.mul/sqr/add/sub/neg      reference dist/elliptic.js:7106-7175
//   MPrime.ireduce + K256.split/imulK   dist/elliptic.js:6904-6934, 6952-7009
// Design (B200-first, not a port of the 26-bit-limb JS code): 8 x 32-bit limbs
// held in registers, "weakly reduced" values in [0, 2^256) (canonical form is
// produced only where the reference exposes a residue: fe_normalize), products
// folded with 2^256 = 2^32 + 977 (mod p).  One field multiplication is 64
// IMAD.WIDE.U32 for the product + 9 for the fold; the add/sub chains run on
// the ALU pipe in the shadow of the fma pipe.
#pragma once
#include "limbs.cuh"
#include "mul_cs.cuh"

// EB_MUL_CS=1: wide products from carry-out-only MACs (mul_cs.cuh) instead of the mad.lo.cc / madc.hi.cc chains
#ifndef EB_MUL_CS
#define EB_MUL_CS 0
#endif

namespace eb {

struct fe { u32 v[8]; };

#define K256_C0 977u  // p = 2^256 - (2^32 + K256_C0)

EB_HD fe fe_zero() { fe r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
EB_HD fe fe_one() { fe r = fe_zero(); r.v[0] = 1; return r; }

// r += k * (2^32 + 977) over the full 8 limbs, k in {0,1}; returns carry out.
EB_HD u32 fe_addc_k(u32* r, u32 k) {
  u32 t[8] = {k ? K256_C0 : 0u, k, 0, 0, 0, 0, 0, 0};
  return add_n<8>(r, r, t);
}

#if defined(__CUDA_ARCH__)
// Device reduction: lo + hi*(2^32 + 977) as two 3-address carry chains (the even-indexed limbs of hi
// fold onto lo, the odd-indexed ones onto hi << 32), one merge, then the 33-bit top folded again.
// 9 fused MACs + ~32 adds instead of the ~57 instructions the portable body compiles to (the
// reduction is ~30% of all instructions in the verify kernel, ncu r01).  One PTX instruction per asm
// statement (see fp_mont.cuh for why).
#define EB_MADLO_CC(d, a, b, c) asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define EB_MADCLO_CC(d, a, b, c) asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define EB_MADCHI_CC(d, a, b, c) asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define EB_ADD_CC(d, a, b) asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define EB_ADDC_CC(d, a, b) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define EB_ADDC(d, a, b) asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
EB_D void fe_reduce512_ptx(u32* r, const u32* t) {
  const u32 K = K256_C0, Z = 0;
  u32 E[9], O[10], A[10];
  EB_MADLO_CC(E[0], t[8], K, t[0]);   EB_MADCHI_CC(E[1], t[8], K, t[1]);
  EB_MADCLO_CC(E[2], t[10], K, t[2]); EB_MADCHI_CC(E[3], t[10], K, t[3]);
  EB_MADCLO_CC(E[4], t[12], K, t[4]); EB_MADCHI_CC(E[5], t[12], K, t[5]);
  EB_MADCLO_CC(E[6], t[14], K, t[6]); EB_MADCHI_CC(E[7], t[14], K, t[7]);
  EB_ADDC(E[8], Z, Z);
  EB_MADLO_CC(O[1], t[9], K, t[8]);    EB_MADCHI_CC(O[2], t[9], K, t[9]);
  EB_MADCLO_CC(O[3], t[11], K, t[10]); EB_MADCHI_CC(O[4], t[11], K, t[11]);
  EB_MADCLO_CC(O[5], t[13], K, t[12]); EB_MADCHI_CC(O[6], t[13], K, t[13]);
  EB_MADCLO_CC(O[7], t[15], K, t[14]); EB_MADCHI_CC(O[8], t[15], K, t[15]);
  EB_ADDC(O[9], Z, Z);
  A[0] = E[0];
  EB_ADD_CC(A[1], E[1], O[1]);
#pragma unroll
  for (int k = 2; k < 9; k++) EB_ADDC_CC(A[k], E[k], O[k]);
  EB_ADDC(A[9], O[9], Z);                                   // A < 2^289: A[9] is 0 or 1
  // r = A[0..7] + (A[8] + 2^32 A[9]) * (977 + 2^32)
  u32 c1, c2, c3 = 0;
  EB_MADLO_CC(r[0], A[8], K, A[0]); EB_MADCHI_CC(r[1], A[8], K, A[1]);
#pragma unroll
  for (int k = 2; k < 8; k++) EB_ADDC_CC(r[k], A[k], Z);
  EB_ADDC(c1, Z, Z);
  EB_ADD_CC(r[1], r[1], A[8]);
  EB_ADDC_CC(r[2], r[2], A[9]);
#pragma unroll
  for (int k = 3; k < 8; k++) EB_ADDC_CC(r[k], r[k], Z);
  EB_ADDC(c2, Z, Z);
  if (A[9]) {                                               // only for operands within ~2^10 of 2^256
    EB_ADD_CC(r[1], r[1], K);
#pragma unroll
    for (int k = 2; k < 8; k++) EB_ADDC_CC(r[k], r[k], Z);
    EB_ADDC(c3, Z, Z);
  }
  u32 k = c1 + c2 + c3;                                     // at most one wrap in total
  u32 kK = k * K;
  EB_ADD_CC(r[0], r[0], kK);
  EB_ADDC_CC(r[1], r[1], k);
  EB_ADDC(r[2], r[2], Z);
}
#endif

// Fold a 512-bit value t[16] to 8 limbs in [0, 2^256).
EB_HD void fe_reduce512(u32* r, const u32* t) {
#if defined(__CUDA_ARCH__) && !defined(EB_REDUCE_C)
  fe_reduce512_ptx(r, t);
  return;
#endif
  u32 A[10];
  u64 c = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {  // A = lo + hi*977   (9 limbs)
    c += (u64)t[8 + j] * K256_C0 + t[j];
    A[j] = (u32)c;
    c >>= 32;
  }
  A[8] = (u32)c;
  A[9] = add_n<8>(A + 1, A + 1, t + 8);  // += hi << 32
  // second fold: top = A[8] + 2^32*A[9]  (< 2^33 + 2^11)
  c = (u64)A[8] * K256_C0 + A[0];
  r[0] = (u32)c; c >>= 32;
  c += (u64)A[1] + A[8] + (A[9] ? K256_C0 : 0u);
  r[1] = (u32)c; c >>= 32;
  c += (u64)A[2] + A[9];
  r[2] = (u32)c; c >>= 32;
#pragma unroll
  for (int j = 3; j < 8; j++) {
    c += A[j];
    r[j] = (u32)c; c >>= 32;
  }
  // c in {0,1}; if 1 the wrapped value is < 2^66 so adding 2^32+977 cannot carry past limb 2
  u32 k = (u32)c;
  c = (u64)r[0] + (k ? K256_C0 : 0u);
  r[0] = (u32)c; c >>= 32;
  c += (u64)r[1] + k;
  r[1] = (u32)c; c >>= 32;
  r[2] += (u32)c;
}

// Squaring: 36 MACs (28 off-diagonal products once, doubled, + 8 squares) as PTX
// carry chains generated by tools/gen_sqr.py; the host-emulation build uses the
// general product.
#if defined(__CUDACC__) && !defined(EB_SQR8_INCLUDED)
#define EB_SQR8_INCLUDED
#include "sqr_gen.inc"
#endif
EB_HD void fe_sqr_wide(u32* r, const u32* a) {
#if defined(__CUDA_ARCH__) && EB_MUL_CS
  sqr_wide_cs<8>(r, a);
#elif defined(__CUDA_ARCH__) && !defined(EB_SQR_AS_MUL)
  sqr_wide8_ptx(r, a);
#else
  mul_wide<8>(r, a, a);
#endif
}

EB_HD fe fe_mul_inl(const fe& a, const fe& b) {
  u32 t[16];
#if defined(__CUDA_ARCH__) && EB_MUL_CS
  mul_wide_cs<8>(t, a.v, b.v);
#else
  mul_wide<8>(t, a.v, b.v);
#endif
  fe r;
  fe_reduce512(r.v, t);
  return r;
}

EB_HD fe fe_sqr_inl(const fe& a) {
  u32 t[16];
  fe_sqr_wide(t, a.v);
  fe r;
  fe_reduce512(r.v, t);
  return r;
}

// The multiplier is ~130 instructions (2 KB).  Inlining it at every use makes the
// double/add loop ~43 KB, which thrashes the instruction cache (ncu r01: stall
// "no_instruction" was the top stall).  Out-of-line, operands in registers, the
// whole hot loop is ~12 KB.
#ifndef EB_FE_OUTLINE
#define EB_FE_OUTLINE 1
#endif
#ifndef EB_FE_SQR_INLINE
#define EB_FE_SQR_INLINE 0     // r01 (27 ms kernel): inlining the squarer in the group-law bodies gained 1.5 %.
#endif                         // r02 (21.7 ms kernel, I-cache hit rate 89 %): out of line is 0.5 % faster (21.63 vs
                               // 21.75 ms) and the kernel shrinks from 78 KB to 53 KB -- out of line it is.
#if defined(__CUDACC__) && EB_FE_OUTLINE
__host__ __device__ __noinline__ fe fe_mul(fe a, fe b) { return fe_mul_inl(a, b); }
__host__ __device__ __noinline__ fe fe_sqr(fe a) { return fe_sqr_inl(a); }
// the group-law bodies (jac_dbl_inl / jac_madd_inl) take the squarer inline: 8 of their 18 products then need
// no call marshalling; everything else (inversion and square-root chains) keeps the out-of-line copy
#if EB_FE_SQR_INLINE
EB_HD fe fe_sqr_hot(const fe& a) { return fe_sqr_inl(a); }
#else
EB_HD fe fe_sqr_hot(const fe& a) { return fe_sqr(a); }
#endif
#else
EB_HD fe fe_mul(const fe& a, const fe& b) { return fe_mul_inl(a, b); }
EB_HD fe fe_sqr(const fe& a) { return fe_sqr_inl(a); }
EB_HD fe fe_sqr_hot(const fe& a) { return fe_sqr_inl(a); }
#endif

// Two independent products per call (EB_FE_MUL2=1): half the calls of the group-law bodies, and two carry chains
// for the scheduler to interleave inside one body.
#ifndef EB_FE_MUL2
#define EB_FE_MUL2 0
#endif
struct fe2 { fe a, b; };
#if defined(__CUDACC__) && EB_FE_OUTLINE
__host__ __device__ __noinline__ fe2 fe_mul2(fe a, fe b, fe c, fe d) { fe2 r; r.a = fe_mul_inl(a, b); r.b = fe_mul_inl(c, d); return r; }
#else
EB_HD fe2 fe_mul2(const fe& a, const fe& b, const fe& c, const fe& d) { fe2 r; r.a = fe_mul_inl(a, b); r.b = fe_mul_inl(c, d); return r; }
#endif

#if defined(__CUDA_ARCH__)
// Device add / sub: the 2^256 wrap touches only limbs 0..1 unless a carry leaves limb 1 (probability
// 2^-32 on random data) -- that propagation, and the second wrap behind it, live in a cold branch.
EB_D fe fe_add_ptx(const fe& a, const fe& b) {
  fe r;
  const u32 Z = 0;
  u32 cy, c2;
  EB_ADD_CC(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) EB_ADDC_CC(r.v[i], a.v[i], b.v[i]);
  EB_ADDC(cy, Z, Z);
  u32 kK = (0u - cy) & K256_C0;          // mask, not a multiply: IMAD would take the multiplier pipe
  EB_ADD_CC(r.v[0], r.v[0], kK);
  EB_ADDC_CC(r.v[1], r.v[1], cy);
  EB_ADDC(c2, Z, Z);
  if (c2) {
    u32 c3;
    EB_ADD_CC(r.v[2], r.v[2], c2);
#pragma unroll
    for (int i = 3; i < 8; i++) EB_ADDC_CC(r.v[i], r.v[i], Z);
    EB_ADDC(c3, Z, Z);
    u32 k3 = c3 * K256_C0;                 // wrapped twice: the value is now tiny, no further carry
    EB_ADD_CC(r.v[0], r.v[0], k3);
    EB_ADDC(r.v[1], r.v[1], c3);
  }
  return r;
}
EB_D fe fe_sub_ptx(const fe& a, const fe& b) {
  fe r;
  const u32 Z = 0;
  u32 bw, b2;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r.v[0]) : "r"(a.v[0]), "r"(b.v[0]));
#pragma unroll
  for (int i = 1; i < 8; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r.v[i]) : "r"(a.v[i]), "r"(b.v[i]));
  asm volatile("subc.u32 %0, %1, %1;" : "=r"(bw) : "r"(Z));      // 0 or 0xFFFFFFFF
  u32 kK = bw & K256_C0;
  bw &= 1;
  asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(r.v[0]) : "r"(kK));
  asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(r.v[1]) : "r"(bw));
  asm volatile("subc.u32 %0, %1, %1;" : "=r"(b2) : "r"(Z));
  if (b2) {
    u32 b3;
    asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(r.v[2]) : "r"(1u));
#pragma unroll
    for (int i = 3; i < 8; i++) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(r.v[i]) : "r"(Z));
    asm volatile("subc.u32 %0, %1, %1;" : "=r"(b3) : "r"(Z));
    b3 &= 1;
    u32 k3 = b3 * K256_C0;                 // wrapped below zero twice: value is now >= 2^256 - 2c
    asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(r.v[0]) : "r"(k3));
    asm volatile("subc.u32 %0, %0, %1;" : "+r"(r.v[1]) : "r"(b3));
  }
  return r;
}
#endif

EB_HD fe fe_add(const fe& a, const fe& b) {
#if defined(__CUDA_ARCH__) && !defined(EB_ADDSUB_C)
  return fe_add_ptx(a, b);
#endif
  fe r;
  u32 cy = add_n<8>(r.v, a.v, b.v);
  cy = fe_addc_k(r.v, cy);
  // second wrap only when a,b were both >= 2^256 - c - eps: value now < 2^33
  u64 c = (u64)r.v[0] + (cy ? K256_C0 : 0u);
  r.v[0] = (u32)c; c >>= 32;
  r.v[1] += (u32)c + cy;
  return r;
}

EB_HD fe fe_sub(const fe& a, const fe& b) {
#if defined(__CUDA_ARCH__) && !defined(EB_ADDSUB_C)
  return fe_sub_ptx(a, b);
#endif
  fe r;
  u32 bw = sub_n<8>(r.v, a.v, b.v);
  u32 t[8] = {bw ? K256_C0 : 0u, bw, 0, 0, 0, 0, 0, 0};
  bw = sub_n<8>(r.v, r.v, t);
  // second borrow: value is now >= 2^256 - c, low 64 bits absorb another -c
  u64 lo = ((u64)r.v[1] << 32) | r.v[0];
  lo -= bw ? (((u64)1 << 32) + K256_C0) : 0;
  r.v[0] = (u32)lo; r.v[1] = (u32)(lo >> 32);
  return r;
}

EB_HD fe fe_neg(const fe& a) { return fe_sub(fe_zero(), a); }

// r = a * k for a small constant k (k <= 2^16).
EB_HD fe fe_mul_small(const fe& a, u32 k) {
  fe r;
  u64 c = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    c += (u64)a.v[j] * k;
    r.v[j] = (u32)c; c >>= 32;
  }
  // top = c < k : fold top * (2^32 + 977)
  u32 top = (u32)c;
  u32 t[8] = {0, top, 0, 0, 0, 0, 0, 0};
  u64 m = (u64)top * K256_C0;  // < 2^26
  t[0] = (u32)m;
  t[1] += (u32)(m >> 32);
  u32 cy = add_n<8>(r.v, r.v, t);
  u64 c2 = (u64)r.v[0] + (cy ? K256_C0 : 0u);
  r.v[0] = (u32)c2; c2 >>= 32;
  r.v[1] += (u32)c2 + cy;
  return r;
}

EB_HD fe fe_dbl(const fe& a) { return fe_add(a, a); }

// p as limbs
EB_HD void fe_p(u32* p) {
  p[0] = 0xFFFFFC2Fu; p[1] = 0xFFFFFFFEu;
  for (int i = 2; i < 8; i++) p[i] = 0xFFFFFFFFu;
}

// canonical residue in [0,p)  (what bn.js fromRed() exposes)
EB_HD fe fe_normalize(const fe& a) {
  u32 p[8]; fe_p(p);
  fe r;
  u32 bw = sub_n<8>(r.v, a.v, p);
  cmov_n<8>(r.v, a.v, bw != 0);
  return r;
}

// a == 0 (mod p) for a weakly reduced a
EB_HD bool fe_is_zero(const fe& a) {
  u32 o = 0, n = 0xFFFFFFFFu;
#pragma unroll
  for (int i = 2; i < 8; i++) { o |= a.v[i]; n &= a.v[i]; }
  bool z = (o | a.v[0] | a.v[1]) == 0;
  bool isp = (n == 0xFFFFFFFFu) && a.v[1] == 0xFFFFFFFEu && a.v[0] == 0xFFFFFC2Fu;
  return z || isp;
}

EB_HD bool fe_eq(const fe& a, const fe& b) { return fe_is_zero(fe_sub(a, b)); }

EB_HD bool fe_is_odd(const fe& a) { return fe_normalize(a).v[0] & 1; }

EB_HD fe fe_cmov(const fe& a, const fe& b, bool c) {  // c ? b : a
  fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c ? b.v[i] : a.v[i];
  return r;
}

// Load a 32-byte big-endian value and reduce mod p (reference: `toRed` ->
// Red.convertTo, dist:7292-7296; quirk Q2: coordinates >= p are accepted).
EB_HD fe fe_from_be(const uint8_t* p32) {
  fe r;
  load_be<8>(r.v, p32);
  return r;  // already in [0, 2^256): a valid weak representative
}

// a^e for the fixed exponents the path needs, by square-and-multiply on the
// exponent's bits (MSB first).  e is given as 8 limbs.
EB_HD fe fe_pow(const fe& a, const u32* e) {
  fe r = fe_one();
  bool started = false;
  for (int i = 255; i >= 0; i--) {
    if (started) r = fe_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) {
      r = started ? fe_mul(r, a) : a;
      started = true;
    }
  }
  return r;
}

// a^(p-2)  (0 -> 0, matching bn.js _invmp(0) == 0, dist:6568-6579)
EB_HD fe fe_inv(const fe& a) {
  u32 e[8]; fe_p(e);
  e[0] -= 2;
  return fe_pow(a, e);
}

// sqrt candidate a^((p+1)/4)  (Red.sqrt fast path for p = 3 mod 4, dist:7183-7187)
EB_HD fe fe_sqrt_candidate(const fe& a) {
  // (p+1)/4 = 2^254 - 2^30 - 244
  u32 e[8] = {0xBFFFFF0Cu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu,
              0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x3FFFFFFFu};
  return fe_pow(a, e);
}

}  // namespace eb
