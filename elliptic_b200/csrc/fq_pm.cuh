// fq_pm.cuh -- carry-free arithmetic in GF(p) for the two pseudo-Mersenne primes of the path:
//   secp256k1  p = 2^256 - 2^32 - 977       (bn.js `Red` over K256,   dist/elliptic.js:6944-7009)
//   25519      p = 2^255 - 19               (bn.js `Red` over P25519, dist/elliptic.js:7027-7051)
// It replaces Red.mul/sqr/add/sub/neg (dist/elliptic.js:7106-7175) inside the hot double/add loops.
//
// Why a second field layer (fe_k256.cuh keeps the packed 8x32 form for everything off the hot loop):
// measured on B200 (profiles/r01_imad_peak.json), IMAD.WIDE.U32 with a carry predicate in or out
// issues at HALF the rate of the plain form, so a 32-bit-limb schoolbook product (64 carry-chained
// MACs) costs 64 x 4 issue cycles per warp.  Here an element is 9 limbs of 29 bits in 32-bit
// registers: column sums of up to 9 products (58 bits each) fit a 64-bit accumulator, so every MAC is
// a plain IMAD.WIDE.U32 (81 x 2 cycles), carries are shifts on the ALU pipe, and add / sub are 9
// independent 32-bit adds with no carry chain and no reduction ("lazy" limbs, 3 spare bits).
//
// Magnitude discipline (checked at COMPILE time): fq<P, M> promises every limb <= M * U with
// U = 2^29 + 2^12 (top limb <= M * (2^TOPBITS + 2^12)).  mul / sqr need MA * MB <= 7 so that
// 9 * MA * MB * U^2 + carries < 2^64; add gives MA + MB, sub gives MA + MB + 1 (it adds (MB + 1) * p);
// weak() brings any M <= 7 back to 1.  A formula that could overflow does not compile.
#pragma once
#include "limbs.cuh"

namespace eb {

struct PmK256 {
  static constexpr int TOPBITS = 24;                 // 8 * 29 + 24 = 256
  static constexpr u32 R0 = 31264u, R1 = 256u;       // 2^261 = R0 + R1 * 2^29   (mod p)
  static constexpr u32 T0 = 977u, T1 = 8u;           // 2^256 = T0 + T1 * 2^29   (mod p)
};
struct Pm25519 {
  static constexpr int TOPBITS = 23;                 // 8 * 29 + 23 = 255
  static constexpr u32 R0 = 1216u, R1 = 0u;          // 2^261 = 64 * 19
  static constexpr u32 T0 = 19u, T1 = 0u;            // 2^255 = 19
};

constexpr int FQ_L = 9, FQ_B = 29;
constexpr u32 FQ_MASK = (1u << FQ_B) - 1;

template <class P> EB_HD u32 fq_p_limb(int i) {
  return i == 0 ? (1u << FQ_B) - P::T0 : i == 1 ? FQ_MASK - P::T1 : i < 8 ? FQ_MASK : (1u << P::TOPBITS) - 1;
}

template <class P, int M>
struct fq {
  static_assert(M >= 1 && M <= 7, "lazy limbs must stay below 2^32");
  u32 v[FQ_L];
  fq() = default;
  template <int M2>
  EB_HD fq(const fq<P, M2>& o) {
    static_assert(M2 <= M, "magnitude can only be widened implicitly");
#pragma unroll
    for (int i = 0; i < FQ_L; i++) v[i] = o.v[i];
  }
};

// acc += a * b as ONE IMAD.WIDE.U32.  Written as PTX on the device: left to the front end, the column sums
// are re-associated into independent partial sums joined by 64-bit carry-chain adds (IADD3 + IADD3.X, the
// very instructions this representation exists to avoid), and `u * 256` is strength-reduced into a
// 6-instruction shift/add sequence.
EB_HD void fq_mac(u64& acc, u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
#else
  acc += (u64)a * b;
#endif
}

// product MAC of the column sums: FQ_MAC_MODE 1 = PTX (ptxas then picks between chained IMAD.WIDE and
// IMAD.WIDE + 3-input IADD3 trees), 0 = C expression (the front end re-associates)
#ifndef FQ_MAC_MODE
#define FQ_MAC_MODE 1
#endif
EB_HD void fq_pmac(u64& acc, u32 a, u32 b) {
#if FQ_MAC_MODE
  fq_mac(acc, a, b);
#else
  acc += (u64)a * b;
#endif
}

// ---------------------------------------------------------------------------------------------
// Product / square cores.  Two interleaved accumulators: d walks columns 8..16 (its 29-bit
// digits u are folded down by 2^261 = R0 + R1 * 2^29), c walks columns 0..7 and collects the folds.
template <class P>
EB_HD void fq_finish(u32* r, u64 c, u64 d, u32 t8, const u32* t) {
  // c: carry into limb 8; d: what is left above column 16 (weight 2^(29*17) = 2^(29*8) * 2^261).
  // d < 2^32 here because the top limbs are short (<= M * 2^TOPBITS, the fq contract).
  const u32 dl = (u32)d;
  fq_mac(c, t8, 1u);
  fq_mac(c, dl, P::R0);
  r[8] = (u32)c & ((1u << P::TOPBITS) - 1);
  c >>= P::TOPBITS;
  if (P::R1) fq_mac(c, dl, P::R1 << (FQ_B - P::TOPBITS));
  // c * 2^(232 + TOPBITS) = c * (T0 + T1 * 2^29); c < 2^46
  const u32 cl = (u32)c, ch = (u32)(c >> 32);
  u64 e = ((u64)(ch * P::T0) << 32) | t[0];
  fq_mac(e, cl, P::T0);
  r[0] = (u32)e & FQ_MASK; e >>= FQ_B;
  if (P::T1) e += c * P::T1;
  fq_mac(e, t[1], 1u);
  r[1] = (u32)e & FQ_MASK; e >>= FQ_B;
  u32 x = (u32)e + t[2];                        // e < 2^24 here
  r[2] = x & FQ_MASK;
  r[3] = t[3] + (x >> FQ_B);                    // <= 2^29 + 1: stays lazy
  r[4] = t[4]; r[5] = t[5]; r[6] = t[6]; r[7] = t[7];
}

template <class P>
EB_HD void fq_mul_core(u32* r, const u32* a, const u32* b) {
  u64 c = 0, d = 0;
  u32 t[8];
#pragma unroll
  for (int i = 0; i < 9; i++) fq_pmac(d, a[i], b[8 - i]);
  u32 t8 = (u32)d & FQ_MASK; d >>= FQ_B;
#pragma unroll
  for (int k = 0; k < 8; k++) {
#pragma unroll
    for (int i = k + 1; i < 9; i++) fq_pmac(d, a[i], b[9 + k - i]);
    u32 u = (u32)d & FQ_MASK; d >>= FQ_B;
    fq_mac(c, u, P::R0);          // first in the chain, so that it stays one IMAD.WIDE on the accumulator
#pragma unroll
    for (int i = 0; i <= k; i++) fq_pmac(c, a[i], b[k - i]);
    t[k] = (u32)c & FQ_MASK; c >>= FQ_B;
    if (P::R1) fq_mac(c, u, P::R1);
  }
  fq_finish<P>(r, c, d, t8, t);
}

template <class P>
EB_HD void fq_sqr_core(u32* r, const u32* a) {
  u32 a2[9];
#pragma unroll
  for (int i = 0; i < 9; i++) a2[i] = a[i] << 1;
  u64 c = 0, d = 0;
  u32 t[8];
  // column 8: 2 (a0 a8 + a1 a7 + a2 a6 + a3 a5) + a4^2
#pragma unroll
  for (int i = 0; i < 4; i++) fq_pmac(d, a2[i], a[8 - i]);
  fq_pmac(d, a[4], a[4]);
  u32 t8 = (u32)d & FQ_MASK; d >>= FQ_B;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    // high column 9 + k: pairs (i, 9 + k - i), k + 1 <= i < 9 + k - i
    const int s = 9 + k;
#pragma unroll
    for (int i = k + 1; 2 * i < s; i++) fq_pmac(d, a2[i], a[s - i]);
    if ((s & 1) == 0) fq_pmac(d, a[s / 2], a[s / 2]);
    u32 u = (u32)d & FQ_MASK; d >>= FQ_B;
    fq_mac(c, u, P::R0);
    // low column k: pairs (i, k - i), i < k - i
#pragma unroll
    for (int i = 0; 2 * i < k; i++) fq_pmac(c, a2[i], a[k - i]);
    if ((k & 1) == 0) fq_pmac(c, a[k / 2], a[k / 2]);
    t[k] = (u32)c & FQ_MASK; c >>= FQ_B;
    if (P::R1) fq_mac(c, u, P::R1);
  }
  fq_finish<P>(r, c, d, t8, t);
}

template <class P, int MA, int MB>
EB_HD fq<P, 1> fq_mul(const fq<P, MA>& a, const fq<P, MB>& b) {
  static_assert(MA * MB <= 7, "column sums would overflow 64 bits: weak() an operand first");
  fq<P, 1> r;
  fq_mul_core<P>(r.v, a.v, b.v);
  return r;
}
template <class P, int MA>
EB_HD fq<P, 1> fq_sqr(const fq<P, MA>& a) {
  static_assert(MA * MA <= 7, "column sums would overflow 64 bits: weak() the operand first");
  fq<P, 1> r;
  fq_sqr_core<P>(r.v, a.v);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Lazy linear operations
template <class P, int MA, int MB>
EB_HD fq<P, MA + MB> fq_add(const fq<P, MA>& a, const fq<P, MB>& b) {
  fq<P, MA + MB> r;
#pragma unroll
  for (int i = 0; i < FQ_L; i++) r.v[i] = a.v[i] + b.v[i];
  return r;
}
// a - b + (MB + 1) p : every limb of (MB + 1) p covers the corresponding limb of b
template <class P, int MA, int MB>
EB_HD fq<P, MA + MB + 1> fq_sub(const fq<P, MA>& a, const fq<P, MB>& b) {
  fq<P, MA + MB + 1> r;
#pragma unroll
  for (int i = 0; i < FQ_L; i++) r.v[i] = a.v[i] + (u32)(MB + 1) * fq_p_limb<P>(i) - b.v[i];
  return r;
}
template <class P, int MA>
EB_HD fq<P, MA + 1> fq_neg(const fq<P, MA>& a) {
  fq<P, MA + 1> r;
#pragma unroll
  for (int i = 0; i < FQ_L; i++) r.v[i] = (u32)(MA + 1) * fq_p_limb<P>(i) - a.v[i];
  return r;
}
// neg ? -a : a, branch-free; the result is typed with the larger magnitude
template <class P, int MA>
EB_HD fq<P, MA + 1> fq_cneg(const fq<P, MA>& a, bool neg) {
  fq<P, MA + 1> r;
  u32 s = 0u - (u32)neg;
#pragma unroll
  for (int i = 0; i < FQ_L; i++) r.v[i] = (a.v[i] ^ s) - s + (s & ((u32)(MA + 1) * fq_p_limb<P>(i)));
  return r;
}
template <int K, class P, int MA>
EB_HD fq<P, K * MA> fq_mul_int(const fq<P, MA>& a) {
  fq<P, K * MA> r;
#pragma unroll
  for (int i = 0; i < FQ_L; i++) r.v[i] = a.v[i] * (u32)K;
  return r;
}

// carry pass: any magnitude -> 1 (limbs 0..7 < 2^29, top limb <= 2^TOPBITS + 7); the value is unchanged mod p
template <class P, int MA>
EB_HD fq<P, 1> fq_weak(const fq<P, MA>& a) {
  fq<P, 1> r;
  u32 t = a.v[8] >> P::TOPBITS;
  u32 v8 = a.v[8] & ((1u << P::TOPBITS) - 1);
  u32 x = a.v[0] + t * P::T0;
  r.v[0] = x & FQ_MASK;
  x = (x >> FQ_B) + a.v[1] + t * P::T1;
#pragma unroll
  for (int i = 1; i < 8; i++) {
    r.v[i] = x & FQ_MASK;
    x = (x >> FQ_B) + (i < 7 ? a.v[i + 1] : v8);
  }
  r.v[8] = x;
  return r;
}

// canonical residue in [0, p)  (what bn.js fromRed() exposes)
template <class P, int MA>
EB_HD fq<P, 1> fq_canon(const fq<P, MA>& a) {
  fq<P, 1> r = fq_weak(a);                      // value < 2^(232+TOPBITS) + 8 * 2^232
  const u32 TM = (1u << P::TOPBITS) - 1;
  // x = 1 iff the value is >= p
  u32 m = r.v[2] & r.v[3] & r.v[4] & r.v[5] & r.v[6] & r.v[7];
  u32 x = (r.v[8] >> P::TOPBITS) |
          ((r.v[8] == TM) & (m == FQ_MASK) &
           (((u64)r.v[1] << FQ_B | r.v[0]) >= ((u64)fq_p_limb<P>(1) << FQ_B | fq_p_limb<P>(0))));
  // add x * (2^(232+TOPBITS) - p) and drop bit 232+TOPBITS
  u32 c = r.v[0] + x * P::T0;
  r.v[0] = c & FQ_MASK;
  c = (c >> FQ_B) + r.v[1] + x * P::T1;
#pragma unroll
  for (int i = 1; i < 8; i++) {
    r.v[i] = c & FQ_MASK;
    c = (c >> FQ_B) + r.v[i + 1];
  }
  r.v[8] = c & TM;
  return r;
}

// a == 0 (mod p)
template <class P, int MA>
EB_HD bool fq_is_zero(const fq<P, MA>& a) {
  fq<P, 1> r = fq_weak(a);                      // < 2 p : zero iff 0 or p
  u32 o = 0, n = FQ_MASK;
#pragma unroll
  for (int i = 0; i < 8; i++) { o |= r.v[i]; n &= r.v[i] ^ (FQ_MASK ^ fq_p_limb<P>(i)); }
  o |= r.v[8];
  return o == 0 || (n == FQ_MASK && r.v[8] == fq_p_limb<P>(8));
}
template <class P, int MA, int MB>
EB_HD bool fq_eq(const fq<P, MA>& a, const fq<P, MB>& b) { return fq_is_zero(fq_sub(a, b)); }

// ---------------------------------------------------------------------------------------------
// packed 8 x 32-bit words (value < 2^256) <-> 9 x 29-bit limbs
template <class P>
EB_HD fq<P, 1> fq_from_words(const u32* w) {
  fq<P, 1> r;
  r.v[0] = w[0] & FQ_MASK;
#pragma unroll
  for (int i = 1; i < 8; i++) {
    const int bit = FQ_B * i, k = bit >> 5, s = bit & 31;     // limb i = bits [29 i, 29 i + 29)
    u32 lo = w[k] >> s;
    u32 hi = (s > 3 && k + 1 < 8) ? (w[k + 1] << (32 - s)) : 0u;
    r.v[i] = (lo | hi) & FQ_MASK;
  }
  r.v[8] = w[7] >> 8;                                        // bits 232..255
  if (P::TOPBITS < 24) {                                     // 25519: bit 255 is worth 19
    u32 t = r.v[8] >> P::TOPBITS;
    r.v[8] &= (1u << P::TOPBITS) - 1;
    r.v[0] += t * P::T0;                                     // <= 2^29 + 18: still magnitude 1
  }
  return r;
}
// canonical limbs -> packed words
template <class P>
EB_HD void fq_to_words(u32* w, const fq<P, 1>& c) {
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int bit = 32 * k, i = bit / FQ_B, s = bit % FQ_B;   // word k = bits [32 k, 32 k + 32)
    u32 lo = c.v[i] >> s;
    u32 hi = (i + 1 < 9) ? (c.v[i + 1] << (FQ_B - s)) : 0u;
    u32 hi2 = (FQ_B - s + FQ_B < 32 && i + 2 < 9) ? (c.v[i + 2] << (2 * FQ_B - s)) : 0u;
    w[k] = lo | hi | hi2;
  }
}

}  // namespace eb
