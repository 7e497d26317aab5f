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

#if defined(__CUDA_ARCH__)
  // CIOS with two accumulators so that every (lo,hi) product pair stays on a fixed register pair
  // (ptxas then emits one IMAD.WIDE.U32.X per 32x32+64 MAC and no realignment moves):
  //   T = E + O * 2^32,  E takes the even-indexed limbs' products, O the odd-indexed ones.
  // After a row, E[0] == 0; dividing by 2^32 swaps the roles: the old O becomes the new E (absorbing
  // E[1], whose carry is exactly what the new O's lowest limb must receive), and the old E shifted
  // down by two limbs becomes the new O -- the shift is done for free by the 3-address mad.
  // Arrays carry two extra limbs because p256 / p384 / the group orders use every bit of their top limb.
  // One PTX instruction per asm statement: with n0inv == 1 the multiplier m IS E[0], and a tied "+r"
  // operand inside a multi-instruction asm would be overwritten before its second use.
  template <bool FIRST>
  static EB_D void cios_row(u32* E, u32* X, const u32* a, u32 bi, const u32* p) {
    // X: previous even array on entry; rewritten in place as the new odd array.
    if (!FIRST) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(E[0]) : "r"(X[1]));
#pragma unroll
    for (int j = 1; j < N; j += 2) {
      if (FIRST) {
        if (j == 1) asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(X[0]) : "r"(a[1]), "r"(bi));
        else asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(X[j - 1]) : "r"(a[j]), "r"(bi));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(X[j]) : "r"(a[j]), "r"(bi));
      } else {
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(X[j - 1]) : "r"(a[j]), "r"(bi), "r"(X[j + 1]));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(X[j]) : "r"(a[j]), "r"(bi), "r"(X[j + 2]));
      }
    }
    if (FIRST) { X[N] = 0; X[N + 1] = 0; }
    else {
      asm volatile("addc.cc.u32 %0, 0, 0;" : "=r"(X[N]));
      asm volatile("addc.u32 %0, 0, 0;" : "=r"(X[N + 1]));
    }
    // E += even-indexed limbs of a times bi
    if (FIRST) {
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(E[j]) : "r"(a[j]), "r"(bi));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(E[j + 1]) : "r"(a[j]), "r"(bi));
      }
      E[N] = 0; E[N + 1] = 0;
    } else {
      asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(E[0]) : "r"(a[0]), "r"(bi));
      asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(E[1]) : "r"(a[0]), "r"(bi));
#pragma unroll
      for (int j = 2; j < N; j += 2) {
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(E[j]) : "r"(a[j]), "r"(bi));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(E[j + 1]) : "r"(a[j]), "r"(bi));
      }
      asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(E[N]));
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(E[N + 1]));
    }
    u32 m = E[0] * P::n0inv;
    // O (= X) += odd-indexed limbs of p times m
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(X[0]) : "r"(p[1]), "r"(m));
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(X[1]) : "r"(p[1]), "r"(m));
#pragma unroll
    for (int j = 3; j < N; j += 2) {
      asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(X[j - 1]) : "r"(p[j]), "r"(m));
      asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(X[j]) : "r"(p[j]), "r"(m));
    }
    asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(X[N]));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(X[N + 1]));
    // E += even-indexed limbs of p times m   (E[0] becomes 0)
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(E[0]) : "r"(p[0]), "r"(m));
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(E[1]) : "r"(p[0]), "r"(m));
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(E[j]) : "r"(p[j]), "r"(m));
      asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(E[j + 1]) : "r"(p[j]), "r"(m));
    }
    asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(E[N]));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(E[N + 1]));
  }
  static EB_D fe mul_ptx(const fe& a, const fe& b) {
    static_assert(N % 2 == 0, "PTX CIOS path is written for an even limb count");
    u32 p[N]; P::mod(p);
    u32 A[N + 3], B[N + 3];
    A[N + 2] = 0; B[N + 2] = 0;
    // row 0: E = A, new odd = B.  Afterwards roles alternate: the new even array is the old odd one.
    cios_row<true>(A, B, a.v, b.v[0], p);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
      cios_row<false>(B, A, a.v, b.v[i], p);                 // even = B (old odd), A: old even -> new odd
      if (i + 1 < N) cios_row<false>(A, B, a.v, b.v[i + 1], p);
    }
    // N even: the last row had even = B, odd = A.  T / 2^32 = (B >> 32) + A
    u32 t[N + 1];
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(t[0]) : "r"(B[1]), "r"(A[0]));
#pragma unroll
    for (int k = 1; k < N; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(t[k]) : "r"(B[k + 1]), "r"(A[k]));
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(t[N]) : "r"(B[N + 1]), "r"(A[N]));
    fe r, d;
    u32 bw = sub_n<N>(d.v, t, p);
    bool ge = t[N] != 0 || bw == 0;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = ge ? d.v[i] : t[i];
    return r;
  }
#endif

  // a*b*R^-1 mod p.  Requires a < R, b < p (or a < p, b < R); result in [0, p).
  // Out-of-line on the device (operands in registers): one copy of the ~250-instruction multiplier
  // per field instead of one per use keeps the double/add loop inside the instruction cache.
#if defined(__CUDACC__)
  static __device__ __noinline__ fe mul_ol(fe a, fe b) { return mul_ptx_or_c(a, b); }
#endif
  static EB_HD fe mul(const fe& a, const fe& b) {
#if defined(__CUDA_ARCH__) && !defined(EB_MONT_INLINE)
    return mul_ol(a, b);
#else
    return mul_ptx_or_c(a, b);
#endif
  }
  static EB_HD fe mul_ptx_or_c(const fe& a, const fe& b) {
#if defined(__CUDA_ARCH__) && !defined(EB_MONT_PORTABLE)
    if (N % 2 == 0) return mul_ptx(a, b);
#endif
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
  // K a b and K a^2 for the small constants of the doubling formulas (K = 3, 4, 8): here by modular doublings of the
  // product; FpS (fp_special.cuh) folds K into the reduction instead
  template <int K> static EB_HD fe scale_k(const fe& r) {
    static_assert(K == 1 || K == 3 || K == 4 || K == 8, "scale");
    if (K == 3) return add(add(r, r), r);
    if (K == 4) { fe t = add(r, r); return add(t, t); }
    if (K == 8) { fe t = add(r, r); t = add(t, t); return add(t, t); }
    return r;
  }
  static EB_HD fe canon(const fe& a) { return a; }          // elements are always canonical here (FpS: see fp_special.cuh)
  template <int K> static EB_HD fe mul_k(const fe& a, const fe& b) { return scale_k<K>(mul(a, b)); }
  template <int K> static EB_HD fe sqr_k(const fe& a) { return scale_k<K>(sqr(a)); }

#if defined(__CUDA_ARCH__)
  // Device add / sub for moduli whose top limb is 0xFFFFFFFF (p256, p384 and their group orders): a carry out of
  // a + b is undone by adding R - p under a mask, and a sum in [p, R) can only occur when its top limb is all ones
  // (2^-32 of the time), which is left to a cold branch; a borrow out of a - b is undone by adding p under a mask.
  // No compare-and-select chain on the common path.
  static EB_D fe add_fast(const fe& a, const fe& b) {
    u32 p[N], rp[N];
    P::mod(p); P::r1(rp);                                  // R - p == R mod p because p > R / 2
    fe r;
    const u32 Z = 0;
    u32 cy;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.v[0]) : "r"(a.v[0]), "r"(b.v[0]));
#pragma unroll
    for (int i = 1; i < N; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.v[i]) : "r"(a.v[i]), "r"(b.v[i]));
    asm volatile("addc.u32 %0, %1, %1;" : "=r"(cy) : "r"(Z));
    u32 mask = 0u - cy;
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r.v[0]) : "r"(rp[0] & mask));
#pragma unroll
    for (int i = 1; i < N - 1; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r.v[i]) : "r"(rp[i] & mask));
    asm volatile("addc.u32 %0, %0, %1;" : "+r"(r.v[N - 1]) : "r"(rp[N - 1] & mask));
    if (!cy && r.v[N - 1] == 0xffffffffu) {                // cold: the sum may lie in [p, R)
      fe d;
      u32 bw = sub_n<N>(d.v, r.v, p);
      if (!bw) r = d;
    }
    return r;
  }
  static EB_D fe sub_fast(const fe& a, const fe& b) {
    u32 p[N];
    P::mod(p);
    fe r;
    const u32 Z = 0;
    u32 bw;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r.v[0]) : "r"(a.v[0]), "r"(b.v[0]));
#pragma unroll
    for (int i = 1; i < N; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r.v[i]) : "r"(a.v[i]), "r"(b.v[i]));
    asm volatile("subc.u32 %0, %1, %1;" : "=r"(bw) : "r"(Z));        // 0 or 0xFFFFFFFF
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r.v[0]) : "r"(p[0] & bw));
#pragma unroll
    for (int i = 1; i < N - 1; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r.v[i]) : "r"(p[i] & bw));
    asm volatile("addc.u32 %0, %0, %1;" : "+r"(r.v[N - 1]) : "r"(p[N - 1] & bw));
    return r;
  }
  static EB_D bool top_limb_all_ones() { u32 p[N]; P::mod(p); return p[N - 1] == 0xffffffffu; }
#endif

  static EB_HD fe add(const fe& a, const fe& b) {
#if defined(__CUDA_ARCH__) && defined(EB_MONT_FAST_ADDSUB) && EB_MONT_FAST_ADDSUB   // opt-in until re-verified on the GPU
    if (top_limb_all_ones()) return add_fast(a, b);
#endif
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
#if defined(__CUDA_ARCH__) && defined(EB_MONT_FAST_ADDSUB) && EB_MONT_FAST_ADDSUB
    if (top_limb_all_ones()) return sub_fast(a, b);
#endif
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
    u32 e[N], two[N];
    P::mod(e);
    for (int i = 0; i < N; i++) two[i] = i == 0 ? 2u : 0u;
    sub_n<N>(e, e, two);    // full borrow chain: p224's lowest limb is 1
    return pow(a, e);
  }
};

template <int N>
EB_HD fe_n<N> load_fe_n(const u32* src) { fe_n<N> a; for (int i = 0; i < N; i++) a.v[i] = src[i]; return a; }
template <int N>
EB_HD void store_fe_n(u32* dst, const fe_n<N>& a) { for (int i = 0; i < N; i++) dst[i] = a.v[i]; }

}  // namespace eb
