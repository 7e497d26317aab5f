// limbs.cuh -- multi-limb (32-bit) integer primitives for sm_100a.
//
// Replaces the role of bn.js's comb10MulTo / smallMulTo / iadd / isub
// (reference dist/elliptic.js:4941-5557, 4505-4607) with register-resident
// 32-bit-limb arithmetic: the device path is inline PTX carry chains
// (mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32[.X]
// with predicate carries -- one fma-pipe instruction per 32x32+64 MAC).
//
// When this header is compiled for the host (no __CUDA_ARCH__) every
// primitive has a portable C++ body.  That body exists ONLY so the unit
// tests can run the kernel logic on a CPU-only box (tests/_hostemu); the
// product library never executes it -- see capi.cu, which fails loudly when
// no CUDA device is present.
#pragma once
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;

#if defined(__CUDACC__)
#define EB_HD __host__ __device__ __forceinline__
#define EB_D __device__ __forceinline__
#else
#define EB_HD inline
#define EB_D inline
#endif

namespace eb {

// ---------------------------------------------------------------------------
// r = a + b; returns carry out.
template <int N>
EB_HD u32 add_n(u32* r, const u32* a, const u32* b) {
#if defined(__CUDA_ARCH__)
  u32 cout;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r[0]) : "r"(a[0]), "r"(b[0]));
#pragma unroll
  for (int i = 1; i < N; i++)
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(a[i]), "r"(b[i]));
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(cout));
  return cout;
#else
  u64 c = 0;
  for (int i = 0; i < N; i++) {
    c += (u64)a[i] + b[i];
    r[i] = (u32)c;
    c >>= 32;
  }
  return (u32)c;
#endif
}

// r = a - b; returns borrow out (0/1).
template <int N>
EB_HD u32 sub_n(u32* r, const u32* a, const u32* b) {
#if defined(__CUDA_ARCH__)
  u32 bout;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r[0]) : "r"(a[0]), "r"(b[0]));
#pragma unroll
  for (int i = 1; i < N; i++)
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(a[i]), "r"(b[i]));
  asm volatile("subc.u32 %0, 0, 0;" : "=r"(bout));
  return bout & 1;
#else
  u64 c = 0;
  for (int i = 0; i < N; i++) {
    u64 t = (u64)a[i] - b[i] - c;
    r[i] = (u32)t;
    c = (t >> 32) & 1;
  }
  return (u32)c;
#endif
}

// ---------------------------------------------------------------------------
// Wide multiply r[2N] = a[N] * b[N].
// Device: even/odd column split so that every (lo,hi) product pair lands on a
// register pair and each row is a single carry chain; ptxas turns each
// mad.lo.cc/madc.hi.cc pair into one IMAD.WIDE.U32.X.
#if defined(__CUDA_ARCH__)
template <int n>
EB_D void mul_row0(u32* acc, const u32* a, u32 bi) {  // acc[j],acc[j+1] = a[j]*bi, j even < n
#pragma unroll
  for (int j = 0; j < n; j += 2)
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;"
                 : "=r"(acc[j]), "=r"(acc[j + 1]) : "r"(a[j]), "r"(bi));
}
template <int n>
EB_D void cmad_row(u32* acc, const u32* a, u32 bi) {  // (acc[j],acc[j+1]) += a[j]*bi, carry chained, CC live-out
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(bi));
#pragma unroll
  for (int j = 2; j < n; j += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                 : "+r"(acc[j]), "+r"(acc[j + 1]) : "r"(a[j]), "r"(bi));
}
template <int n>
EB_D void mad_row(u32* odd, u32* even, const u32* a, u32 bi) {
  cmad_row<n - 2>(odd, a + 1, bi);
  asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;"
               : "=r"(odd[n - 2]), "=r"(odd[n - 1]) : "r"(a[n - 1]), "r"(bi));
  cmad_row<n>(even, a, bi);
  asm volatile("addc.u32 %0, %0, 0;" : "+r"(odd[n - 1]));
}
#endif

template <int N>
EB_HD void mul_wide(u32* r, const u32* a, const u32* b) {
#if defined(__CUDA_ARCH__)
  static_assert(N % 2 == 0, "even limb count");
  u32 even[2 * N], odd[2 * N];
  mul_row0<N>(even, a, b[0]);
  mul_row0<N>(odd, a + 1, b[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i += 2) {
    mad_row<N>(&even[i + 1], &odd[i - 1], a, b[i]);
    mad_row<N>(&odd[i + 1], &even[i + 1], a, b[i + 1]);
  }
  mad_row<N>(&even[N], &odd[N - 2], a, b[N - 1]);
  r[0] = even[0];
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r[1]) : "r"(even[1]), "r"(odd[0]));
#pragma unroll
  for (int i = 2; i < 2 * N - 1; i++)
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(even[i]), "r"(odd[i - 1]));
  asm volatile("addc.u32 %0, %1, 0;" : "=r"(r[2 * N - 1]) : "r"(even[2 * N - 1]));
#else
  for (int i = 0; i < 2 * N; i++) r[i] = 0;
  for (int i = 0; i < N; i++) {
    u64 c = 0;
    for (int j = 0; j < N; j++) {
      c += (u64)a[j] * b[i] + r[i + j];
      r[i + j] = (u32)c;
      c >>= 32;
    }
    r[i + N] = (u32)c;
  }
#endif
}

// Wide square r[2N] = a[N]^2.  (First version: plain product; a dedicated
// N(N+1)/2-MAC squaring replaces it in fe_k256.cuh once measured.)
template <int N>
EB_HD void sqr_wide(u32* r, const u32* a) {
  mul_wide<N>(r, a, a);
}

// Generic (portable, compiler-scheduled) rectangular multiply, used by the
// scalar-field / GLV code where throughput is not critical:
// r[NA+NB] = a[NA] * b[NB].
template <int NA, int NB>
EB_HD void mul_rect(u32* r, const u32* a, const u32* b) {
#pragma unroll
  for (int i = 0; i < NA + NB; i++) r[i] = 0;
#pragma unroll
  for (int i = 0; i < NB; i++) {
    u64 c = 0;
#pragma unroll
    for (int j = 0; j < NA; j++) {
      c += (u64)a[j] * b[i] + r[i + j];
      r[i + j] = (u32)c;
      c >>= 32;
    }
    r[i + NA] = (u32)c;
  }
}

template <int N>
EB_HD bool is_zero_n(const u32* a) {
  u32 t = 0;
#pragma unroll
  for (int i = 0; i < N; i++) t |= a[i];
  return t == 0;
}

template <int N>
EB_HD bool eq_n(const u32* a, const u32* b) {
  u32 t = 0;
#pragma unroll
  for (int i = 0; i < N; i++) t |= a[i] ^ b[i];
  return t == 0;
}

// a >= b ?
template <int N>
EB_HD bool geq_n(const u32* a, const u32* b) {
  u32 t[N];
  return sub_n<N>(t, a, b) == 0;
}

template <int N>
EB_HD void copy_n(u32* r, const u32* a) {
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = a[i];
}

// r = cond ? a : r   (branch-free select)
template <int N>
EB_HD void cmov_n(u32* r, const u32* a, bool cond) {
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = cond ? a[i] : r[i];
}

// Load N limbs from a big-endian byte string of 4N bytes (wire format of the
// reference's toArray('be', len), base.js:298-306).
template <int N>
EB_HD void load_be(u32* r, const uint8_t* p) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint8_t* q = p + 4 * (N - 1 - i);
    r[i] = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
  }
}
template <int N>
EB_HD void store_be(uint8_t* p, const u32* a) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint8_t* q = p + 4 * (N - 1 - i);
    q[0] = (uint8_t)(a[i] >> 24); q[1] = (uint8_t)(a[i] >> 16);
    q[2] = (uint8_t)(a[i] >> 8);  q[3] = (uint8_t)a[i];
  }
}

// Big-endian byte strings whose length is not a multiple of 4 (p521: 66 bytes in 18 limbs).
template <int N>
EB_HD void load_be_len(u32* r, const uint8_t* p, int len) {
  for (int w = 0; w < N; w++) r[w] = 0;
  for (int k = 0; k < len; k++) r[k >> 2] |= (u32)p[len - 1 - k] << (8 * (k & 3));
}
template <int N>
EB_HD void store_be_len(uint8_t* p, const u32* a, int len) {
  for (int k = 0; k < len; k++) p[len - 1 - k] = (uint8_t)(a[k >> 2] >> (8 * (k & 3)));
}

template <int N>
EB_HD void load_le(u32* r, const uint8_t* p) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint8_t* q = p + 4 * i;
    r[i] = ((u32)q[3] << 24) | ((u32)q[2] << 16) | ((u32)q[1] << 8) | (u32)q[0];
  }
}

}  // namespace eb
