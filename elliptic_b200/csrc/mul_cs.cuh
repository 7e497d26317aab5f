// mul_cs.cuh -- "carry-save" wide products for the packed 32-bit-limb fields (device only).
#pragma once
#include "limbs.cuh"
namespace eb {
#if defined(__CUDACC__)
// carry-save products: every 32x32 MAC is a plain IMAD.WIDE.U32 with a carry-OUT only (half the issue cost of the
// carry-in form IMAD.WIDE.U32.X on sm_100); the carries are counted on the ALU pipe and enter the accumulator two
// columns up as the addend of its first product.
template <int N>
EB_D void mul_wide_cs(u32* r, const u32* a, const u32* b) {
  u32 lo[2 * N - 1], hi[2 * N - 1], cn[2 * N - 1];
#pragma unroll
  for (int s = 0; s < 2 * N - 1; s++) {
    const int i0 = s < N ? 0 : s - N + 1, i1 = s < N ? s : N - 1;
    u64 t = (u64)a[i0] * b[s - i0] + (s >= 2 ? cn[s - 2] : 0u);
    lo[s] = (u32)t; hi[s] = (u32)(t >> 32); cn[s] = 0;
#pragma unroll
    for (int i = i0 + 1; i <= i1; i++)
      asm("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
          : "+r"(lo[s]), "+r"(hi[s]), "+r"(cn[s]) : "r"(a[i]), "r"(b[s - i]));
  }
  r[0] = lo[0];
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r[1]) : "r"(lo[1]), "r"(hi[0]));
#pragma unroll
  for (int t = 2; t < 2 * N - 1; t++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[t]) : "r"(lo[t]), "r"(hi[t - 1]));
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(r[2 * N - 1]) : "r"(hi[2 * N - 2]), "r"(cn[2 * N - 3]));
}

template <int N>
EB_D void sqr_wide_cs(u32* r, const u32* a) {
  // T = sum_{i<j} a_i a_j 2^(32(i+j)) on accumulators s = i + j = 1 .. 2N-3
  u32 lo[2 * N - 2], hi[2 * N - 2], cn[2 * N - 2];
#pragma unroll
  for (int s = 1; s <= 2 * N - 3; s++) {
    const int i0 = s < N ? 0 : s - N + 1, i1 = (s - 1) / 2;
    u64 t = (u64)a[i0] * a[s - i0] + (s >= 3 ? cn[s - 2] : 0u);
    lo[s] = (u32)t; hi[s] = (u32)(t >> 32); cn[s] = 0;
#pragma unroll
    for (int i = i0 + 1; i <= i1; i++)
      asm("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
          : "+r"(lo[s]), "+r"(hi[s]), "+r"(cn[s]) : "r"(a[i]), "r"(a[s - i]));
  }
  u32 T[2 * N];
  T[0] = 0; T[1] = lo[1];
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(T[2]) : "r"(lo[2]), "r"(hi[1]));
#pragma unroll
  for (int t = 3; t <= 2 * N - 3; t++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(T[t]) : "r"(lo[t]), "r"(hi[t - 1]));
  asm volatile("addc.u32 %0, %1, 0;" : "=r"(T[2 * N - 2]) : "r"(hi[2 * N - 3]));
  // U = 2 T
  u32 U[2 * N];
  U[0] = 0;
  asm volatile("add.cc.u32 %0, %1, %1;" : "=r"(U[1]) : "r"(T[1]));
#pragma unroll
  for (int t = 2; t <= 2 * N - 2; t++) asm volatile("addc.cc.u32 %0, %1, %1;" : "=r"(U[t]) : "r"(T[t]));
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(U[2 * N - 1]));
  // r = U + D
  u64 d0 = (u64)a[0] * a[0];
  r[0] = (u32)d0;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r[1]) : "r"(U[1]), "r"((u32)(d0 >> 32)));
#pragma unroll
  for (int i = 1; i < N; i++) {
    u64 d = (u64)a[i] * a[i];
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[2 * i]) : "r"(U[2 * i]), "r"((u32)d));
    if (i < N - 1) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[2 * i + 1]) : "r"(U[2 * i + 1]), "r"((u32)(d >> 32)));
    else asm volatile("addc.u32 %0, %1, %2;" : "=r"(r[2 * i + 1]) : "r"(U[2 * i + 1]), "r"((u32)(d >> 32)));
  }
}
#endif
}  // namespace eb
