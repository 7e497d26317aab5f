// imad_peak.cu -- measures the integer-multiply roofline denominators on the GPU:
// IMAD / IMAD.WIDE.U32(.X) issue rates, IADD3, their co-issue, and the field
// multiplication built from them.  Output: one JSON object on stdout.
// (MEASURED_PEAKS.json has no integer peak, so it is measured here.)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../elliptic_b200/csrc/ecdsa_k256_body.cuh"
using namespace eb;

#define ITERS 4096

// 8 independent 32x32+64 MAC chains with carry (the mul_wide inner pattern)
__global__ void k_wide_cc(u32* out, u32 seed) {
  u32 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x + threadIdx.x * 7;
  u32 x[16];
  for (int i = 0; i < 16; i++) x[i] = a * (i + 1);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[0]), "+r"(x[1]) : "r"(a), "r"(b));
      asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[2]), "+r"(x[3]) : "r"(a), "r"(b));
      asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[4]), "+r"(x[5]) : "r"(a), "r"(b));
      asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(x[6]), "+r"(x[7]) : "r"(a), "r"(b));
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[8]), "+r"(x[9]) : "r"(b), "r"(a));
      asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[10]), "+r"(x[11]) : "r"(b), "r"(a));
      asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[12]), "+r"(x[13]) : "r"(b), "r"(a));
      asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(x[14]), "+r"(x[15]) : "r"(b), "r"(a));
    }
  }
  u32 s = 0;
  for (int i = 0; i < 16; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 32 wide MACs / iteration

// independent mad.wide.u32 (no carries)
__global__ void k_wide(u32* out, u32 seed) {
  u32 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x + threadIdx.x * 7;
  u64 x[8];
  for (int i = 0; i < 8; i++) x[i] = a * (i + 1);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a), "r"(b));
    }
  }
  u64 s = 0;
  for (int i = 0; i < 8; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (u32)s ^ (u32)(s >> 32);
}

// plain 32-bit IMAD
__global__ void k_imad(u32* out, u32 seed) {
  u32 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x + threadIdx.x * 7;
  u32 x[8];
  for (int i = 0; i < 8; i++) x[i] = a * (i + 1);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
  }
  u32 s = 0;
  for (int i = 0; i < 8; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// IADD3 carry chains
__global__ void k_iadd(u32* out, u32 seed) {
  u32 a = seed + threadIdx.x;
  u32 x[8];
  for (int i = 0; i < 8; i++) x[i] = a * (i + 1);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(a));
#pragma unroll
      for (int i = 1; i < 7; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
      asm volatile("addc.u32 %0, %0, %1;" : "+r"(x[7]) : "r"(a));
    }
  }
  u32 s = 0;
  for (int i = 0; i < 8; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 1:1 mix of wide MACs and adds (do the two pipes overlap?)
__global__ void k_mix(u32* out, u32 seed) {
  u32 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x + threadIdx.x * 7;
  u64 x[8];
  u32 y[8];
  for (int i = 0; i < 8; i++) { x[i] = a * (i + 1); y[i] = b + i; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a), "r"(b));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a));
      }
    }
  }
  u64 s = 0;
  for (int i = 0; i < 8; i++) s ^= x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (u32)s ^ (u32)(s >> 32);
}

template <int OP>
__global__ void k_fe(u32* out, u32 seed) {
  fe a, b;
  for (int i = 0; i < 8; i++) { a.v[i] = seed * (i + 1) + threadIdx.x; b.v[i] = seed * (i + 7) + blockIdx.x + 3 * threadIdx.x; }
  for (int it = 0; it < ITERS / 4; it++) {
    if (OP == 0) { a = fe_mul(a, b); b = fe_mul(b, a); }
    if (OP == 1) { a = fe_sqr(a); b = fe_sqr(b); }
    if (OP == 2) { a = fe_add(a, b); b = fe_sub(b, a); }
  }
  u32 s = 0;
  for (int i = 0; i < 8; i++) s ^= a.v[i] ^ b.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
__global__ void k_ge(u32* out, u32 seed) {
  ge_jac p;
  ge_aff q;
  for (int i = 0; i < 8; i++) {
    p.x.v[i] = seed * (i + 1) + threadIdx.x; p.y.v[i] = seed * (i + 7) + blockIdx.x; p.z.v[i] = seed + i;
    q.x.v[i] = seed * (i + 3); q.y.v[i] = seed * (i + 5) + threadIdx.x;
  }
  for (int it = 0; it < ITERS / 32; it++) {
    if (OP == 0) p = jac_dbl(p);
    if (OP == 1) p = jac_madd(p, q);
  }
  u32 s = 0;
  for (int i = 0; i < 8; i++) s ^= p.x.v[i] ^ p.y.v[i] ^ p.z.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- carry-free 9 x 29-bit field (fq_pm.cuh) and the group law on it (gq_k256.cuh)
template <int OP>
__global__ void k_fq(u32* out, u32 seed) {
  fqk<1> a, b;
  for (int i = 0; i < 9; i++) { a.v[i] = (seed * (i + 1) + threadIdx.x) & FQ_MASK; b.v[i] = (seed * (i + 7) + blockIdx.x + 3 * threadIdx.x) & FQ_MASK; }
  a.v[8] &= 0xFFFFFF; b.v[8] &= 0xFFFFFF;
  for (int it = 0; it < ITERS / 4; it++) {
    if (OP == 0) { a = fq_mul(a, b); b = fq_mul(b, a); }
    if (OP == 1) { a = fq_sqr(a); b = fq_sqr(b); }
    if (OP == 2) { a = fq_weak(fq_add(a, b)); b = fq_weak(fq_sub(b, a)); }
    if (OP == 3) { fqk<3> t = fq_sub(a, b); fqk<6> u = fq_sub(t, fq_add(a, b)); a = fq_weak(u); b = fq_weak(fq_add(b, a)); }
  }
  u32 s = 0;
  for (int i = 0; i < 9; i++) s ^= a.v[i] ^ b.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
__global__ void k_gq(u32* out, u32 seed) {
  gq_jac p;
  gq_aff q;
  for (int i = 0; i < 9; i++) {
    p.x.v[i] = (seed * (i + 1) + threadIdx.x) & FQ_MASK; p.y.v[i] = (seed * (i + 7) + blockIdx.x) & FQ_MASK; p.z.v[i] = (seed + i) & FQ_MASK;
    q.x.v[i] = (seed * (i + 3)) & FQ_MASK; q.y.v[i] = (seed * (i + 5) + threadIdx.x) & FQ_MASK;
  }
  for (int it = 0; it < ITERS / 32; it++) {
    if (OP == 0) p = gq_dbl(p);
    if (OP == 1) p = gq_madd(p, q);
    if (OP == 2) p = gq_dbl_inl(p);
    if (OP == 3) p = gq_madd_inl(p, q);
  }
  u32 s = 0;
  for (int i = 0; i < 9; i++) s ^= p.x.v[i] ^ p.y.v[i] ^ p.z.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 1:1 mix of wide MACs and funnel shifts (the carry extraction of a 29-bit-limb column sum)
__global__ void k_mix_shf(u32* out, u32 seed) {
  u32 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x + threadIdx.x * 7;
  u64 x[8];
  u32 y[8];
  for (int i = 0; i < 8; i++) { x[i] = a * (i + 1); y[i] = b + i; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a), "r"(b));
        asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(y[i]) : "r"(a));
      }
    }
  }
  u64 s = 0;
  for (int i = 0; i < 8; i++) s ^= x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (u32)s ^ (u32)(s >> 32);
}

template <typename K>
double run(K kern, int blocks, int threads, u32* d_out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(d_out, 12345u);
  cudaDeviceSynchronize();
  double best = 1e30;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    kern<<<blocks, threads>>>(d_out, 12345u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best * 1e-3;
}

int main(int argc, char** argv) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
  int sms = prop.multiProcessorCount;
  u32* d_out;
  cudaMalloc(&d_out, (size_t)sms * 64 * 1024 * 4);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);
  int cfgs[][2] = {{4, 128}, {4, 256}, {2, 512}, {8, 128}, {1, 1024}, {2, 1024}};
  for (auto& c : cfgs) {
    int blocks = sms * c[0], threads = c[1];
    double n = (double)blocks * threads;
    double t;
    t = run(k_wide_cc, blocks, threads, d_out);
    printf(",\n \"wide_cc_%dx%d_Tmac\": %.3f", c[0], c[1], n * ITERS * 32 / t / 1e12);
    t = run(k_wide, blocks, threads, d_out);
    printf(", \"wide_%dx%d_Tmac\": %.3f", c[0], c[1], n * ITERS * 32 / t / 1e12);
    t = run(k_imad, blocks, threads, d_out);
    printf(", \"imad_%dx%d_Tops\": %.3f", c[0], c[1], n * ITERS * 32 / t / 1e12);
    t = run(k_iadd, blocks, threads, d_out);
    printf(", \"iadd_%dx%d_Tops\": %.3f", c[0], c[1], n * ITERS * 32 / t / 1e12);
    t = run(k_mix, blocks, threads, d_out);
    printf(", \"mix_%dx%d_Tpairs\": %.3f", c[0], c[1], n * ITERS * 32 / t / 1e12);
    t = run(k_fe<0>, blocks, threads, d_out);
    printf(", \"fe_mul_%dx%d_G\": %.2f", c[0], c[1], n * (ITERS / 4) * 2 / t / 1e9);
    t = run(k_fe<1>, blocks, threads, d_out);
    printf(", \"fe_sqr_%dx%d_G\": %.2f", c[0], c[1], n * (ITERS / 4) * 2 / t / 1e9);
    t = run(k_fe<2>, blocks, threads, d_out);
    printf(", \"fe_addsub_%dx%d_G\": %.2f", c[0], c[1], n * (ITERS / 4) * 2 / t / 1e9);
    t = run(k_fq<0>, blocks, threads, d_out);
    printf(", \"fq_mul_%dx%d_G\": %.2f", c[0], c[1], n * (ITERS / 4) * 2 / t / 1e9);
    t = run(k_fq<1>, blocks, threads, d_out);
    printf(", \"fq_sqr_%dx%d_G\": %.2f", c[0], c[1], n * (ITERS / 4) * 2 / t / 1e9);
    t = run(k_fq<2>, blocks, threads, d_out);
    printf(", \"fq_addsub_weak_%dx%d_G\": %.2f", c[0], c[1], n * (ITERS / 4) * 2 / t / 1e9);
    t = run(k_fq<3>, blocks, threads, d_out);
    printf(", \"fq_lazy4_weak2_%dx%d_G\": %.2f", c[0], c[1], n * (ITERS / 4) / t / 1e9);
    t = run(k_mix_shf, blocks, threads, d_out);
    printf(", \"mix_wide_shf_%dx%d_Tpairs\": %.3f", c[0], c[1], n * ITERS * 32 / t / 1e12);
    if (threads <= 256) {
      t = run(k_gq<0>, blocks, threads, d_out);
      printf(", \"gq_dbl_%dx%d_G\": %.3f", c[0], c[1], n * (ITERS / 32) / t / 1e9);
      t = run(k_gq<1>, blocks, threads, d_out);
      printf(", \"gq_madd_%dx%d_G\": %.3f", c[0], c[1], n * (ITERS / 32) / t / 1e9);
      t = run(k_gq<2>, blocks, threads, d_out);
      printf(", \"gq_dbl_inl_%dx%d_G\": %.3f", c[0], c[1], n * (ITERS / 32) / t / 1e9);
      t = run(k_gq<3>, blocks, threads, d_out);
      printf(", \"gq_madd_inl_%dx%d_G\": %.3f", c[0], c[1], n * (ITERS / 32) / t / 1e9);
      t = run(k_ge<0>, blocks, threads, d_out);
      printf(", \"jac_dbl_%dx%d_G\": %.3f", c[0], c[1], n * (ITERS / 32) / t / 1e9);
      t = run(k_ge<1>, blocks, threads, d_out);
      printf(", \"jac_madd_%dx%d_G\": %.3f", c[0], c[1], n * (ITERS / 32) / t / 1e9);
    }
  }
  printf("}\n");
  return 0;
}
