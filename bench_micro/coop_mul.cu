// coop_mul.cu -- layout experiment behind DESIGN.md's "thread per item" decision.
//
// BASELINE.json's north-star sketches "one warp per scalar-mul ... CIOS Montgomery with warp shuffles for
// carry propagation"; SURVEY 7 asks for both layouts to be measured.  This micro-benchmark times the same
// p256 / p384 Montgomery multiplication (CIOS, 32-bit limbs) two ways on all SMs:
//   T: one thread owns all N limbs (fp_mont.cuh, the product's code path)
//   W: N lanes of a warp own one limb each (4 or 2 multiplications per warp), operand limbs broadcast and
//      the running sum shifted one lane down per row with __shfl_sync, carries kept per lane in a 64-bit
//      accumulator and resolved once at the end
// and checks W against T value by value.  Output: one JSON object (multiplications per second).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../elliptic_b200/csrc/ecdsa_sw_body.cuh"
using namespace eb;

#define ITERS 2048

template <class P>
__global__ void k_thread(u32* out, u32 seed, int check) {
  typedef Fp<P> F;
  constexpr int N = P::N;
  typename F::fe a, b;
  for (int i = 0; i < N; i++) { a.v[i] = seed * (i + 3) + threadIdx.x * 2654435761u + blockIdx.x; b.v[i] = seed * (i + 7) ^ (threadIdx.x * 40503u + blockIdx.x * 977u); }
  a.v[N - 1] >>= 1; b.v[N - 1] >>= 1;                         // below p
  for (int it = 0; it < (check ? 4 : ITERS); it++) { a = F::mul(a, b); b = F::mul(b, a); }
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (check) for (int i = 0; i < N; i++) out[t * N + i] = a.v[i];
  else { u32 s = 0; for (int i = 0; i < N; i++) s ^= a.v[i] ^ b.v[i]; out[t] = s; }
}

// One Montgomery product with limb j of every operand in lane j of a GROUP-lane segment.
template <class P, int GROUP>
__device__ __forceinline__ u32 coop_mul(u32 a, u32 b, u32 p, u32 n0inv, int lane) {
  constexpr int N = P::N;
  const unsigned full = 0xffffffffu;
  u64 t = 0;                                                   // this lane's column: digit + pending carries
#pragma unroll
  for (int i = 0; i < N; i++) {
    u32 ai = __shfl_sync(full, a, i, GROUP);
    u64 prod = (u64)ai * b + (u32)t;                           // fits: (2^32-1)^2 + 2^32-1 < 2^64
    u64 carry = (t >> 32) + (prod >> 32);
    u32 lo0 = __shfl_sync(full, (u32)prod, 0, GROUP);
    u32 m = lo0 * n0inv;
    u64 red = (u64)m * p + (u32)prod;                          // lane 0: low word becomes 0
    carry += red >> 32;
    u32 down = __shfl_down_sync(full, (u32)red, 1, GROUP);     // column j+1 moves to lane j
    if (lane >= N - 1) down = 0;
    t = carry + down;
  }
  // resolve the pending carries: two ripple rounds leave 0/1 carries, a ballot-based carry-lookahead
  // (generate = pending carry, propagate = digit all ones) settles the rest in one step
  const int gbase = (threadIdx.x & 31) - lane;                 // first lane of this group within the warp
  const u32 gmask = (N == 32) ? 0xffffffffu : ((1u << N) - 1);
  u32 over = 0;
#pragma unroll
  for (int r = 0; r < 2; r++) {
    u32 c = (u32)(t >> 32);
    u32 up = __shfl_up_sync(full, c, 1, GROUP);
    if (lane == 0) up = 0;
    if (lane == N - 1) over += c;
    t = (u64)(u32)t + up;
  }
  u32 digit = (u32)t;
  u32 c1 = (u32)(t >> 32);
  {
    u32 G = (__ballot_sync(full, c1 != 0) >> gbase) & gmask;
    u32 Pm = (__ballot_sync(full, digit == 0xffffffffu) >> gbase) & gmask;
    u32 X = (Pm | G), Y = G;                                   // bit j of (X + Y) ^ X ^ Y = carry into lane j
    u32 sum = X + Y;
    u32 cin = sum ^ X ^ Y;
    digit += (cin >> lane) & 1;
    over += (cin >> N) & 1;
  }
  u32 top = __shfl_sync(full, over, N - 1, GROUP);             // overflow word of the N+1-limb result
  // conditional subtraction of p with the same lookahead on borrows
  u32 d = digit - p;
  {
    u32 G = (__ballot_sync(full, digit < p) >> gbase) & gmask;
    u32 Pm = (__ballot_sync(full, digit == p) >> gbase) & gmask;
    u32 X = (Pm | G), Y = G;
    u32 sum = X + Y;
    u32 bin = sum ^ X ^ Y;
    d -= (bin >> lane) & 1;
    u32 final_borrow = (bin >> N) & 1;
    bool ge = top || !final_borrow;
    if (lane >= N) return 0;                                   // idle lanes of a partly filled group
    return ge ? d : digit;
  }
}

template <class P, int GROUP>
__global__ void k_warp(u32* out, u32 seed, int check) {
  constexpr int N = P::N;
  const int lane = threadIdx.x % GROUP;
  const size_t item = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / GROUP;
  // the same operands item `item` gets in k_thread when its thread index equals `item`
  const u32 tix = (u32)(item % blockDim.x), bix = (u32)(item / blockDim.x);
  u32 pm[N]; P::mod(pm);
  u32 p = 0, a = 0, b = 0;
  for (int i = 0; i < N; i++) if (i == lane) {
    p = pm[i];
    a = seed * (i + 3) + tix * 2654435761u + bix;
    b = seed * (i + 7) ^ (tix * 40503u + bix * 977u);
    if (i == N - 1) { a >>= 1; b >>= 1; }
  }
  for (int it = 0; it < (check ? 4 : ITERS); it++) {
    a = coop_mul<P, GROUP>(a, b, p, P::n0inv, lane);
    b = coop_mul<P, GROUP>(b, a, p, P::n0inv, lane);
  }
  if (check) { if (lane < N) out[item * N + lane] = a; }
  else out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a ^ b;
}

template <class K>
static double time_kernel(K launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 3; r++) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3 / 3;
}

template <class P, int GROUP>
static void run(const char* name, int sms, u32* d_a, u32* d_b) {
  constexpr int N = P::N;
  // correctness: 64 blocks x 128 items
  const int cb = 64, ct = 128;
  k_thread<P><<<cb, ct>>>(d_a, 12345u, 1);
  k_warp<P, GROUP><<<cb * GROUP, ct>>>(d_b, 12345u, 1);
  cudaDeviceSynchronize();
  size_t words = (size_t)cb * ct * N;
  u32* ha = (u32*)malloc(words * 4); u32* hb = (u32*)malloc(words * 4);
  cudaMemcpy(ha, d_a, words * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hb, d_b, words * 4, cudaMemcpyDeviceToHost);
  size_t bad = 0;
  for (size_t i = 0; i < words; i++) bad += ha[i] != hb[i];
  if (bad) {   // diagnostic: the first differing item, limb by limb
    size_t it = 0;
    for (size_t i = 0; i < words; i++) if (ha[i] != hb[i]) { it = i / N; break; }
    fprintf(stderr, "%s first differing item %zu\n thread:", name, it);
    for (int i = 0; i < N; i++) fprintf(stderr, " %08x", ha[it * N + i]);
    fprintf(stderr, "\n warp:  ");
    for (int i = 0; i < N; i++) fprintf(stderr, " %08x", hb[it * N + i]);
    fprintf(stderr, "\n");
  }
  free(ha); free(hb);
  double best_t = 0, best_w = 0;
  int cfgs[][2] = {{4, 128}, {4, 256}, {8, 128}, {2, 512}};
  for (auto& c : cfgs) {
    int blocks = sms * c[0], threads = c[1];
    double t = time_kernel([&] { k_thread<P><<<blocks, threads>>>(d_a, 99u, 0); });
    double r = (double)blocks * threads * ITERS * 2 / t;
    if (r > best_t) best_t = r;
    t = time_kernel([&] { k_warp<P, GROUP><<<blocks, threads>>>(d_b, 99u, 0); });
    r = (double)blocks * threads / GROUP * ITERS * 2 / t;
    if (r > best_w) best_w = r;
  }
  printf(",\n \"%s\": {\"limbs\": %d, \"lanes_per_item\": %d, \"mismatching_words\": %zu, \"thread_per_item_Gmul_s\": %.2f, "
         "\"warp_cooperative_Gmul_s\": %.2f, \"ratio\": %.2f}",
         name, N, GROUP, bad, best_t / 1e9, best_w / 1e9, best_t / best_w);
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
  int sms = prop.multiProcessorCount;
  u32 *d_a, *d_b;
  cudaMalloc(&d_a, (size_t)sms * 8 * 1024 * 16 * 4);
  cudaMalloc(&d_b, (size_t)sms * 8 * 1024 * 16 * 4);
  printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
  run<P256_FP, 8>("p256", sms, d_a, d_b);
  run<P384_FP, 16>("p384", sms, d_a, d_b);
  printf("}\n");
  return 0;
}
