// addsub_check.cu -- differential test ON THE DEVICE of the masked add / sub (Fp<P>::add_fast / sub_fast, the
// EB_MONT_FAST_ADDSUB option of fp_mont.cuh) against the compare-and-select forms, for every Montgomery modulus whose
// top limb is 0xFFFFFFFF.  Operands: uniformly random reduced values and the edges (0, 1, p-1, p-2, R-p, values with
// all-ones upper limbs).  Prints one JSON object: mismatch counts per modulus.
#include <cuda_runtime.h>
#include <stdio.h>
#include "../elliptic_b200/csrc/sw_params.cuh"
#include "../elliptic_b200/csrc/sc_k256.cuh"
using namespace eb;

template <class P>
__global__ void k_check(unsigned long long* bad, u32 seed) {
  typedef Fp<P> F;
  constexpr int N = F::N;
  u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
  u32 s = seed + idx * 2654435761u;
  u32 p[N]; P::mod(p);
  typename F::fe a, b;
  for (int k = 0; k < N; k++) {
    s = s * 1664525u + 1013904223u; a.v[k] = s;
    s = s * 1664525u + 1013904223u; b.v[k] = s;
  }
  int ea = idx & 15, eb_ = (idx >> 4) & 15;                  // 1/16 of the lanes per edge class, crossed
  auto edge = [&](typename F::fe& x, int e) {
    if (e == 1) for (int k = 0; k < N; k++) x.v[k] = 0;
    if (e == 2) { for (int k = 0; k < N; k++) x.v[k] = 0; x.v[0] = 1; }
    if (e == 3) { for (int k = 0; k < N; k++) x.v[k] = p[k]; x.v[0] -= 1; }       // p - 1 (p odd)
    if (e == 4) { for (int k = 0; k < N; k++) x.v[k] = p[k]; x.v[0] -= 2; }
    if (e == 5) { u32 r1[N]; P::r1(r1); for (int k = 0; k < N; k++) x.v[k] = r1[k]; }   // R - p
    if (e == 6) for (int k = N / 2; k < N; k++) x.v[k] = 0xffffffffu;             // upper half all ones
    if (e == 7) for (int k = 1; k < N; k++) x.v[k] = p[k];                          // differs from p only in limb 0
  };
  edge(a, ea); edge(b, eb_);
  typename F::fe d;
  if (!sub_n<N>(d.v, a.v, p)) a = d;                          // reduce into [0, p)  (R < 2p for these moduli)
  if (!sub_n<N>(d.v, b.v, p)) b = d;
  // compare-and-select references
  typename F::fe ra, rs, t;
  {
    u32 cy = add_n<N>(ra.v, a.v, b.v);
    u32 bw = sub_n<N>(t.v, ra.v, p);
    if (cy || !bw) ra = t;
    bw = sub_n<N>(rs.v, a.v, b.v);
    add_n<N>(t.v, rs.v, p);
    if (bw) rs = t;
  }
  typename F::fe fa = F::add_fast(a, b), fs = F::sub_fast(a, b);
  bool ok = true;
  for (int k = 0; k < N; k++) ok = ok && fa.v[k] == ra.v[k] && fs.v[k] == rs.v[k];
  if (!ok) atomicAdd(bad, 1ull);
}

// Operands NOT reduced (uniform in [0, R), edges included): which of the two forms still returns something congruent
// to a + b / a - b?  counts[0] masked != compare-select, [1] compare-select not congruent, [2] masked not congruent
template <class P>
__global__ void k_weak(unsigned long long* counts, u32 seed) {
  typedef Fp<P> F;
  constexpr int N = F::N;
  u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
  u32 s = seed + idx * 2654435761u;
  u32 p[N]; P::mod(p);
  typename F::fe a, b;
  for (int k = 0; k < N; k++) {
    s = s * 1664525u + 1013904223u; a.v[k] = s;
    s = s * 1664525u + 1013904223u; b.v[k] = s;
  }
  if ((idx & 3) == 1) for (int k = N / 2; k < N; k++) a.v[k] = 0xffffffffu;      // a >= p
  if ((idx & 3) == 2) for (int k = N / 2; k < N; k++) { a.v[k] = 0xffffffffu; b.v[k] = 0xffffffffu; }
  auto canon = [&](typename F::fe x) { typename F::fe d; for (int t = 0; t < 3; t++) if (!sub_n<N>(d.v, x.v, p)) x = d; return x; };
  typename F::fe ca = canon(a), cb = canon(b), ra, rs, t;
  {                                                     // the true residues, from canonical copies
    u32 cy = add_n<N>(ra.v, ca.v, cb.v);
    u32 bw = sub_n<N>(t.v, ra.v, p);
    if (cy || !bw) ra = t;
    bw = sub_n<N>(rs.v, ca.v, cb.v);
    add_n<N>(t.v, rs.v, p);
    if (bw) rs = t;
  }
  typename F::fe sa, ss;                                // compare-and-select on the raw operands
  {
    u32 cy = add_n<N>(sa.v, a.v, b.v);
    u32 bw = sub_n<N>(t.v, sa.v, p);
    if (cy || !bw) sa = t;
    bw = sub_n<N>(ss.v, a.v, b.v);
    add_n<N>(t.v, ss.v, p);
    if (bw) ss = t;
  }
  typename F::fe fa = F::add_fast(a, b), fs = F::sub_fast(a, b);
  bool differ = false, slow_bad = false, fast_bad = false;
  typename F::fe csa = canon(sa), css = canon(ss), cfa = canon(fa), cfs = canon(fs);
  for (int k = 0; k < N; k++) {
    differ = differ || fa.v[k] != sa.v[k] || fs.v[k] != ss.v[k];
    slow_bad = slow_bad || csa.v[k] != ra.v[k] || css.v[k] != rs.v[k];
    fast_bad = fast_bad || cfa.v[k] != ra.v[k] || cfs.v[k] != rs.v[k];
  }
  if (differ) atomicAdd(counts, 1ull);
  if (slow_bad) atomicAdd(counts + 1, 1ull);
  if (fast_bad) atomicAdd(counts + 2, 1ull);
}

template <class P>
void run_weak(const char* name, bool first) {
  u32 pm[Fp<P>::N]; P::mod(pm);
  if (pm[Fp<P>::N - 1] != 0xffffffffu) return;
  unsigned long long* d; unsigned long long c[3] = {0, 0, 0};
  cudaMalloc(&d, 24); cudaMemset(d, 0, 24);
  k_weak<P><<<4096, 256>>>(d, 424242u);
  cudaMemcpy(c, d, 24, cudaMemcpyDeviceToHost);
  printf("%s\"%s\": {\"masked_ne_select\": %llu, \"select_wrong_mod_p\": %llu, \"masked_wrong_mod_p\": %llu}", first ? "" : ", ", name, c[0], c[1], c[2]);
  cudaFree(d);
}

template <class P>
unsigned long long run(const char* name, bool first) {
  u32 pm[Fp<P>::N]; P::mod(pm);
  if (pm[Fp<P>::N - 1] != 0xffffffffu) {                    // Fp<P>::add / sub never take the masked path for this modulus
    printf("%s\"%s\": \"not eligible (top limb %08x)\"", first ? "" : ", ", name, pm[Fp<P>::N - 1]);
    return 0;
  }
  unsigned long long* d_bad; unsigned long long bad = 0;
  cudaMalloc(&d_bad, 8); cudaMemset(d_bad, 0, 8);
  for (int rep = 0; rep < 4; rep++) k_check<P><<<4096, 256>>>(d_bad, 1234567u * (rep + 1));
  cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%s\"%s\": %lld", first ? "" : ", ", name, e == cudaSuccess ? (long long)bad : -1ll);
  cudaFree(d_bad);
  return bad;
}

int main() {
  printf("{\"pairs_per_modulus\": %d, \"mismatches\": {", 4 * 4096 * 256);
  run<P192_FP>("p192.p", true); run<P224_FP>("p224.p", false);
  run<P192_FN>("p192.n", false); run<P224_FN>("p224.n", false);
  run<P256_FN>("p256.n", false); run<P384_FN>("p384.n", false); run<K256_FN>("secp256k1.n", false);
  printf("}, \"unreduced_operands_1048576_pairs\": {");
  run_weak<P192_FP>("p192.p", true); run_weak<P256_FN>("p256.n", false); run_weak<P384_FN>("p384.n", false); run_weak<K256_FN>("secp256k1.n", false);
  printf("}}\n");
  return 0;
}
