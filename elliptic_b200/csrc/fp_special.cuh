// fp_special.cuh -- prime-field arithmetic for the NIST primes with their own fast reductions:
//   p256 = 2^256 - 2^224 + 2^192 + 2^96 - 1,  p384 = 2^384 - 2^128 - 2^96 + 2^32 - 1,  p521 = 2^521 - 1.
//
// The reference runs these curves on bn.js `Mont` (lib/elliptic/curves.js:73-134 give `prime: null` ->
// BN.mont(p), curve/base.js:14; Mont.mul / imul dist/elliptic.js:7312-7381): every product is followed by an
// N x N multiplication by the modulus.  Only canonical residues are observable (fromRed), so the engine is
// free to hold PLAIN residues and reduce the double-width product with the primes' word structure instead
// (FIPS 186-4 D.2): word additions / subtractions of the high half for p256 / p384, one 521-bit fold for
// p521.  Against the generic CIOS of fp_mont.cuh that drops N^2 + N of the 2 N^2 + N multiplies of a product
// and lets squarings use the N (N + 1) / 2-multiply generated squarer (tools/gen_sqr.py).
//
// Same interface as Fp<P> (fp_mont.cuh), so SW<C> (ecdsa_sw_body.cuh) is unchanged: `to_mont` reduces a raw
// value mod p (the reference's toRed, dist:7292-7296), `from_mont` is the identity, `one()` is 1.
#pragma once
#include "fp_mont.cuh"

namespace eb {

#if defined(__CUDACC__) && !defined(EB_SQR8_INCLUDED)
#define EB_SQR8_INCLUDED
#include "sqr_gen.inc"
#endif

// EB_SOLINAS_COLUMNS=1 (default): the p256 / p384 reductions read the standard's word vectors down their columns
// (tools/gen_solinas.py -> solinas_gen.inc); 0: the vector-wise carry chains below (round-2 first version).
// EB_SOLINAS_SCALED=1: the doubling's constants (3, 4, 8) are applied inside the reduction (F::mul_k / sqr_k through one
// extra out-of-line body).  Measured in round 2 and left off: p256 43.9 vs 44.0 ms, p384 143.3 vs 139.8 ms -- the extra
// body costs the instruction cache what the saved modular doublings gain.
// EB_FPS_WEAK=1: p256 / p384 elements are held WEAKLY reduced -- any representative in [0, 2^(32N)) -- the way
// fe_k256.cuh holds secp256k1's: products need no final subtraction, add / sub fold the carry / borrow with
// 2^(32N) = K (a masked N-word add instead of subtract-compare-select), only is_zero / eq / from_mont look at the
// value mod p.  Bit-exact (host emulation and the GPU suites pass with it) and SLOWER, so off: ptxas predicates the
// rare second fold of every add / sub instead of branching around it, the group-law bodies grow (p256 dbl 529 ->
// 615 instructions, add 1038 -> 1365) past the instruction-cache budget: p256 50.4 vs 44.1 ms, p384 141.5 vs 139.5.
#ifndef EB_FPS_WEAK
#define EB_FPS_WEAK 0
#endif
#ifndef EB_SOLINAS_SCALED
#define EB_SOLINAS_SCALED 0
#endif
#ifndef EB_SOLINAS_COLUMNS
#define EB_SOLINAS_COLUMNS 1
#endif
#include "solinas_gen.inc"

// acc (N words) += / -= v, returns the carry / borrow
template <int N> EB_HD int sp_addv(u32* acc, const u32* v) { return (int)add_n<N>(acc, acc, v); }
template <int N> EB_HD int sp_subv(u32* acc, const u32* v) { return (int)sub_n<N>(acc, acc, v); }

// acc += u * K (u small); returns the carry word
template <int N> EB_HD u32 sp_addmul_small(u32* acc, const u32* K, u32 u) {
  u64 c = 0;
#pragma unroll
  for (int j = 0; j < N; j++) {
    c += (u64)K[j] * u + acc[j];
    acc[j] = (u32)c;
    c >>= 32;
  }
  return (u32)c;
}
// (carry, acc) < 2p  ->  canonical residue
template <int N> EB_HD void sp_final(u32* r, const u32* acc, u32 carry, const u32* p) {
  u32 d[N];
  u32 bw = sub_n<N>(d, acc, p);
  bool ge = carry != 0 || bw == 0;
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = ge ? d[i] : acc[i];
}

// v in [0, 2^(32N)), v < 2p, top limb of p all ones  ->  canonical residue.  v >= p needs v's top limb to be all
// ones, which a product's residue class hits with probability 2^-32: the comparison and the subtraction sit in a
// branch that a warp takes (all lanes together) only when one of its lanes is in that sliver.
template <int N> EB_HD void sp_final_rare(u32* v, const u32* p) {
  if (v[N - 1] == 0xffffffffu) {
    u32 d[N];
    if (!sub_n<N>(d, v, p)) {
#pragma unroll
      for (int i = 0; i < N; i++) v[i] = d[i];
    }
  }
}

// r = a + b + c word-wise with the carries counted (a, b, c: N words); returns the carry count (0..2)
template <int N> EB_HD int sp_add3(u32* r, const u32* a, const u32* b, const u32* c) {
  int t = (int)add_n<N>(r, a, b);
  return t + (int)add_n<N>(r, r, c);
}

// ---- p256: r = s1 + 2 s2 + 2 s3 + s4 + s5 - s6 - s7 - s8 - s9  (FIPS 186-4 D.2.3; word vectors (w7..w0)) ----
// The nine vectors are summed as four independent carry chains (two positive, two negative) that are only
// joined at the end: the chains of a naive left-to-right sum are ~90 dependent add-with-carry instructions, and
// with two warps per scheduler that latency, not the instruction count, sets the pace.
struct RedP256 {
  static constexpr int N = 8, WN = 8;
  static constexpr bool SCALED = EB_SOLINAS_COLUMNS != 0 && EB_SOLINAS_SCALED != 0;
  static constexpr bool WEAK = EB_SOLINAS_COLUMNS != 0 && EB_FPS_WEAK != 0;
  static EB_HD void kwords(u32* k) {           // 2^256 mod p = 2^224 - 2^192 - 2^96 + 1
    const u32 K[8] = {0x00000001u, 0, 0, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xfffffffeu, 0};
    for (int i = 0; i < 8; i++) k[i] = K[i];
  }
  static EB_HD void reduce_weak(u32* r, const u32* c, int k = 1) { solinas_p256(r, c, k); }   // [0, 2^256), not < p
  static EB_HD void reduce_scaled(u32* r, const u32* c, const u32* p, int k) {
    solinas_p256(r, c, k);
    sp_final_rare<8>(r, p);
  }
  static EB_HD void reduce(u32* r, const u32* c, const u32* p) {
#if EB_SOLINAS_COLUMNS
    reduce_scaled(r, c, p, 1);
    return;
#endif
    // value = acc + top * 2^256, top in [-4, 5].  2^256 = K (mod p), K = 2^224 - 2^192 - 2^96 + 1:
    // acc + (top + 4) K + (-4 K mod p), all terms non-negative; C4 = -4 K mod p rides along with s1
    const u32 K[8] = {0x00000001u, 0, 0, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xfffffffeu, 0};
    const u32 C4[8] = {0xfffffffbu, 0xffffffffu, 0xffffffffu, 0x00000004u, 0x00000000u, 0x00000000u, 0x00000005u, 0xfffffffbu};
    const u32 s2[8] = {0, 0, 0, c[11], c[12], c[13], c[14], c[15]};
    const u32 s3[8] = {0, 0, 0, c[12], c[13], c[14], c[15], 0};
    const u32 s4[8] = {c[8], c[9], c[10], 0, 0, 0, c[14], c[15]};
    const u32 s5[8] = {c[9], c[10], c[11], c[13], c[14], c[15], c[13], c[8]};
    const u32 s6[8] = {c[11], c[12], c[13], 0, 0, 0, c[8], c[10]};
    const u32 s7[8] = {c[12], c[13], c[14], c[15], 0, 0, c[9], c[11]};
    const u32 s8[8] = {c[13], c[14], c[15], c[8], c[9], c[10], 0, c[12]};
    const u32 s9[8] = {c[14], c[15], 0, c[9], c[10], c[11], 0, c[13]};
    u32 p1[8], p2[8], p3[8], n1[8], n2[8];
    int top = sp_add3<8>(p1, c, C4, s2);               // s1 + C4 + s2
    top += sp_add3<8>(p2, s2, s3, s3);                 // s2 + 2 s3
    top += (int)add_n<8>(p3, s4, s5);
    top -= (int)add_n<8>(n1, s6, s7);
    top -= (int)add_n<8>(n2, s8, s9);
    u32 acc[8];
    top += sp_add3<8>(acc, p1, p2, p3);
    top -= (int)sub_n<8>(acc, acc, n1);
    top -= (int)sub_n<8>(acc, acc, n2);
    u32 cy = sp_addmul_small<8>(acc, K, (u32)(top + 4));
    sp_final<8>(r, acc, cy, p);
  }
};

// ---- p384: r = s1 + 2 s2 + s3 + s4 + s5 + s6 + s7 - s8 - s9 - s10  (FIPS 186-4 D.2.4) ----------------------
struct RedP384 {
  static constexpr int N = 12, WN = 12;
  static constexpr bool SCALED = EB_SOLINAS_COLUMNS != 0 && EB_SOLINAS_SCALED != 0;
  static constexpr bool WEAK = EB_SOLINAS_COLUMNS != 0 && EB_FPS_WEAK != 0;
  static EB_HD void kwords(u32* k) {           // 2^384 mod p = 2^128 + 2^96 - 2^32 + 1
    const u32 K[12] = {0x00000001u, 0xffffffffu, 0xffffffffu, 0, 0x00000001u, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 12; i++) k[i] = K[i];
  }
  static EB_HD void reduce_weak(u32* r, const u32* c, int k = 1) { solinas_p384(r, c, k); }
  static EB_HD void reduce_scaled(u32* r, const u32* c, const u32* p, int k) {
    solinas_p384(r, c, k);
    sp_final_rare<12>(r, p);
  }
  static EB_HD void reduce(u32* r, const u32* c, const u32* p) {
#if EB_SOLINAS_COLUMNS
    reduce_scaled(r, c, p, 1);
    return;
#endif
    // value = acc + top * 2^384, top in [-3, 7].  2^384 = K (mod p), K = 2^128 + 2^96 - 2^32 + 1; C3 = -3 K mod p
    const u32 K[12] = {0x00000001u, 0xffffffffu, 0xffffffffu, 0, 0x00000001u, 0, 0, 0, 0, 0, 0, 0};
    const u32 C3[12] = {0xfffffffcu, 0x00000003u, 0x00000000u, 0xfffffffcu, 0xfffffffbu, 0xffffffffu,
                        0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    const u32 s2[12] = {0, 0, 0, 0, c[21], c[22], c[23], 0, 0, 0, 0, 0};
    const u32 s4[12] = {c[21], c[22], c[23], c[12], c[13], c[14], c[15], c[16], c[17], c[18], c[19], c[20]};
    const u32 s5[12] = {0, c[23], 0, c[20], c[12], c[13], c[14], c[15], c[16], c[17], c[18], c[19]};
    const u32 s6[12] = {0, 0, 0, 0, c[20], c[21], c[22], c[23], 0, 0, 0, 0};
    const u32 s7[12] = {c[20], 0, 0, c[21], c[22], c[23], 0, 0, 0, 0, 0, 0};
    const u32 s8[12] = {c[23], c[12], c[13], c[14], c[15], c[16], c[17], c[18], c[19], c[20], c[21], c[22]};
    const u32 s9[12] = {0, c[20], c[21], c[22], c[23], 0, 0, 0, 0, 0, 0, 0};
    const u32 s10[12] = {0, 0, 0, c[23], c[23], 0, 0, 0, 0, 0, 0, 0};
    u32 p1[12], p2[12], p3[12], n1[12];
    int top = sp_add3<12>(p1, c, C3, c + 12);          // s1 + C3 + s3
    top += sp_add3<12>(p2, s2, s2, s4);
    top += sp_add3<12>(p3, s5, s6, s7);
    top -= sp_add3<12>(n1, s8, s9, s10);
    u32 acc[12];
    top += sp_add3<12>(acc, p1, p2, p3);
    top -= (int)sub_n<12>(acc, acc, n1);
    u32 cy = sp_addmul_small<12>(acc, K, (u32)(top + 3));
    sp_final<12>(r, acc, cy, p);
  }
};

// ---- p521 = 2^521 - 1 in an 18-word container (521 bits = 16 words + 9 bits; word 17 is always 0) -------------
struct RedP521 {
  static constexpr int N = 18, WN = 17;
  static constexpr bool SCALED = false;
  static constexpr bool WEAK = false;
  static EB_HD void kwords(u32*) {}
  static EB_HD void reduce_weak(u32* r, const u32* c, int = 1) { reduce(r, c, nullptr); }                 // never used
  static EB_HD void reduce_scaled(u32* r, const u32* c, const u32* p, int) { reduce(r, c, p); }   // never used
  static EB_HD void reduce(u32* r, const u32* c, const u32* /*p*/) {
    // c < 2^1042 (36 words, the top ones zero): (c mod 2^521) + (c >> 521), twice, then p -> 0
    u32 lo[17], hi[17];
#pragma unroll
    for (int i = 0; i < 16; i++) lo[i] = c[i];
    lo[16] = c[16] & 0x1ffu;
#pragma unroll
    for (int i = 0; i < 17; i++) hi[i] = (c[16 + i] >> 9) | (c[17 + i] << 23);
    add_n<17>(lo, lo, hi);                       // < 2^522: no carry out of word 16
    u32 k = lo[16] >> 9;
    lo[16] &= 0x1ffu;
    u32 one[17];
#pragma unroll
    for (int i = 0; i < 17; i++) one[i] = i == 0 ? k : 0u;
    add_n<17>(lo, lo, one);                      // <= 2^521 - 1 (see the derivation in DESIGN.md: the sum was <= 2^522 - 2)
#pragma unroll
    for (int i = 0; i < 17; i++) r[i] = lo[i];
    r[17] = 0;
    if (lo[15] == 0xffffffffu) {                 // lo == p (all 521 bits set) -> 0; one word screens it out (2^-32)
      u32 all = lo[16] ^ 0x1ffu;
#pragma unroll
      for (int i = 0; i < 16; i++) all |= ~lo[i];
      if (all == 0) {
#pragma unroll
        for (int i = 0; i < 17; i++) r[i] = 0;
      }
    }
  }
};

template <class P, class RED>
struct FpS {
  static constexpr int N = P::N;
  typedef P Params;
  typedef fe_n<N> fe;
  typedef Fp<P> G;                       // generic canonical add / sub / compare

  static EB_HD fe zero() { fe r; for (int i = 0; i < N; i++) r.v[i] = 0; return r; }
  static EB_HD fe one() { fe r = zero(); r.v[0] = 1; return r; }

  static EB_HD void wide_mul(u32* t, const fe& a, const fe& b) {
    mul_wide<N>(t, a.v, b.v);
    t[2 * N] = 0; t[2 * N + 1] = 0;
  }
  static EB_HD void wide_sqr(u32* t, const fe& a) {
#if defined(__CUDA_ARCH__) && !defined(EB_SQR_AS_MUL)
    if (RED::WN == 8) sqr_wide8_ptx(t, a.v);
    else if (RED::WN == 12) sqr_wide12_ptx(t, a.v);
    else { sqr_wide17_ptx(t, a.v); t[34] = 0; t[35] = 0; }
#else
    mul_wide<N>(t, a.v, a.v);
#endif
    t[2 * N] = 0; t[2 * N + 1] = 0;
  }
  static EB_HD fe mul_inl(const fe& a, const fe& b) {
    u32 t[2 * N + 2], p[N];
    P::mod(p);
    wide_mul(t, a, b);
    fe r;
    if (RED::WEAK) RED::reduce_weak(r.v, t);
    else RED::reduce(r.v, t, p);
    return r;
  }
  static EB_HD fe sqr_inl(const fe& a) {
    u32 t[2 * N + 2], p[N];
    P::mod(p);
    wide_sqr(t, a);
    fe r;
    if (RED::WEAK) RED::reduce_weak(r.v, t);
    else RED::reduce(r.v, t, p);
    return r;
  }
  // k a b, k in {3, 4, 8} (the constants of the a = -3 doubling): the factor rides through the column sums of the
  // reduction (RED::SCALED) instead of two or three modular doublings of the result.  ONE out-of-line body with k
  // as a run-time argument, and k a^2 goes through it as k a a: three specialised bodies (r02: mulk<3>, mulk<4>,
  // sqrk<8>, +720 instructions) pushed the p256 window loop past the instruction cache (hit rate 99 % -> 88 %,
  // 44.0 -> 46.5 ms) and lost more than the doublings cost.
  static EB_HD fe mulk_inl(const fe& a, const fe& b, int k) {
    u32 t[2 * N + 2], p[N];
    P::mod(p);
    wide_mul(t, a, b);
    fe r;
    if (RED::WEAK) RED::reduce_weak(r.v, t, k);
    else RED::reduce_scaled(r.v, t, p, k);
    return r;
  }
#if defined(__CUDACC__)
  static __device__ __noinline__ fe mul_ol(fe a, fe b) { return mul_inl(a, b); }
  static __device__ __noinline__ fe sqr_ol(fe a) { return sqr_inl(a); }
  static __device__ __noinline__ fe mulk_ol(fe a, fe b, int k) { return mulk_inl(a, b, k); }
#endif
  template <int K> static EB_HD fe scale_k(const fe& r) {          // K r by this field's own additions
    static_assert(K == 1 || K == 3 || K == 4 || K == 8, "scale");
    if (K == 3) return add(add(r, r), r);
    if (K == 4) { fe t = add(r, r); return add(t, t); }
    if (K == 8) { fe t = add(r, r); t = add(t, t); return add(t, t); }
    return r;
  }
  template <int K> static EB_HD fe mul_k(const fe& a, const fe& b) {
    if (!RED::SCALED) return scale_k<K>(mul(a, b));
#if defined(__CUDA_ARCH__) && !defined(EB_MONT_INLINE)
    return mulk_ol(a, b, K);
#else
    return mulk_inl(a, b, K);
#endif
  }
  template <int K> static EB_HD fe sqr_k(const fe& a) {
    if (!RED::SCALED) return scale_k<K>(sqr(a));
    return mul_k<K>(a, a);
  }
  static EB_HD fe mul(const fe& a, const fe& b) {
#if defined(__CUDA_ARCH__) && !defined(EB_MONT_INLINE)
    return mul_ol(a, b);
#else
    return mul_inl(a, b);
#endif
  }
  static EB_HD fe sqr(const fe& a) {
#if defined(__CUDA_ARCH__) && !defined(EB_MONT_INLINE)
    return sqr_ol(a);
#else
    return sqr_inl(a);
#endif
  }

  // weak forms: a carry out of a + b is worth K = 2^(32N) mod p, a borrow out of a - b is worth -K.  A second
  // carry / borrow needs the first result within K of the wrap (2^-32 for p256, 2^-256 for p384): rare branch.
  static EB_HD fe wadd(const fe& a, const fe& b) {
    fe r;
    u32 k[N], km[N];
    RED::kwords(k);
    u32 m = 0u - add_n<N>(r.v, a.v, b.v);
#pragma unroll
    for (int i = 0; i < N; i++) km[i] = k[i] & m;
    if (add_n<N>(r.v, r.v, km)) add_n<N>(r.v, r.v, k);
    return r;
  }
  static EB_HD fe wsub(const fe& a, const fe& b) {
    fe r;
    u32 k[N], km[N];
    RED::kwords(k);
    u32 m = 0u - sub_n<N>(r.v, a.v, b.v);
#pragma unroll
    for (int i = 0; i < N; i++) km[i] = k[i] & m;
    if (sub_n<N>(r.v, r.v, km)) sub_n<N>(r.v, r.v, k);
    return r;
  }
  // the representative in [0, p): only a top limb of all ones can be >= p (both primes' top limbs are all ones)
  static EB_HD fe canon(const fe& a) {
    if (!RED::WEAK) return a;
    fe r = a;
    u32 p[N];
    P::mod(p);
    sp_final_rare<N>(r.v, p);
    return r;
  }
  static EB_HD fe add(const fe& a, const fe& b) { return RED::WEAK ? wadd(a, b) : G::add(a, b); }
  static EB_HD fe sub(const fe& a, const fe& b) { return RED::WEAK ? wsub(a, b) : G::sub(a, b); }
  static EB_HD fe neg(const fe& a) { return sub(zero(), a); }
  static EB_HD fe dbl(const fe& a) { return add(a, a); }
  static EB_HD bool is_zero(const fe& a) {
    if (is_zero_n<N>(a.v)) return true;
    if (!RED::WEAK || a.v[N - 1] != 0xffffffffu) return false;
    u32 p[N];
    P::mod(p);
    return eq_n<N>(a.v, p);                  // 2p > 2^(32N): 0 and p are the only representatives of zero
  }
  static EB_HD bool eq(const fe& a, const fe& b) { return RED::WEAK ? is_zero(wsub(a, b)) : eq_n<N>(a.v, b.v); }
  static EB_HD fe cmov(const fe& a, const fe& b, bool c) { return G::cmov(a, b, c); }

  // raw integer (< 2^(32N)) -> residue (`toRed`): reduce as a double-width value whose high half is zero
  static EB_HD fe to_mont(const fe& a) {
    u32 t[2 * N + 2], p[N];
    P::mod(p);
#pragma unroll
    for (int i = 0; i < 2 * N + 2; i++) t[i] = i < N ? a.v[i] : 0u;
    fe r;
    RED::reduce(r.v, t, p);
    return r;
  }
  static EB_HD fe from_mont(const fe& a) { return canon(a); }
  static EB_HD bool geq_mod(const u32* a) { u32 p[N]; P::mod(p); return geq_n<N>(a, p); }

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
  static EB_HD fe inv(const fe& a) {
    u32 e[N], two[N];
    P::mod(e);
    for (int i = 0; i < N; i++) two[i] = i == 0 ? 2u : 0u;
    sub_n<N>(e, e, two);
    return pow(a, e);
  }
};

}  // namespace eb
