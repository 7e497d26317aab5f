// ecdsa_k256_smem.cuh -- the secp256k1 double-scalar core with the accumulator kept in shared memory.
//
// Same schedule and same results as k256_dsm (ecdsa_k256_body.cuh); what changes is where the running point
// lives.  With acc = (X, Y, Z) in registers and passed by value through the out-of-line double / mixed-add
// the kernel needs 162 registers, i.e. 3 warps per scheduler, and ncu shows the multiplier pipe 75 % busy
// with `wait` as the top stall.  Here X, Y, Z (and the table's shared Z) sit in shared memory, word-major so
// that a warp's 32 lanes hit 32 banks, and each formula loads a coordinate when it needs it and stores a
// result as soon as it is final; operands from the tables are read from global memory at their point of
// use.  That caps the live state of double / add at four to five field elements and lets the kernel run at
// 128 registers (4 warps per scheduler).  Build with -DEB_VERIFY_SMEM=1 -DEB_VERIFY_MINBLOCKS=4 to select it.
#pragma once
#include "ecdsa_k256_body.cuh"

namespace eb {

// STRIDE = threads per block on the device (word w of a field element is at p[w * STRIDE]); 1 on the host.
template <int STRIDE> EB_HD fe sm_ld(const u32* p) {
  fe r;
#pragma unroll
  for (int w = 0; w < 8; w++) r.v[w] = p[w * STRIDE];
  return r;
}
template <int STRIDE> EB_HD void sm_st(u32* p, const fe& a) {
#pragma unroll
  for (int w = 0; w < 8; w++) p[w * STRIDE] = a.v[w];
}

constexpr int SM_X = 0, SM_Y = 8, SM_Z = 16, SM_ZG = 24;   // field slots (x STRIDE words each)
constexpr int SM_WORDS = 32;                               // per thread

// acc <- 2 acc   (dbl-2009-l, as jac_dbl_inl)
template <int STRIDE>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
void jac_dbl_sm(u32* acc) {
  fe Y = sm_ld<STRIDE>(acc + SM_Y * STRIDE);
  {
    fe Z = sm_ld<STRIDE>(acc + SM_Z * STRIDE);
    sm_st<STRIDE>(acc + SM_Z * STRIDE, fe_dbl(fe_mul(Y, Z)));
  }
  fe B = fe_sqr(Y);
  fe X = sm_ld<STRIDE>(acc + SM_X * STRIDE);
  fe A = fe_sqr(X);
  fe t = fe_sqr(fe_add(X, B));
  fe C = fe_sqr(B);
  t = fe_sub(fe_sub(t, A), C);
  fe D = fe_dbl(t);
  fe E = fe_mul_small(A, 3);
  fe x3 = fe_sub(fe_sqr(E), fe_dbl(D));
  sm_st<STRIDE>(acc + SM_X * STRIDE, x3);
  fe y3 = fe_sub(fe_mul(E, fe_sub(D, x3)), fe_mul_small(C, 8));
  sm_st<STRIDE>(acc + SM_Y * STRIDE, y3);
}

// acc <- acc + (px, +-py) for an affine table entry read at its point of use; exceptional cases exact.
template <int STRIDE>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
void jac_madd_sm(u32* acc, const u32* px, const u32* py, bool neg) {
  fe Z = sm_ld<STRIDE>(acc + SM_Z * STRIDE);
  fe z2 = fe_sqr(Z);
  fe u2 = fe_mul(load_fe(px), z2);
  fe s2;
  {
    fe y = load_fe(py);
    y = fe_cmov(y, fe_neg(y), neg);
    s2 = fe_mul(fe_mul(y, z2), Z);
  }
  fe X = sm_ld<STRIDE>(acc + SM_X * STRIDE);
  fe h = fe_sub(X, u2);
  fe z3 = fe_mul(Z, h);
  if (fe_is_zero(z3)) {                       // cold: acc == O, or h == 0 -- redo with the all-cases formula
    ge_jac a;
    a.x = X; a.y = sm_ld<STRIDE>(acc + SM_Y * STRIDE); a.z = Z;
    ge_aff p;
    p.x = load_fe(px); p.y = load_fe(py);
    ge_jac r = jac_madd_inl(a, aff_neg_if(p, neg));
    sm_st<STRIDE>(acc + SM_X * STRIDE, r.x); sm_st<STRIDE>(acc + SM_Y * STRIDE, r.y); sm_st<STRIDE>(acc + SM_Z * STRIDE, r.z);
    return;
  }
  sm_st<STRIDE>(acc + SM_Z * STRIDE, z3);
  fe rr = fe_sub(sm_ld<STRIDE>(acc + SM_Y * STRIDE), s2);
  fe h2 = fe_sqr(h);
  fe h3 = fe_mul(h2, h);
  fe v = fe_mul(X, h2);
  fe x3 = fe_sub(fe_sub(fe_add(fe_sqr(rr), h3), v), v);
  sm_st<STRIDE>(acc + SM_X * STRIDE, x3);
  fe y3 = fe_sub(fe_mul(rr, fe_sub(v, x3)), fe_mul(sm_ld<STRIDE>(acc + SM_Y * STRIDE), h3));
  sm_st<STRIDE>(acc + SM_Y * STRIDE, y3);
}

// u1*G + u2*Q into acc (shared memory); same table layout and digit recoding as k256_dsm.
template <int STRIDE>
EB_HD void k256_dsm_sm(size_t i, size_t N, const ge_aff& Q, u32 flags, const u32* ws, const u32* gtab, u32* qtab,
                       u32* acc) {
  u32* tab = qtab + (size_t)i * QTAB_WORDS;
  {
    ge_jac D = jac_dbl(jac_from_aff(Q));
    fe C2 = fe_sqr(D.z);
    fe C3 = fe_mul(C2, D.z);
    ge_aff Dp; Dp.x = D.x; Dp.y = D.y;
    ge_jac P;
    P.x = fe_mul(Q.x, C2);
    P.y = fe_mul(Q.y, C3);
    P.z = fe_one();
    store_fe(tab + 0, P.x); store_fe(tab + 8, P.y);
    for (int k = 1; k < QTAB_ENTRIES; k++) {
      madd_out o = jac_madd_h(P, Dp);
      P = o.r;
      store_fe(tab + 24 * k, P.x); store_fe(tab + 24 * k + 8, P.y);
      store_fe(tab + 24 * k + 16, o.h);
    }
    sm_st<STRIDE>(acc + SM_ZG * STRIDE, fe_mul(P.z, D.z));
    fe beta = fe_beta();
    fe zs = fe_one();
    for (int k = QTAB_ENTRIES - 1; k >= 0; k--) {
      fe X = load_fe(tab + 24 * k), Y = load_fe(tab + 24 * k + 8);
      fe hk = fe_one();
      if (k > 0) hk = load_fe(tab + 24 * k + 16);
      if (k < QTAB_ENTRIES - 1) {
        fe zs2 = fe_sqr(zs);
        fe zs3 = fe_mul(zs2, zs);
        X = fe_mul(X, zs2);
        Y = fe_mul(Y, zs3);
        store_fe(tab + 24 * k, X); store_fe(tab + 24 * k + 8, Y);
      }
      store_fe(tab + 24 * k + 16, fe_mul(X, beta));
      zs = fe_mul(zs, hk);
    }
  }
  for (int w = 32; w >= 0; w--) {
    if (w != 32)
      for (int d = 0; d < 4; d++) jac_dbl_sm<STRIDE>(acc);
    for (int h = 0; h < 2; h++) {
      u32 word = ws[(size_t)((h ? 13 : 8) + (w >> 3)) * N + i];
      u32 nib = (word >> (4 * (w & 7))) & 15;
      bool dneg = (w != 32) && (nib < 8);
      u32 idx = (w == 32) ? (nib & 7) : (dneg ? 7 - nib : nib - 8);
      bool neg = dneg != (((flags & (h ? FL_NEG2 : FL_NEG1)) != 0));
      const u32* px = tab + 24 * idx + (h ? 16 : 0);
      const u32* py = tab + 24 * idx + 8;
      if (w == 32 && h == 0) {
        fe y = load_fe(py);
        sm_st<STRIDE>(acc + SM_X * STRIDE, load_fe(px));
        sm_st<STRIDE>(acc + SM_Y * STRIDE, fe_cmov(y, fe_neg(y), neg));
        sm_st<STRIDE>(acc + SM_Z * STRIDE, fe_one());
      } else {
        jac_madd_sm<STRIDE>(acc, px, py, neg);
      }
    }
  }
  sm_st<STRIDE>(acc + SM_Z * STRIDE, fe_mul(sm_ld<STRIDE>(acc + SM_Z * STRIDE), sm_ld<STRIDE>(acc + SM_ZG * STRIDE)));
  if (flags & FL_NOG) return;
  for (int j = 0; j < GTAB_WINDOWS; j++) {
    const int pos = GTAB_W * j;
    u32 lo = ws[(size_t)(pos >> 5) * N + i];
    u32 hi = ((pos >> 5) < 7) ? ws[(size_t)((pos >> 5) + 1) * N + i] : 0u;
    u64 both = ((u64)hi << 32) | lo;
    u32 chunk = (u32)(both >> (pos & 31)) & ((1u << GTAB_W) - 1);
    const u32 half = 1u << (GTAB_W - 1);
    bool dneg = (j != GTAB_WINDOWS - 1) && (chunk < half);
    u32 idx = (j == GTAB_WINDOWS - 1) ? (chunk & (half - 1)) : (dneg ? half - 1 - chunk : chunk - half);
    bool neg = dneg != ((flags & FL_NEGG) != 0);
    const u32* ent = gtab + ((size_t)j * GTAB_ENTRIES + idx) * 16;
    jac_madd_sm<STRIDE>(acc, ent, ent + 8, neg);
  }
}

template <int STRIDE>
EB_HD uint8_t verify_item_sm(size_t i, size_t N, const uint8_t* pub, const uint8_t* r, const u32* ws, const u32* gtab,
                             u32* qtab, u32* acc) {
  u32 flags = ws[(size_t)18 * N + i];
  if (flags & FL_INVALID) return ST_FALSE;
  ge_aff Q;
  Q.x = fe_from_be(pub + 64 * i);
  Q.y = fe_from_be(pub + 64 * i + 32);
  if (!aff_on_curve(Q)) return ST_NEEDS_HOST;
  k256_dsm_sm<STRIDE>(i, N, Q, flags, ws, gtab, qtab, acc);
  fe Z = sm_ld<STRIDE>(acc + SM_Z * STRIDE);
  if (fe_is_zero(Z)) return ST_FALSE;
  fe z2 = fe_sqr(Z);
  fe X = sm_ld<STRIDE>(acc + SM_X * STRIDE);
  fe rf = fe_from_be(r + 32 * i);
  if (fe_eq(X, fe_mul(rf, z2))) return ST_TRUE;
  const u32 pmn[8] = {0x2fc9baeeu, 0x402da172u, 0x50b75fc4u, 0x45512319u, 0x00000001u, 0, 0, 0};  // p - n
  if (!geq_n<8>(rf.v, pmn)) {
    u32 nn[8]; K256N::n(nn);
    fe rn;
    add_n<8>(rn.v, rf.v, nn);
    if (fe_eq(X, fe_mul(rn, z2))) return ST_TRUE;
  }
  return ST_FALSE;
}

}  // namespace eb
