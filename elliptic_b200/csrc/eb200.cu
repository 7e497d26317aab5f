// eb200.cu -- CUDA kernels (sm_100a) and the C ABI of libelliptic_b200.so.
// See include/elliptic_b200.h for the boundary, ecdsa_k256_body.cuh (secp256k1) and
// ecdsa_sw_body.cuh (p256 / p384) for the algorithms.  No CPU fallback exists in this
// library by design: without a CUDA device every compute entry point fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <cstdlib>
#include <mutex>
#include "../../include/elliptic_b200.h"
#include "ecdsa_k256_body.cuh"
#include "ecdsa_k256_replay.cuh"
#include "ecdsa_sw_replay.cuh"
#include "ecdsa_k256_sign_fast.cuh"
#include "der_sig.cuh"
#include "ecdsa_sw_sign.cuh"

// Calls F(curve-parameter type) for the non-GLV short curve `curve`.
#define SW_DISPATCH(curve, F)                                                                    \
  ((curve) == EB200_CURVE_P256 ? F(P256) : (curve) == EB200_CURVE_P384 ? F(P384) :               \
   (curve) == EB200_CURVE_P521 ? F(P521) : (curve) == EB200_CURVE_P192 ? F(P192) : F(P224))
#include "ecdsa_k256_sign.cuh"
#include "ecdsa_sw_body.cuh"
#include "ed25519_body.cuh"

using namespace eb;

#ifndef EB_VERIFY_BLOCK
#define EB_VERIFY_BLOCK 128
#endif
#ifndef EB_VERIFY_MINBLOCKS
#define EB_VERIFY_MINBLOCKS 3
#endif

// ---------------------------------------------------------------------------
// secp256k1 kernels
__global__ void __launch_bounds__(128) k256_gtab_kernel(u32* gtab) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)GTAB_WINDOWS * GTAB_ENTRIES) return;
  int j = (int)(t / GTAB_ENTRIES), idx = (int)(t % GTAB_ENTRIES);
  gtab_entry(j, idx, gtab + t * 16);
}

__global__ void __launch_bounds__(128) k256_prep_kernel(size_t N, const uint8_t* __restrict__ e,
                                                        const uint8_t* __restrict__ r,
                                                        const uint8_t* __restrict__ s,
                                                        u32* __restrict__ ws, u32* __restrict__ scratch) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t T = (size_t)gridDim.x * blockDim.x;
  prep_thread(tid, T, N, e, r, s, ws, scratch);
}

__global__ void __launch_bounds__(EB_VERIFY_BLOCK, EB_VERIFY_MINBLOCKS)
k256_verify_kernel(size_t N, const uint8_t* __restrict__ pub, const uint8_t* __restrict__ r,
                   const u32* __restrict__ ws, const u32* __restrict__ gtab,
                   u32* __restrict__ qtab, const uint8_t* __restrict__ pre, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (pre && pre[i]) { status[i] = pre[i]; return; }   // the reference throws while importing the key
  status[i] = verify_item(i, N, pub, r, ws, gtab, qtab);
}

__global__ void __launch_bounds__(128) k256_prep_recover_kernel(size_t N, const uint8_t* __restrict__ e,
                                                                const uint8_t* __restrict__ r,
                                                                const uint8_t* __restrict__ s,
                                                                u32* __restrict__ ws, u32* __restrict__ scratch) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t T = (size_t)gridDim.x * blockDim.x;
  prep_thread(tid, T, N, e, r, s, ws, scratch, 1);
}
__global__ void __launch_bounds__(EB_VERIFY_BLOCK, EB_VERIFY_MINBLOCKS)
k256_recover_kernel(size_t N, const uint8_t* __restrict__ r, const uint8_t* __restrict__ recid,
                    const u32* __restrict__ ws, const u32* __restrict__ gtab, u32* __restrict__ qtab,
                    uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  status[i] = recover_item(i, N, r, recid, ws, gtab, qtab, out);
}

// Two-kernel signing pipeline (ecdsa_k256_sign_fast.cuh); k256_sign_slow_kernel redoes flagged items with
// the literal retry loop of k256_sign_item.
__global__ void __launch_bounds__(128)
k256_sign_nonce_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ priv,
                       const u32* __restrict__ gtab, u32* __restrict__ ws, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) k256_sign_nonce_item(i, N, e, priv, gtab, ws, status);
}
__global__ void __launch_bounds__(128)
k256_sign_finish_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ priv, u32 canonical,
                        const u32* __restrict__ ws, u32* __restrict__ scratch, uint8_t* __restrict__ r,
                        uint8_t* __restrict__ s, uint8_t* __restrict__ recid, uint8_t* __restrict__ status) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t T = (size_t)gridDim.x * blockDim.x;
  k256_sign_finish_thread(tid, T, N, e, priv, canonical, ws, scratch, r, s, recid, status);
}
__global__ void __launch_bounds__(128)
k256_sign_slow_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ priv, u32 canonical,
                      const u32* __restrict__ gtab, uint8_t* __restrict__ r, uint8_t* __restrict__ s,
                      uint8_t* __restrict__ recid, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N || status[i] != ST_NEEDS_HOST) return;
  status[i] = k256_sign_item(i, e, priv, canonical, gtab, r, s, recid);
}

// Exact replay of the reference's own GLV/JSF/wNAF schedule for the items the fast kernel flagged
// (un-validated off-curve keys).  Divergent by nature; flagged items are rare.
__global__ void __launch_bounds__(128) k256_replay_tab_kernel(u32* tab) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 2 * REPLAY_NAF_PTS) rp_tab_entry(t, tab + 16 * t);
}
__global__ void __launch_bounds__(128)
k256_replay_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ r, const uint8_t* __restrict__ s,
                   const uint8_t* __restrict__ pub, const u32* __restrict__ tab, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N || status[i] != ST_NEEDS_HOST) return;
  status[i] = rp_verify_item(i, e, r, s, pub, tab);
}

// Point.mul / Point.mulAdd batches (short.js:422-441): k1*G + k2*P, k2*P, or k*G
__global__ void __launch_bounds__(128)
k256_prep_scalars_kernel(size_t N, const uint8_t* __restrict__ k1, const uint8_t* __restrict__ k2, u32* __restrict__ ws) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) prep_scalars_item(i, N, k1, k2, ws);
}
__global__ void __launch_bounds__(EB_VERIFY_BLOCK, EB_VERIFY_MINBLOCKS)
k256_mul_add_kernel(size_t N, const uint8_t* __restrict__ pts, const u32* __restrict__ ws, const u32* __restrict__ gtab,
                    u32* __restrict__ qtab, uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  status[i] = mul_add_item(i, N, pts, ws, gtab, qtab, out);
}
__global__ void __launch_bounds__(128)
k256_mul_add_replay_kernel(size_t N, const uint8_t* __restrict__ k1, const uint8_t* __restrict__ k2,
                           const uint8_t* __restrict__ pts, const u32* __restrict__ tab, uint8_t* __restrict__ out,
                           uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N || status[i] != ST_NEEDS_HOST) return;
  status[i] = rp_mul_add_item(i, k1, k2, pts, tab, out);
}
__global__ void __launch_bounds__(128)
k256_mul_g_kernel(size_t N, const uint8_t* __restrict__ k, const u32* __restrict__ gtab, uint8_t* __restrict__ out,
                  uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  status[i] = k256_mul_g_item(i, k, gtab, out);
}

__global__ void status_map_kernel(size_t N, uint8_t* __restrict__ status, uint8_t from, uint8_t to) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N && status[i] == from) status[i] = to;
}

// DER signatures -> fixed-width r, s (Signature._importDER, ec/signature.js:73-134).  `pre` carries the
// key-decoding verdict when there is one: a key that throws wins, as keyFromPublic runs first
// (ec/index.js:194-195).
__global__ void __launch_bounds__(128)
der_decode_kernel(size_t N, u32 len, const uint8_t* __restrict__ der, const unsigned long long* __restrict__ off,
                  uint8_t* __restrict__ r, uint8_t* __restrict__ s, uint8_t* __restrict__ pre, int pre_valid) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  bool ok = der_import(der + off[i], (size_t)(off[i + 1] - off[i]), len, r + (size_t)len * i, s + (size_t)len * i);
  uint8_t st = pre_valid ? pre[i] : 0;
  if (!st && !ok) st = ST_THROW_SIG_FORMAT;
  pre[i] = st;
}

// SEC1 decode (BaseCurve.decodePoint, lib/elliptic/curve/base.js:270-292; pointFromX short.js:187-204)
// fmt 1: 65-byte 04|06|07 || x || y ; fmt 2: 33-byte 02|03 || x.  Writes x||y (64 B) + a pre-status.
__global__ void __launch_bounds__(128) k256_decode_pub_kernel(size_t N, const uint8_t* __restrict__ in, u32 fmt,
                                                              uint8_t* __restrict__ xy, uint8_t* __restrict__ pre) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint8_t st = 0;
  if (fmt == EB200_PUB_SEC1_65) {
    const uint8_t* p = in + 65 * i;
    uint8_t tag = p[0];
    if (tag != 4 && tag != 6 && tag != 7) st = ST_THROW_POINT_FORMAT;
    else if ((tag == 6 && (p[64] & 1)) || (tag == 7 && !(p[64] & 1))) st = ST_THROW_ASSERT;   // base.js:278-281
    for (int k = 0; k < 64; k++) xy[64 * i + k] = p[1 + k];
  } else {
    const uint8_t* p = in + 33 * i;
    uint8_t tag = p[0];
    if (tag != 2 && tag != 3) st = ST_THROW_POINT_FORMAT;
    fe x = fe_from_be(p + 1);
    fe seven = fe_zero(); seven.v[0] = 7;
    fe y2 = fe_add(fe_mul(fe_sqr(x), x), seven);
    fe y = fe_sqrt_candidate(y2);
    if (!st && !fe_eq(fe_sqr(y), y2)) st = ST_THROW_INVALID_POINT;       // short.js:194-195
    y = fe_normalize(y);
    bool odd = tag == 3;
    if (((y.v[0] & 1) != 0) != odd) y = fe_normalize(fe_neg(y));
    x = fe_normalize(x);
    store_be<8>(xy + 64 * i, x.v);
    store_be<8>(xy + 64 * i + 32, y.v);
  }
  pre[i] = st;
}

__global__ void k256_selftest_fe_kernel(int op, size_t n, const u32* a, const u32* b, u32* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe A = load_fe(a + 8 * i), B = load_fe(b + 8 * i), R;
  switch (op) {
    case 0: R = fe_mul(A, B); break;
    case 1: R = fe_sqr(A); break;
    case 2: R = fe_add(A, B); break;
    case 3: R = fe_sub(A, B); break;
    case 4: R = fe_neg(A); break;
    case 5: R = fe_mul_small(A, b[8 * i]); break;
    case 6: R = fe_normalize(A); break;
    case 7: R = fe_inv(A); break;
    case 8: R = fe_sqrt_candidate(A); break;
    default: R = fe_zero();
  }
  store_fe(out + 8 * i, R);
}

// ---------------------------------------------------------------------------
// generic short-Weierstrass (a = -3) kernels
template <class C>
__global__ void __launch_bounds__(128) sw_gtab_kernel(u32* gtab) {
  typedef SW<C> W;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)W::GWINDOWS * W::GENTRIES) return;
  int j = (int)(t / W::GENTRIES), idx = (int)(t % W::GENTRIES);
  W::gtab_entry(j, idx, gtab + t * 2 * W::N);
}
template <class C>
__global__ void __launch_bounds__(128) sw_prep_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ r,
                                                      const uint8_t* __restrict__ s, u32* __restrict__ ws,
                                                      u32* __restrict__ scratch) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t T = (size_t)gridDim.x * blockDim.x;
  SW<C>::prep_thread(tid, T, N, e, r, s, ws, scratch);
}
template <class C>
__global__ void __launch_bounds__(128, 2)
sw_verify_kernel(size_t N, const uint8_t* __restrict__ pub, const uint8_t* __restrict__ r, const u32* __restrict__ ws,
                 const u32* __restrict__ gtab, u32* __restrict__ qtab, const uint8_t* __restrict__ pre,
                 uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (pre && pre[i]) { status[i] = pre[i]; return; }
  status[i] = SW<C>::verify_item(i, N, pub, r, ws, gtab, qtab);
}
// Exact replay of the reference's wNAF schedule for off-curve keys (ecdsa_sw_replay.cuh)
template <class C>
__global__ void __launch_bounds__(128) sw_replay_tab_kernel(u32* tab) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < SWReplay<C>::NAF_PTS) SWReplay<C>::tab_entry(t, tab + 2 * C::N * t);
}
template <class C>
__global__ void __launch_bounds__(128)
sw_replay_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ r, const uint8_t* __restrict__ s,
                 const uint8_t* __restrict__ pub, const u32* __restrict__ tab, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N || status[i] != ST_NEEDS_HOST) return;
  status[i] = SWReplay<C>::verify_item(i, e, r, s, pub, tab);
}
// EC.recoverPubKey on the non-GLV curves
template <class C>
__global__ void __launch_bounds__(128)
sw_prep_recover_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ r, const uint8_t* __restrict__ s,
                       u32* __restrict__ ws) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) SW<C>::prep_recover_item(i, N, e, r, s, ws);
}
template <class C>
__global__ void __launch_bounds__(128, 2)
sw_recover_kernel(size_t N, const uint8_t* __restrict__ r, const uint8_t* __restrict__ recid, const u32* __restrict__ ws,
                  const u32* __restrict__ gtab, u32* __restrict__ qtab, uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  status[i] = SW<C>::recover_item(i, N, r, recid, ws, gtab, qtab, out);
}

// EC.sign on p256 / p384 (ecdsa_sw_sign.cuh)
template <class SG>
__global__ void __launch_bounds__(128)
sw_sign_nonce_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ priv,
                     const u32* __restrict__ gtab, u32* __restrict__ ws, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) SG::nonce_item(i, N, e, priv, gtab, ws, status);
}
template <class SG>
__global__ void __launch_bounds__(128)
sw_sign_finish_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ priv, u32 canonical,
                      const u32* __restrict__ ws, u32* __restrict__ scratch, uint8_t* __restrict__ r,
                      uint8_t* __restrict__ s, uint8_t* __restrict__ recid, uint8_t* __restrict__ status) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t T = (size_t)gridDim.x * blockDim.x;
  SG::finish_thread(tid, T, N, e, priv, canonical, ws, scratch, r, s, recid, status);
}
template <class SG>
__global__ void __launch_bounds__(128)
sw_sign_slow_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ priv, u32 canonical,
                    const u32* __restrict__ gtab, uint8_t* __restrict__ r, uint8_t* __restrict__ s,
                    uint8_t* __restrict__ recid, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N || status[i] != ST_NEEDS_HOST) return;
  status[i] = SG::slow_item(i, e, priv, canonical, gtab, r, s, recid);
}

// Point.mul / mulAdd batches on the non-GLV short curves
template <class C>
__global__ void __launch_bounds__(128)
sw_prep_scalars_kernel(size_t N, const uint8_t* __restrict__ k1, const uint8_t* __restrict__ k2, u32* __restrict__ ws) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) SW<C>::prep_scalars_item(i, N, k1, k2, ws);
}
template <class C>
__global__ void __launch_bounds__(128, 2)
sw_mul_add_kernel(size_t N, const uint8_t* __restrict__ pts, const u32* __restrict__ ws, const u32* __restrict__ gtab,
                  u32* __restrict__ qtab, uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  status[i] = SW<C>::mul_add_item(i, N, pts, ws, gtab, qtab, out);
}
template <class C>
__global__ void __launch_bounds__(128)
sw_mul_add_replay_kernel(size_t N, const uint8_t* __restrict__ k1, const uint8_t* __restrict__ k2,
                         const uint8_t* __restrict__ pts, const u32* __restrict__ tab, uint8_t* __restrict__ out,
                         uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N || status[i] != ST_NEEDS_HOST) return;
  status[i] = SWReplay<C>::mul_add_item(i, k1, k2, pts, tab, out);
}
template <class C>
__global__ void __launch_bounds__(128)
sw_mul_g_kernel(size_t N, const uint8_t* __restrict__ k, const u32* __restrict__ gtab, uint8_t* __restrict__ out,
                uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  status[i] = SW<C>::mul_g_item(i, k, gtab, out);
}
template <class C>
__global__ void __launch_bounds__(128) sw_decode_pub_kernel(size_t N, const uint8_t* __restrict__ in, u32 fmt,
                                                            uint8_t* __restrict__ xy, uint8_t* __restrict__ pre) {
  typedef SW<C> W;
  typedef typename W::F F;
  constexpr int NL = W::N;
  constexpr size_t LEN = C::LEN;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint8_t st = 0;
  if (fmt == EB200_PUB_SEC1_65) {
    const uint8_t* p = in + (1 + 2 * LEN) * i;
    uint8_t tag = p[0];
    if (tag != 4 && tag != 6 && tag != 7) st = ST_THROW_POINT_FORMAT;
    else if ((tag == 6 && (p[2 * LEN] & 1)) || (tag == 7 && !(p[2 * LEN] & 1))) st = ST_THROW_ASSERT;
    for (size_t k = 0; k < 2 * LEN; k++) xy[2 * LEN * i + k] = p[1 + k];
  } else {
    const uint8_t* p = in + (1 + LEN) * i;
    uint8_t tag = p[0];
    if (tag != 2 && tag != 3) st = ST_THROW_POINT_FORMAT;
    typename F::fe t;
    W::ldb(t.v, p + 1);
    typename F::fe x = F::to_mont(t);
    typename F::fe y2 = F::add(F::sub(F::mul(F::sqr(x), x), F::add(F::dbl(x), x)), C::b());
    typename F::fe y = F::zero();
    uint8_t ss = W::sqrt_ref(y2, &y);                 // Red.sqrt: a^((p+1)/4), or Tonelli-Shanks on p224
    if (!st && ss) st = ss;
    if (!st && !F::eq(F::sqr(y), y2)) st = ST_THROW_INVALID_POINT;
    typename F::fe yp = F::from_mont(y);
    bool odd = tag == 3;
    if (((yp.v[0] & 1) != 0) != odd) yp = F::from_mont(F::neg(y));
    typename F::fe xp = F::from_mont(x);
    W::stb(xy + 2 * LEN * i, xp.v);
    W::stb(xy + 2 * LEN * i + LEN, yp.v);
  }
  pre[i] = st;
}
template <class C>
__global__ void sw_selftest_fe_kernel(int op, size_t n, const u32* a, const u32* b, u32* out) {
  typedef typename SW<C>::F F;
  constexpr int NL = SW<C>::N;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename F::fe A = F::to_mont(load_fe_n<NL>(a + NL * i)), B = F::to_mont(load_fe_n<NL>(b + NL * i)), R;
  switch (op) {
    case 0: R = F::mul(A, B); break;
    case 1: R = F::sqr(A); break;
    case 2: R = F::add(A, B); break;
    case 3: R = F::sub(A, B); break;
    case 4: R = F::neg(A); break;
    case 7: R = F::inv(A); break;
    default: R = A;
  }
  store_fe_n<NL>(out + NL * i, F::from_mont(R));
}

// ---------------------------------------------------------------------------
// ed25519 / curve25519 kernels
__global__ void __launch_bounds__(128) ed_gtab_kernel(u32* gtab) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)ED_GWINDOWS * ED_GENTRIES) return;
  ed_gtab_entry((int)(t / ED_GENTRIES), (int)(t % ED_GENTRIES), gtab + t * 24);
}
__global__ void __launch_bounds__(128, 3)
ed25519_verify_kernel(size_t N, const uint8_t* __restrict__ R, const uint8_t* __restrict__ S,
                      const uint8_t* __restrict__ A, const uint8_t* __restrict__ h,
                      const u32* __restrict__ gtab, u32* __restrict__ atab, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  status[i] = ed25519_verify_item(i, R, S, A, h, gtab, atab);
}
__global__ void f25_selftest_kernel(int op, size_t n, const u32* a, const u32* b, u32* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  f25 A = f25_load(a + 8 * i), B = f25_load(b + 8 * i), R;
  switch (op) {
    case 0: R = f25_mul(A, B); break;
    case 1: R = f25_sqr(A); break;
    case 2: R = f25_add(A, B); break;
    case 3: R = f25_sub(A, B); break;
    case 4: R = f25_neg(A); break;
    case 5: R = f25_mul_small(A, b[8 * i]); break;
    case 6: R = f25_normalize(A); break;
    case 7: R = f25_inv(A); break;
    case 8: R = f25_pow_p58(A); break;
    default: R = f25_zero();
  }
  f25_store(out + 8 * i, R);
}
__global__ void __launch_bounds__(128)
ed25519_hash_kernel(size_t N, const uint8_t* __restrict__ R, const uint8_t* __restrict__ A,
                    const uint8_t* __restrict__ msgs, const u64* __restrict__ msg_off, uint8_t* __restrict__ h) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  ed25519_hash_item(i, R, A, msgs, msg_off, h);
}
__global__ void __launch_bounds__(128, 4)
x25519_derive_kernel(size_t N, const uint8_t* __restrict__ priv, const uint8_t* __restrict__ pubx,
                     uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  status[i] = x25519_derive_item(i, priv, pubx, out);
}

// ---------------------------------------------------------------------------
// context
namespace {
constexpr int MAX_CHUNKS = 16;
struct Ctx {
  bool ready = false;
  int device = -1;
  cudaStream_t stream = nullptr, stream2 = nullptr, copy_stream = nullptr;
  u32* gtab[16] = {};
  u32* sw_replay_tab[16] = {};         // p256/p384: the reference's wnd-8 NAF table of G
  u32* replay_tab = nullptr;          // secp256k1: the reference's wnd-7 NAF table of G and its beta image
  uint8_t* d_in = nullptr; size_t d_in_cap = 0;
  uint8_t* d_ws = nullptr; size_t d_ws_cap = 0;
  uint8_t* d_status = nullptr; size_t d_status_cap = 0;
  cudaEvent_t ev[6] = {};
  cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_k0[MAX_CHUNKS] = {}, ev_k1[MAX_CHUNKS] = {}, ev_done[MAX_CHUNKS] = {};
  eb200_timing timing = {};
  bool dev_timing_pending = false;
};
Ctx g;
std::mutex g_mu;
thread_local char g_err[256] = "";

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? EB200_ERR_NO_DEVICE : EB200_ERR_CUDA;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

int grow(uint8_t** p, size_t* cap, size_t need) {
  if (*cap >= need) return EB200_OK;
  if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
  CK(cudaMalloc(p, need));
  *cap = need;
  return EB200_OK;
}
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t fe_len(int curve) {   // field-element bytes for the selftest hooks (also the 25519 curves)
  if (curve == EB200_CURVE_P192) return 24;
  return (curve == EB200_CURVE_P384) ? 48 : (curve == EB200_CURVE_P521) ? 72 :     // 18 limbs; p224 = 8 limbs
         ((curve >= EB200_CURVE_SECP256K1 && curve <= EB200_CURVE_CURVE25519) || curve == EB200_CURVE_P224) ? 32 : 0;
}
size_t curve_len(int curve) {
  switch (curve) {
    case EB200_CURVE_SECP256K1: case EB200_CURVE_P256: return 32;
    case EB200_CURVE_P384: return 48;
    case EB200_CURVE_P521: return 66;
    case EB200_CURVE_P192: return 24;
    case EB200_CURVE_P224: return 28;
    default: return 0;
  }
}
size_t pub_item_bytes(size_t len, u32 fmt) {
  return fmt == EB200_PUB_XY ? 2 * len : fmt == EB200_PUB_SEC1_65 ? 1 + 2 * len : fmt == EB200_PUB_SEC1_33 ? 1 + len : 0;
}

// workspace: [ws words | scratch words | qtab words | decoded xy | pre-status]
struct WsLayout { size_t ws, scratch, qtab, xy, pre, total; };
WsLayout ws_layout(int curve, size_t n) {
  size_t prep_words, scratch_words, qtab_words, len = curve_len(curve);
  if (curve == EB200_CURVE_SECP256K1) { prep_words = PREP_WORDS; scratch_words = 8; qtab_words = QTAB_WORDS; }
  else {
#define EB_WS(C) (prep_words = SW<C>::PREP_WORDS, scratch_words = SW<C>::N, qtab_words = SW<C>::QTAB_WORDS, 0)
    (void)SW_DISPATCH(curve, EB_WS);
#undef EB_WS
  }
  WsLayout L;
  L.ws = 0;
  L.scratch = align256(L.ws + prep_words * n * 4);
  L.qtab = align256(L.scratch + scratch_words * n * 4);
  L.xy = align256(L.qtab + qtab_words * n * 4);
  L.pre = align256(L.xy + 2 * len * n);
  L.total = align256(L.pre + n);
  return L;
}

template <class C>
int sw_ensure_table(int curve) {
  typedef SW<C> W;
  if (g.gtab[curve]) return EB200_OK;
  size_t entries = (size_t)W::GWINDOWS * W::GENTRIES;
  CK(cudaMalloc(&g.gtab[curve], entries * 2 * W::N * 4));
  sw_gtab_kernel<C><<<(unsigned)((entries + 127) / 128), 128, 0, g.stream>>>(g.gtab[curve]);
  CK(cudaGetLastError());
  CK(cudaMalloc(&g.sw_replay_tab[curve], (size_t)SWReplay<C>::TAB_WORDS * 4));
  sw_replay_tab_kernel<C><<<1, 128, 0, g.stream>>>(g.sw_replay_tab[curve]);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g.stream));
  return EB200_OK;
}
int ensure_table(int curve) {
  if (curve == EB200_CURVE_SECP256K1) {
    if (g.gtab[curve]) return EB200_OK;
    size_t entries = (size_t)GTAB_WINDOWS * GTAB_ENTRIES;
    CK(cudaMalloc(&g.gtab[curve], entries * 16 * 4));
    k256_gtab_kernel<<<(unsigned)((entries + 127) / 128), 128, 0, g.stream>>>(g.gtab[curve]);
    CK(cudaGetLastError());
    CK(cudaMalloc(&g.replay_tab, (size_t)REPLAY_TAB_WORDS * 4));
    k256_replay_tab_kernel<<<2, 128, 0, g.stream>>>(g.replay_tab);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(g.stream));
    return EB200_OK;
  }
  if (curve == EB200_CURVE_P256 || curve == EB200_CURVE_P384 || curve == EB200_CURVE_P521 || curve == EB200_CURVE_P192 ||
      curve == EB200_CURVE_P224) {
#define EB_ENS(C) sw_ensure_table<C>(curve)
    return SW_DISPATCH(curve, EB_ENS);
#undef EB_ENS
  }
  if (curve == EB200_CURVE_ED25519) {
    if (g.gtab[curve]) return EB200_OK;
    size_t entries = (size_t)ED_GWINDOWS * ED_GENTRIES;
    CK(cudaMalloc(&g.gtab[curve], entries * 24 * 4));
    ed_gtab_kernel<<<(unsigned)((entries + 127) / 128), 128, 0, g.stream>>>(g.gtab[curve]);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(g.stream));
    return EB200_OK;
  }
  return EB200_ERR_UNSUPPORTED;
}

template <class C>
int sw_launch_verify(int curve, size_t n, const uint8_t* d_e, const uint8_t* d_r, const uint8_t* d_s, const uint8_t* xy,
                     const uint8_t* pre, u32* ws, u32* scratch, u32* qtab, uint8_t* d_status, cudaStream_t st,
                     unsigned pb, unsigned nb, cudaEvent_t ev_main0, cudaEvent_t* ev_main1) {
  sw_prep_kernel<C><<<pb, 128, 0, st>>>(n, d_e, d_r, d_s, ws, scratch);
  CK(cudaGetLastError());
  if (ev_main0) CK(cudaEventRecord(ev_main0, st));
  sw_verify_kernel<C><<<nb, 128, 0, st>>>(n, xy, d_r, ws, g.gtab[curve], qtab, pre, d_status);
  CK(cudaGetLastError());
  if (*ev_main1) { CK(cudaEventRecord(*ev_main1, st)); *ev_main1 = nullptr; }
  sw_replay_kernel<C><<<nb, 128, 0, st>>>(n, d_e, d_r, d_s, xy, g.sw_replay_tab[curve], d_status);
  return EB200_OK;
}

// Launches decode (if needed) + prep + verify for n items on stream st.  All pointers are device pointers.
int launch_verify(int curve, size_t n, const uint8_t* d_e, const uint8_t* d_r, const uint8_t* d_s,
                  const uint8_t* d_pub, u32 pub_fmt, uint8_t* d_status, uint8_t* d_workspace,
                  cudaStream_t st, cudaEvent_t ev_main0, cudaEvent_t ev_main1, unsigned* launches,
                  const uint8_t* d_der = nullptr, const unsigned long long* d_der_off = nullptr) {
  if (n == 0) return EB200_OK;
  WsLayout L = ws_layout(curve, n);
  u32* ws = (u32*)(d_workspace + L.ws);
  u32* scratch = (u32*)(d_workspace + L.scratch);
  u32* qtab = (u32*)(d_workspace + L.qtab);
  const uint8_t* xy = d_pub;
  const uint8_t* pre = nullptr;
  unsigned nb = (unsigned)((n + 127) / 128);
  size_t T = (n + PREP_BATCH - 1) / PREP_BATCH;
  unsigned pb = (unsigned)((T + 127) / 128);
  unsigned cnt = 0;
  if (pub_fmt != EB200_PUB_XY) {
    uint8_t* dxy = d_workspace + L.xy;
    uint8_t* dpre = d_workspace + L.pre;
    if (curve == EB200_CURVE_SECP256K1) k256_decode_pub_kernel<<<nb, 128, 0, st>>>(n, d_pub, pub_fmt, dxy, dpre);
    else {
#define EB_DEC(C) (sw_decode_pub_kernel<C><<<nb, 128, 0, st>>>(n, d_pub, pub_fmt, dxy, dpre), 0)
      (void)SW_DISPATCH(curve, EB_DEC);
#undef EB_DEC
    }
    CK(cudaGetLastError());
    xy = dxy; pre = dpre; cnt++;
  }
  if (d_der) {     // d_r / d_s are then scratch the decoder fills
    uint8_t* dpre = d_workspace + L.pre;
    der_decode_kernel<<<nb, 128, 0, st>>>(n, (u32)curve_len(curve), d_der, d_der_off, (uint8_t*)d_r, (uint8_t*)d_s, dpre, pre != nullptr);
    CK(cudaGetLastError());
    pre = dpre; cnt++;
  }
  if (curve == EB200_CURVE_SECP256K1) {
    k256_prep_kernel<<<pb, 128, 0, st>>>(n, d_e, d_r, d_s, ws, scratch);
    CK(cudaGetLastError());
    if (ev_main0) CK(cudaEventRecord(ev_main0, st));
    unsigned vb = (unsigned)((n + EB_VERIFY_BLOCK - 1) / EB_VERIFY_BLOCK);
    k256_verify_kernel<<<vb, EB_VERIFY_BLOCK, 0, st>>>(n, xy, d_r, ws, g.gtab[curve], qtab, pre, d_status);
    CK(cudaGetLastError());
    if (ev_main1) { CK(cudaEventRecord(ev_main1, st)); ev_main1 = nullptr; }
    k256_replay_kernel<<<nb, 128, 0, st>>>(n, d_e, d_r, d_s, xy, g.replay_tab, d_status);
    cnt++;
  } else {
#define EB_VER(C) sw_launch_verify<C>(curve, n, d_e, d_r, d_s, xy, pre, ws, scratch, qtab, d_status, st, pb, nb, ev_main0, &ev_main1)
    int rc = SW_DISPATCH(curve, EB_VER);
#undef EB_VER
    if (rc) return rc;
    cnt++;
  }
  CK(cudaGetLastError());
  if (ev_main1) CK(cudaEventRecord(ev_main1, st));
  cnt += 2;
  if (launches) *launches += cnt;
  return EB200_OK;
}

bool curve_ok(int curve) { return curve_len(curve) != 0; }
bool fmt_ok(u32 fmt) { return fmt == EB200_PUB_XY || fmt == EB200_PUB_SEC1_65 || fmt == EB200_PUB_SEC1_33; }
}  // namespace

extern "C" {

const char* eb200_strerror(int code) {
  switch (code) {
    case EB200_OK: return "ok";
    case EB200_ERR_NO_DEVICE: return "no CUDA device available (this library has no CPU fallback)";
    case EB200_ERR_CUDA: return "CUDA error (see eb200_last_error)";
    case EB200_ERR_ARG: return "invalid argument";
    case EB200_ERR_NOT_INIT: return "eb200_init has not been called";
    case EB200_ERR_UNSUPPORTED: return "curve or format not supported by this build";
    default: return "unknown error";
  }
}

const char* eb200_last_error(void) { return g_err; }

int eb200_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g.ready && g.device == device) return EB200_OK;
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess) { cuda_fail(e, "cudaGetDeviceCount"); return EB200_ERR_NO_DEVICE; }
  if (cnt == 0) { snprintf(g_err, sizeof g_err, "no CUDA devices"); return EB200_ERR_NO_DEVICE; }
  if (device < 0 || device >= cnt) return EB200_ERR_ARG;
  CK(cudaSetDevice(device));
  if (!g.stream) CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  if (!g.copy_stream) CK(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
  if (!g.stream2) CK(cudaStreamCreateWithFlags(&g.stream2, cudaStreamNonBlocking));
  for (int i = 0; i < 6; i++) if (!g.ev[i]) CK(cudaEventCreate(&g.ev[i]));
  for (int i = 0; i < MAX_CHUNKS; i++) {
    if (!g.ev_in[i]) CK(cudaEventCreate(&g.ev_in[i]));
    if (!g.ev_k0[i]) CK(cudaEventCreate(&g.ev_k0[i]));
    if (!g.ev_k1[i]) CK(cudaEventCreate(&g.ev_k1[i]));
    if (!g.ev_done[i]) CK(cudaEventCreate(&g.ev_done[i]));
  }
  for (int c = 0; c < 16; c++) if (g.gtab[c]) { cudaFree(g.gtab[c]); g.gtab[c] = nullptr; }
  for (int c = 0; c < 16; c++) if (g.sw_replay_tab[c]) { cudaFree(g.sw_replay_tab[c]); g.sw_replay_tab[c] = nullptr; }
  if (g.replay_tab) { cudaFree(g.replay_tab); g.replay_tab = nullptr; }
  g.device = device;
  g.ready = true;
  // the headline curve's table is built eagerly; the others on first use
  int rc = ensure_table(EB200_CURVE_SECP256K1);
  if (rc) { g.ready = false; return rc; }
  return EB200_OK;
}

int eb200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g.ready) return EB200_OK;
  cudaSetDevice(g.device);
  for (int c = 0; c < 16; c++) if (g.gtab[c]) { cudaFree(g.gtab[c]); g.gtab[c] = nullptr; }
  for (int c = 0; c < 16; c++) if (g.sw_replay_tab[c]) { cudaFree(g.sw_replay_tab[c]); g.sw_replay_tab[c] = nullptr; }
  if (g.replay_tab) { cudaFree(g.replay_tab); g.replay_tab = nullptr; }
  cudaFree(g.d_in); g.d_in = nullptr; g.d_in_cap = 0;
  cudaFree(g.d_ws); g.d_ws = nullptr; g.d_ws_cap = 0;
  cudaFree(g.d_status); g.d_status = nullptr; g.d_status_cap = 0;
  for (int i = 0; i < 6; i++) if (g.ev[i]) { cudaEventDestroy(g.ev[i]); g.ev[i] = nullptr; }
  for (int i = 0; i < MAX_CHUNKS; i++) {
    if (g.ev_in[i]) { cudaEventDestroy(g.ev_in[i]); g.ev_in[i] = nullptr; }
    if (g.ev_k0[i]) { cudaEventDestroy(g.ev_k0[i]); g.ev_k0[i] = nullptr; }
    if (g.ev_k1[i]) { cudaEventDestroy(g.ev_k1[i]); g.ev_k1[i] = nullptr; }
    if (g.ev_done[i]) { cudaEventDestroy(g.ev_done[i]); g.ev_done[i] = nullptr; }
  }
  if (g.stream) { cudaStreamDestroy(g.stream); g.stream = nullptr; }
  if (g.copy_stream) { cudaStreamDestroy(g.copy_stream); g.copy_stream = nullptr; }
  if (g.stream2) { cudaStreamDestroy(g.stream2); g.stream2 = nullptr; }
  g.ready = false;
  return EB200_OK;
}

int eb200_last_timing(eb200_timing* out) {
  if (!out) return EB200_ERR_ARG;
  if (g.dev_timing_pending) {
    // device-pointer call: the caller has synchronised its stream by now
    unsigned l = g.timing.launches;
    g.timing = eb200_timing{};
    if (cudaEventElapsedTime(&g.timing.kernel_ms, g.ev[1], g.ev[2]) != cudaSuccess) return EB200_ERR_CUDA;
    if (cudaEventElapsedTime(&g.timing.main_kernel_ms, g.ev[4], g.ev[5]) != cudaSuccess) return EB200_ERR_CUDA;
    g.timing.launches = l;
    g.dev_timing_pending = false;
  }
  *out = g.timing;
  return EB200_OK;
}

size_t eb200_ecdsa_verify_workspace_bytes(int curve, size_t n) {
  if (!curve_ok(curve)) return 0;
  return ws_layout(curve, n).total;
}

int eb200_ecdsa_verify_batch_dev(int curve, size_t n, const uint8_t* d_e, const uint8_t* d_r,
                                 const uint8_t* d_s, const uint8_t* d_pub, uint32_t pub_fmt,
                                 uint8_t* d_status, void* d_workspace, void* stream) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve) || !fmt_ok(pub_fmt)) return EB200_ERR_UNSUPPORTED;
  if (n && (!d_e || !d_r || !d_s || !d_pub || !d_status || !d_workspace)) return EB200_ERR_ARG;
  int rc = ensure_table(curve);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;   // NULL is the CUDA default stream, as everywhere in CUDA
  // events on the caller's stream: eb200_last_timing() reports them once the stream has been synchronised
  CK(cudaEventRecord(g.ev[1], st));
  unsigned launches = 0;
  rc = launch_verify(curve, n, d_e, d_r, d_s, d_pub, pub_fmt, d_status, (uint8_t*)d_workspace, st, g.ev[4], g.ev[5], &launches);
  if (rc) return rc;
  CK(cudaEventRecord(g.ev[2], st));
  g.timing.launches = launches;
  g.dev_timing_pending = true;
  return EB200_OK;
}

// Host-pointer call.  Large batches are cut into chunks: chunk k+1 is copied host->device on a copy
// stream while chunk k is being verified, and results stream back as each chunk finishes.
int eb200_ecdsa_verify_batch(int curve, size_t n, const uint8_t* e, const uint8_t* r,
                             const uint8_t* s, const uint8_t* pub, uint32_t pub_fmt,
                             uint8_t* status) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve) || !fmt_ok(pub_fmt)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!e || !r || !s || !pub || !status) return EB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  int rc = ensure_table(curve);
  if (rc) return rc;
  const size_t len = curve_len(curve), pb = pub_item_bytes(len, pub_fmt);
  int chunks = 1;
  if (n >= ((size_t)1 << 18)) chunks = 4;     // 2^18-item chunks keep the grid tail small (r01: 8 chunks cost 10%)
  if (n >= ((size_t)1 << 22)) chunks = MAX_CHUNKS;
  if (const char* ev = getenv("EB200_CHUNKS")) { int c = atoi(ev); if (c >= 1 && c <= MAX_CHUNKS) chunks = c; }   // tuning knob
  size_t per = (n + chunks - 1) / chunks;
  per = (per + 127) & ~(size_t)127;
  size_t item_in = 3 * len + pb;
  if ((rc = grow(&g.d_in, &g.d_in_cap, align256(n * item_in) + 1024))) return rc;
  // two chunks in flight (alternating compute streams, so the grid tail of chunk k is filled by chunk k+1)
  const size_t ws_slot = ws_layout(curve, per).total;
  if ((rc = grow(&g.d_ws, &g.d_ws_cap, (chunks > 1 ? 2 : 1) * ws_slot))) return rc;
  if ((rc = grow(&g.d_status, &g.d_status_cap, n))) return rc;
  uint8_t* d_e = g.d_in;
  uint8_t* d_r = d_e + n * len;
  uint8_t* d_s = d_r + n * len;
  uint8_t* d_pub = d_s + n * len;
  cudaStream_t cs = g.copy_stream;
  unsigned launches = 0;
  CK(cudaEventRecord(g.ev[0], cs));
  int used = 0;
  for (int k = 0; k < chunks; k++) {
    size_t lo = (size_t)k * per;
    if (lo >= n) break;
    size_t m = (lo + per <= n) ? per : n - lo;
    used = k + 1;
    CK(cudaMemcpyAsync(d_e + lo * len, e + lo * len, m * len, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(d_r + lo * len, r + lo * len, m * len, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(d_s + lo * len, s + lo * len, m * len, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(d_pub + lo * pb, pub + lo * pb, m * pb, cudaMemcpyHostToDevice, cs));
    CK(cudaEventRecord(g.ev_in[k], cs));
    cudaStream_t ks = (k & 1) ? g.stream2 : g.stream;
    CK(cudaStreamWaitEvent(ks, g.ev_in[k], 0));
    if ((rc = launch_verify(curve, m, d_e + lo * len, d_r + lo * len, d_s + lo * len, d_pub + lo * pb, pub_fmt,
                            g.d_status + lo, g.d_ws + (size_t)(k & 1) * ws_slot, ks, g.ev_k0[k], g.ev_k1[k], &launches))) return rc;
    CK(cudaEventRecord(g.ev_done[k], ks));
  }
  // results: one device->host copy per chunk, behind that chunk's kernels, on the copy stream
  for (int k = 0; k < used; k++) {
    size_t lo = (size_t)k * per;
    size_t m = (lo + per <= n) ? per : n - lo;
    CK(cudaStreamWaitEvent(cs, g.ev_done[k], 0));
    CK(cudaMemcpyAsync(status + lo, g.d_status + lo, m, cudaMemcpyDeviceToHost, cs));
  }
  CK(cudaEventRecord(g.ev[3], cs));
  CK(cudaStreamSynchronize(cs));
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaStreamSynchronize(g.stream2));
  g.dev_timing_pending = false;
  g.timing = eb200_timing{};
  float total = 0, t = 0;
  cudaEventElapsedTime(&total, g.ev[0], g.ev[3]);
  cudaEventElapsedTime(&g.timing.h2d_ms, g.ev[0], g.ev_in[used - 1]);       // all inputs resident
  for (int k = 0; k < used; k++) {       // chunks overlap on two streams: report the span of the main kernels
    cudaEventElapsedTime(&t, g.ev_k0[0], g.ev_k1[k]);
    if (t > g.timing.main_kernel_ms) g.timing.main_kernel_ms = t;
  }
  cudaEventElapsedTime(&t, g.ev_done[used - 1], g.ev[3]);
  g.timing.d2h_ms = t;                                                       // exposed tail copy
  g.timing.kernel_ms = total;                                                // whole call on the GPU timeline
  g.timing.launches = launches;
  return EB200_OK;
}

// DER-encoded signatures, parsed on the GPU (variable length: concatenated bytes + n+1 offsets)
int eb200_ecdsa_verify_batch_der(int curve, size_t n, const uint8_t* e, const uint8_t* sigs, const uint64_t* sig_off,
                                 const uint8_t* pub, uint32_t pub_fmt, uint8_t* status) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve) || !fmt_ok(pub_fmt)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!e || !sigs || !sig_off || !pub || !status) return EB200_ERR_ARG;
  for (size_t i = 0; i < n; i++) if (sig_off[i + 1] < sig_off[i]) return EB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  int rc = ensure_table(curve);
  if (rc) return rc;
  const size_t len = curve_len(curve), pb = pub_item_bytes(len, pub_fmt);
  const size_t sig_bytes = (size_t)(sig_off[n] - sig_off[0]);
  const size_t off_bytes = align256((n + 1) * 8);
  if ((rc = grow(&g.d_in, &g.d_in_cap, off_bytes + align256(n * (3 * len + pb)) + align256(sig_bytes) + 1024))) return rc;
  if ((rc = grow(&g.d_ws, &g.d_ws_cap, ws_layout(curve, n).total))) return rc;
  if ((rc = grow(&g.d_status, &g.d_status_cap, n))) return rc;
  unsigned long long* d_off = (unsigned long long*)g.d_in;
  uint8_t* d_e = g.d_in + off_bytes;
  uint8_t* d_r = d_e + n * len;
  uint8_t* d_s = d_r + n * len;
  uint8_t* d_pub = d_s + n * len;
  uint8_t* d_sig = d_pub + align256(n * pb);
  cudaStream_t st = g.stream;
  unsigned launches = 0;
  CK(cudaEventRecord(g.ev[0], st));
  CK(cudaMemcpyAsync(d_off, sig_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_e, e, n * len, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_pub, pub, n * pb, cudaMemcpyHostToDevice, st));
  if (sig_bytes) CK(cudaMemcpyAsync(d_sig, sigs + sig_off[0], sig_bytes, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(g.ev[1], st));
  // offsets are used relative to sig_off[0] on the device
  if ((rc = launch_verify(curve, n, d_e, d_r, d_s, d_pub, pub_fmt, g.d_status, g.d_ws, st, g.ev[4], g.ev[5], &launches,
                          d_sig - sig_off[0], d_off))) return rc;
  CK(cudaEventRecord(g.ev[2], st));
  CK(cudaMemcpyAsync(status, g.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(g.ev[3], st));
  CK(cudaStreamSynchronize(st));
  g.dev_timing_pending = false;
  g.timing = eb200_timing{};
  cudaEventElapsedTime(&g.timing.h2d_ms, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&g.timing.kernel_ms, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&g.timing.d2h_ms, g.ev[2], g.ev[3]);
  cudaEventElapsedTime(&g.timing.main_kernel_ms, g.ev[4], g.ev[5]);
  g.timing.launches = launches;
  return EB200_OK;
}

// ---- ECDSA public-key recovery (secp256k1) ------------------------------------------------
}  // extern "C"

template <class C>
static int sw_recover_launch(int curve, size_t n, const uint8_t* d_e, const uint8_t* d_r, const uint8_t* d_s, const uint8_t* d_id,
                             uint8_t* d_out, const WsLayout& L, cudaStream_t st) {
  unsigned nb = (unsigned)((n + 127) / 128);
  sw_prep_recover_kernel<C><<<nb, 128, 0, st>>>(n, d_e, d_r, d_s, (u32*)(g.d_ws + L.ws));
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[4], st));
  sw_recover_kernel<C><<<nb, 128, 0, st>>>(n, d_r, d_id, (u32*)(g.d_ws + L.ws), g.gtab[curve], (u32*)(g.d_ws + L.qtab), d_out, g.d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[5], st));
  return EB200_OK;
}

extern "C" {

int eb200_ecdsa_recover_batch(int curve, size_t n, const uint8_t* e, const uint8_t* r, const uint8_t* s,
                              const uint8_t* recid, uint8_t* out_xy, uint8_t* status) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!e || !r || !s || !recid || !out_xy || !status) return EB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  int rc = ensure_table(curve);
  if (rc) return rc;
  const size_t len = curve_len(curve);
  WsLayout L = ws_layout(curve, n);
  if ((rc = grow(&g.d_in, &g.d_in_cap, n * (5 * len + 1) + 256))) return rc;
  if ((rc = grow(&g.d_ws, &g.d_ws_cap, L.total))) return rc;
  if ((rc = grow(&g.d_status, &g.d_status_cap, n))) return rc;
  uint8_t *d_e = g.d_in, *d_r = d_e + len * n, *d_s = d_r + len * n, *d_out = d_s + len * n, *d_id = d_out + 2 * len * n;
  cudaStream_t st = g.stream;
  CK(cudaEventRecord(g.ev[0], st));
  CK(cudaMemcpyAsync(d_e, e, len * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_r, r, len * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_s, s, len * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_id, recid, n, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(g.ev[1], st));
  if (curve == EB200_CURVE_SECP256K1) {
    size_t T = (n + PREP_BATCH - 1) / PREP_BATCH;
    k256_prep_recover_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(n, d_e, d_r, d_s, (u32*)(g.d_ws + L.ws), (u32*)(g.d_ws + L.scratch));
    CK(cudaGetLastError());
    CK(cudaEventRecord(g.ev[4], st));
    k256_recover_kernel<<<(unsigned)((n + EB_VERIFY_BLOCK - 1) / EB_VERIFY_BLOCK), EB_VERIFY_BLOCK, 0, st>>>(
        n, d_r, d_id, (u32*)(g.d_ws + L.ws), g.gtab[curve], (u32*)(g.d_ws + L.qtab), d_out, g.d_status);
    CK(cudaGetLastError());
    CK(cudaEventRecord(g.ev[5], st));
  } else {
#define EB_REC(C) sw_recover_launch<C>(curve, n, d_e, d_r, d_s, d_id, d_out, L, st)
    rc = SW_DISPATCH(curve, EB_REC);
#undef EB_REC
    if (rc) return rc;
  }
  CK(cudaEventRecord(g.ev[2], st));
  CK(cudaMemcpyAsync(out_xy, d_out, 2 * len * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(status, g.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(g.ev[3], st));
  CK(cudaStreamSynchronize(st));
  g.dev_timing_pending = false;
  g.timing = eb200_timing{};
  cudaEventElapsedTime(&g.timing.h2d_ms, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&g.timing.kernel_ms, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&g.timing.d2h_ms, g.ev[2], g.ev[3]);
  cudaEventElapsedTime(&g.timing.main_kernel_ms, g.ev[4], g.ev[5]);
  g.timing.launches = 2;
  return EB200_OK;
}

// ---- Point.mul / Point.mulAdd batches (secp256k1) -----------------------------------------------------
// k1 == NULL: k2*P;  pts == NULL: k2*G;  both given: k1*G + k2*P.
}  // extern "C"

template <class C>
static int sw_mul_add_launch(int curve, size_t n, const uint8_t* d_k1, const uint8_t* d_k2, const uint8_t* d_pts,
                             uint8_t* d_out, const WsLayout& L, cudaStream_t st, unsigned* launches, bool derive) {
  unsigned nb = (unsigned)((n + 127) / 128);
  if (!d_pts) {
    CK(cudaEventRecord(g.ev[4], st));
    sw_mul_g_kernel<C><<<nb, 128, 0, st>>>(n, d_k2, g.gtab[curve], d_out, g.d_status);
    CK(cudaGetLastError());
    CK(cudaEventRecord(g.ev[5], st));
    *launches = 1;
    return EB200_OK;
  }
  sw_prep_scalars_kernel<C><<<nb, 128, 0, st>>>(n, d_k1, d_k2, (u32*)(g.d_ws + L.ws));
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[4], st));
  sw_mul_add_kernel<C><<<nb, 128, 0, st>>>(n, d_pts, (u32*)(g.d_ws + L.ws), g.gtab[curve], (u32*)(g.d_ws + L.qtab), d_out, g.d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[5], st));
  if (derive) status_map_kernel<<<nb, 128, 0, st>>>(n, g.d_status, ST_NEEDS_HOST, ST_THROW_NOT_VALIDATED);
  else sw_mul_add_replay_kernel<C><<<nb, 128, 0, st>>>(n, d_k1, d_k2, d_pts, g.sw_replay_tab[curve], d_out, g.d_status);
  CK(cudaGetLastError());
  *launches = 3;
  return EB200_OK;
}

// derive: KeyPair.derive (ec/key.js:102-107) -- an off-curve point is the reference's
// 'public point not validated' throw instead of a replayed multiplication, and only x is returned.
static int mul_add_common(int curve, size_t n, const uint8_t* k1, const uint8_t* k2, const uint8_t* pts,
                          uint8_t* out_xy, uint8_t* status, bool derive = false) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!k2 || !out_xy || !status || (k1 && !pts)) return EB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  int rc = ensure_table(curve);
  if (rc) return rc;
  const size_t len = curve_len(curve);
  WsLayout L = ws_layout(curve, n);
  if ((rc = grow(&g.d_in, &g.d_in_cap, n * 6 * len + 256))) return rc;
  if ((rc = grow(&g.d_ws, &g.d_ws_cap, L.total))) return rc;
  if ((rc = grow(&g.d_status, &g.d_status_cap, n))) return rc;
  uint8_t *d_k1 = g.d_in, *d_k2 = d_k1 + len * n, *d_pts = d_k2 + len * n, *d_out = d_pts + 2 * len * n;
  cudaStream_t st = g.stream;
  unsigned nb = (unsigned)((n + 127) / 128), launches = 0;
  CK(cudaEventRecord(g.ev[0], st));
  if (k1) CK(cudaMemcpyAsync(d_k1, k1, len * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_k2, k2, len * n, cudaMemcpyHostToDevice, st));
  if (pts) CK(cudaMemcpyAsync(d_pts, pts, 2 * len * n, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(g.ev[1], st));
  if (curve != EB200_CURVE_SECP256K1) {
#define EB_MA(C) sw_mul_add_launch<C>(curve, n, k1 ? d_k1 : nullptr, d_k2, pts ? d_pts : nullptr, d_out, L, st, &launches, derive)
    if ((rc = SW_DISPATCH(curve, EB_MA))) return rc;
#undef EB_MA
  } else if (!pts) {
    CK(cudaEventRecord(g.ev[4], st));
    k256_mul_g_kernel<<<nb, 128, 0, st>>>(n, d_k2, g.gtab[curve], d_out, g.d_status);
    CK(cudaGetLastError());
    CK(cudaEventRecord(g.ev[5], st));
    launches = 1;
  } else {
    k256_prep_scalars_kernel<<<nb, 128, 0, st>>>(n, k1 ? d_k1 : nullptr, d_k2, (u32*)(g.d_ws + L.ws));
    CK(cudaGetLastError());
    CK(cudaEventRecord(g.ev[4], st));
    k256_mul_add_kernel<<<(unsigned)((n + EB_VERIFY_BLOCK - 1) / EB_VERIFY_BLOCK), EB_VERIFY_BLOCK, 0, st>>>(
        n, d_pts, (u32*)(g.d_ws + L.ws), g.gtab[curve], (u32*)(g.d_ws + L.qtab), d_out, g.d_status);
    CK(cudaGetLastError());
    CK(cudaEventRecord(g.ev[5], st));
    if (derive) status_map_kernel<<<nb, 128, 0, st>>>(n, g.d_status, ST_NEEDS_HOST, ST_THROW_NOT_VALIDATED);
    else k256_mul_add_replay_kernel<<<nb, 128, 0, st>>>(n, k1 ? d_k1 : nullptr, d_k2, d_pts, g.replay_tab, d_out, g.d_status);
    CK(cudaGetLastError());
    launches = 3;
  }
  CK(cudaEventRecord(g.ev[2], st));
  if (derive) CK(cudaMemcpy2DAsync(out_xy, len, d_out, 2 * len, len, n, cudaMemcpyDeviceToHost, st));   // x only
  else CK(cudaMemcpyAsync(out_xy, d_out, 2 * len * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(status, g.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(g.ev[3], st));
  CK(cudaStreamSynchronize(st));
  g.dev_timing_pending = false;
  g.timing = eb200_timing{};
  cudaEventElapsedTime(&g.timing.h2d_ms, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&g.timing.kernel_ms, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&g.timing.d2h_ms, g.ev[2], g.ev[3]);
  cudaEventElapsedTime(&g.timing.main_kernel_ms, g.ev[4], g.ev[5]);
  g.timing.launches = launches;
  return EB200_OK;
}

extern "C" {

int eb200_scalar_mul_batch(int curve, size_t n, const uint8_t* k, const uint8_t* points_xy, uint8_t* out_xy,
                           uint8_t* status) {
  return mul_add_common(curve, n, nullptr, k, points_xy, out_xy, status);
}

int eb200_ecdh_derive_batch(int curve, size_t n, const uint8_t* priv, const uint8_t* pub_xy, uint8_t* out_x,
                            uint8_t* status) {
  if (n && !pub_xy) return EB200_ERR_ARG;
  return mul_add_common(curve, n, nullptr, priv, pub_xy, out_x, status, true);
}

int eb200_mul_add_batch(int curve, size_t n, const uint8_t* k1, const uint8_t* k2, const uint8_t* p2_xy,
                        uint8_t* out_xy, uint8_t* status) {
  if (n && (!k1 || !p2_xy)) return EB200_ERR_ARG;
  return mul_add_common(curve, n, k1, k2, p2_xy, out_xy, status);
}

// ---- ECDSA sign (secp256k1, RFC 6979 nonces on the GPU) -------------------------------------------------
}  // extern "C"

template <class SG>
static int sw_sign_launch(int curve, size_t n, const uint8_t* d_e, const uint8_t* d_k, u32 canonical, u32* d_sws, u32* d_scr,
                          uint8_t* d_r, uint8_t* d_s, uint8_t* d_id, cudaStream_t st) {
  unsigned nb = (unsigned)((n + 127) / 128);
  size_t T = (n + SG::BATCH - 1) / SG::BATCH;
  sw_sign_nonce_kernel<SG><<<nb, 128, 0, st>>>(n, d_e, d_k, g.gtab[curve], d_sws, g.d_status);
  CK(cudaGetLastError());
  sw_sign_finish_kernel<SG><<<(unsigned)((T + 127) / 128), 128, 0, st>>>(n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, g.d_status);
  CK(cudaGetLastError());
  sw_sign_slow_kernel<SG><<<nb, 128, 0, st>>>(n, d_e, d_k, canonical, g.gtab[curve], d_r, d_s, d_id, g.d_status);
  CK(cudaGetLastError());
  return EB200_OK;
}

extern "C" {

int eb200_ecdsa_sign_batch(int curve, size_t n, const uint8_t* e, const uint8_t* priv, uint32_t flags,
                           uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid, uint8_t* status) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!e || !priv || !out_r || !out_s || !out_recid || !status) return EB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  int rc = ensure_table(curve);
  if (rc) return rc;
  const size_t len = curve_len(curve), limbs = fe_len(curve) / 4;
  if ((rc = grow(&g.d_in, &g.d_in_cap, n * (4 * len + 1) + 256))) return rc;
  if ((rc = grow(&g.d_status, &g.d_status_cap, n))) return rc;
  const size_t ws_bytes = align256(4 * limbs * 4 * n);                 // X, Y, Z, k
  if ((rc = grow(&g.d_ws, &g.d_ws_cap, ws_bytes + 2 * limbs * 4 * n))) return rc;
  uint8_t *d_e = g.d_in, *d_k = d_e + len * n, *d_r = d_k + len * n, *d_s = d_r + len * n, *d_id = d_s + len * n;
  u32 *d_sws = (u32*)g.d_ws, *d_scr = (u32*)(g.d_ws + ws_bytes);
  const u32 canonical = flags & EB200_SIGN_CANONICAL;
  cudaStream_t st = g.stream;
  CK(cudaEventRecord(g.ev[0], st));
  CK(cudaMemcpyAsync(d_e, e, len * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_k, priv, len * n, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(g.ev[1], st));
  if (curve == EB200_CURVE_P256) {
    if ((rc = sw_sign_launch<SWSign<P256, Sha256W>>(curve, n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, st))) return rc;
  } else if (curve == EB200_CURVE_P384) {
    if ((rc = sw_sign_launch<SWSign<P384, Sha384W>>(curve, n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, st))) return rc;
  } else if (curve == EB200_CURVE_P521) {     // curves.js:124, 50, 65: sha512, sha256, sha256
    if ((rc = sw_sign_launch<SWSign<P521, Sha512W>>(curve, n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, st))) return rc;
  } else if (curve == EB200_CURVE_P192) {
    if ((rc = sw_sign_launch<SWSign<P192, Sha256W>>(curve, n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, st))) return rc;
  } else if (curve == EB200_CURVE_P224) {
    if ((rc = sw_sign_launch<SWSign<P224, Sha256W>>(curve, n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, st))) return rc;
  } else {
    unsigned nb = (unsigned)((n + 127) / 128);
    size_t T = (n + PREP_BATCH - 1) / PREP_BATCH;
    k256_sign_nonce_kernel<<<nb, 128, 0, st>>>(n, d_e, d_k, g.gtab[curve], d_sws, g.d_status);
    CK(cudaGetLastError());
    k256_sign_finish_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, g.d_status);
    CK(cudaGetLastError());
    k256_sign_slow_kernel<<<nb, 128, 0, st>>>(n, d_e, d_k, canonical, g.gtab[curve], d_r, d_s, d_id, g.d_status);
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(g.ev[2], st));
  CK(cudaMemcpyAsync(out_r, d_r, len * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out_s, d_s, len * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out_recid, d_id, n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(status, g.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(g.ev[3], st));
  CK(cudaStreamSynchronize(st));
  g.dev_timing_pending = false;
  g.timing = eb200_timing{};
  cudaEventElapsedTime(&g.timing.h2d_ms, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&g.timing.kernel_ms, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&g.timing.d2h_ms, g.ev[2], g.ev[3]);
  g.timing.main_kernel_ms = g.timing.kernel_ms;
  g.timing.launches = 3;
  return EB200_OK;
}

// ---- EdDSA (ed25519) verify ---------------------------------------------------------------
size_t eb200_eddsa_verify_workspace_bytes(size_t n) { return align256((size_t)ED_ATAB_WORDS * 4 * n); }

int eb200_eddsa_verify_batch_dev(size_t n, const uint8_t* d_R, const uint8_t* d_S, const uint8_t* d_A,
                                 const uint8_t* d_h, uint8_t* d_status, void* d_workspace, void* stream) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (n && (!d_R || !d_S || !d_A || !d_h || !d_status || !d_workspace)) return EB200_ERR_ARG;
  int rc = ensure_table(EB200_CURVE_ED25519);
  if (rc) return rc;
  if (n == 0) return EB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaEventRecord(g.ev[1], st));
  CK(cudaEventRecord(g.ev[4], st));
  ed25519_verify_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, d_R, d_S, d_A, d_h, g.gtab[EB200_CURVE_ED25519],
                                                                   (u32*)d_workspace, d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[5], st));
  CK(cudaEventRecord(g.ev[2], st));
  g.timing.launches = 1;
  g.dev_timing_pending = true;
  return EB200_OK;
}

int eb200_eddsa_verify_batch(size_t n, const uint8_t* R, const uint8_t* S, const uint8_t* A, const uint8_t* h,
                             uint8_t* status) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (n == 0) return EB200_OK;
  if (!R || !S || !A || !h || !status) return EB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  int rc = ensure_table(EB200_CURVE_ED25519);
  if (rc) return rc;
  if ((rc = grow(&g.d_in, &g.d_in_cap, n * 128))) return rc;
  if ((rc = grow(&g.d_ws, &g.d_ws_cap, eb200_eddsa_verify_workspace_bytes(n)))) return rc;
  if ((rc = grow(&g.d_status, &g.d_status_cap, n))) return rc;
  uint8_t *dR = g.d_in, *dS = dR + 32 * n, *dA = dS + 32 * n, *dh = dA + 32 * n;
  cudaStream_t st = g.stream;
  CK(cudaEventRecord(g.ev[0], st));
  CK(cudaMemcpyAsync(dR, R, 32 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dS, S, 32 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dA, A, 32 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dh, h, 32 * n, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(g.ev[1], st));
  ed25519_verify_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, dR, dS, dA, dh, g.gtab[EB200_CURVE_ED25519],
                                                                   (u32*)g.d_ws, g.d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[2], st));
  CK(cudaMemcpyAsync(status, g.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(g.ev[3], st));
  CK(cudaStreamSynchronize(st));
  g.dev_timing_pending = false;
  g.timing = eb200_timing{};
  cudaEventElapsedTime(&g.timing.h2d_ms, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&g.timing.kernel_ms, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&g.timing.d2h_ms, g.ev[2], g.ev[3]);
  g.timing.main_kernel_ms = g.timing.kernel_ms;
  g.timing.launches = 1;
  return EB200_OK;
}

// EdDSA verify from raw messages: SHA-512 on the GPU, then the same verify kernel.
int eb200_eddsa_verify_batch_msgs(size_t n, const uint8_t* R, const uint8_t* S, const uint8_t* A,
                                  const uint8_t* msgs, const uint64_t* msg_off, uint8_t* status) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (n == 0) return EB200_OK;
  if (!R || !S || !A || !msg_off || !status || (!msgs && msg_off[n])) return EB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  int rc = ensure_table(EB200_CURVE_ED25519);
  if (rc) return rc;
  size_t mbytes = (size_t)msg_off[n];
  size_t off_bytes = (n + 1) * sizeof(uint64_t);
  size_t base = align256(n * 128);
  if ((rc = grow(&g.d_in, &g.d_in_cap, base + align256(off_bytes) + align256(mbytes + 1)))) return rc;
  if ((rc = grow(&g.d_ws, &g.d_ws_cap, eb200_eddsa_verify_workspace_bytes(n)))) return rc;
  if ((rc = grow(&g.d_status, &g.d_status_cap, n))) return rc;
  uint8_t *dR = g.d_in, *dS = dR + 32 * n, *dA = dS + 32 * n, *dh = dA + 32 * n;
  uint64_t* doff = (uint64_t*)(g.d_in + base);
  uint8_t* dm = g.d_in + base + align256(off_bytes);
  cudaStream_t st = g.stream;
  CK(cudaEventRecord(g.ev[0], st));
  CK(cudaMemcpyAsync(dR, R, 32 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dS, S, 32 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dA, A, 32 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(doff, msg_off, off_bytes, cudaMemcpyHostToDevice, st));
  if (mbytes) CK(cudaMemcpyAsync(dm, msgs, mbytes, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(g.ev[1], st));
  ed25519_hash_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, dR, dA, dm, doff, dh);
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[4], st));
  ed25519_verify_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, dR, dS, dA, dh, g.gtab[EB200_CURVE_ED25519],
                                                                   (u32*)g.d_ws, g.d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[5], st));
  CK(cudaEventRecord(g.ev[2], st));
  CK(cudaMemcpyAsync(status, g.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(g.ev[3], st));
  CK(cudaStreamSynchronize(st));
  g.dev_timing_pending = false;
  g.timing = eb200_timing{};
  cudaEventElapsedTime(&g.timing.h2d_ms, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&g.timing.kernel_ms, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&g.timing.d2h_ms, g.ev[2], g.ev[3]);
  cudaEventElapsedTime(&g.timing.main_kernel_ms, g.ev[4], g.ev[5]);
  g.timing.launches = 2;
  return EB200_OK;
}

// ---- curve25519 ECDH derive -------------------------------------------------------------------
int eb200_x25519_derive_batch_dev(size_t n, const uint8_t* d_priv, const uint8_t* d_pubx, uint8_t* d_out,
                                  uint8_t* d_status, void* stream) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (n && (!d_priv || !d_pubx || !d_out || !d_status)) return EB200_ERR_ARG;
  if (n == 0) return EB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaEventRecord(g.ev[1], st));
  CK(cudaEventRecord(g.ev[4], st));
  x25519_derive_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, d_priv, d_pubx, d_out, d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[5], st));
  CK(cudaEventRecord(g.ev[2], st));
  g.timing.launches = 1;
  g.dev_timing_pending = true;
  return EB200_OK;
}

int eb200_x25519_derive_batch(size_t n, const uint8_t* priv, const uint8_t* pubx, uint8_t* out, uint8_t* status) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (n == 0) return EB200_OK;
  if (!priv || !pubx || !out || !status) return EB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  int rc;
  if ((rc = grow(&g.d_in, &g.d_in_cap, n * 96))) return rc;
  if ((rc = grow(&g.d_status, &g.d_status_cap, n))) return rc;
  uint8_t *dk = g.d_in, *dx = dk + 32 * n, *dout = dx + 32 * n;
  cudaStream_t st = g.stream;
  CK(cudaEventRecord(g.ev[0], st));
  CK(cudaMemcpyAsync(dk, priv, 32 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dx, pubx, 32 * n, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(g.ev[1], st));
  x25519_derive_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, dk, dx, dout, g.d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[2], st));
  CK(cudaMemcpyAsync(out, dout, 32 * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(status, g.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(g.ev[3], st));
  CK(cudaStreamSynchronize(st));
  g.dev_timing_pending = false;
  g.timing = eb200_timing{};
  cudaEventElapsedTime(&g.timing.h2d_ms, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&g.timing.kernel_ms, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&g.timing.d2h_ms, g.ev[2], g.ev[3]);
  g.timing.main_kernel_ms = g.timing.kernel_ms;
  g.timing.launches = 1;
  return EB200_OK;
}

int eb200_selftest_fe(int curve, int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (!fe_len(curve)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  size_t bytes = n * fe_len(curve);
  u32 *da, *db, *dout;
  CK(cudaMalloc(&da, bytes)); CK(cudaMalloc(&db, bytes)); CK(cudaMalloc(&dout, bytes));
  // stream-ordered copies: a synchronous cudaMemcpy from pageable memory may return before its DMA lands,
  // and g.stream (non-blocking) does not wait for the legacy default stream
  CK(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, g.stream));
  CK(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, g.stream));
  unsigned nb = (unsigned)((n + 127) / 128);
  if (curve == EB200_CURVE_SECP256K1) k256_selftest_fe_kernel<<<nb, 128, 0, g.stream>>>(op, n, da, db, dout);
  else if (curve == EB200_CURVE_ED25519 || curve == EB200_CURVE_CURVE25519) f25_selftest_kernel<<<nb, 128, 0, g.stream>>>(op, n, da, db, dout);
  else {
#define EB_ST(C) (sw_selftest_fe_kernel<C><<<nb, 128, 0, g.stream>>>(op, n, da, db, dout), 0)
    (void)SW_DISPATCH(curve, EB_ST);
#undef EB_ST
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return EB200_OK;
}

int eb200_selftest_gtab_dims(int curve, int* windows, int* entries, int* wbits) {
  if (!windows || !entries || !wbits) return EB200_ERR_ARG;
  if (curve == EB200_CURVE_SECP256K1) { *windows = GTAB_WINDOWS; *entries = GTAB_ENTRIES; *wbits = GTAB_W; }
  else if (curve_len(curve)) {
#define EB_DIM(C) (*windows = SW<C>::GWINDOWS, *entries = SW<C>::GENTRIES, *wbits = SW<C>::GW, 0)
    (void)SW_DISPATCH(curve, EB_DIM);
#undef EB_DIM
  }
  else return EB200_ERR_UNSUPPORTED;
  return EB200_OK;
}

int eb200_selftest_gtab(int curve, uint32_t* out, size_t n_words) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  int w, en, b;
  int rc = eb200_selftest_gtab_dims(curve, &w, &en, &b);
  if (rc) return rc;
  size_t words = (size_t)w * en * 2 * (fe_len(curve) / 4);
  if (!out || n_words < words) return EB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  if ((rc = ensure_table(curve))) return rc;
  CK(cudaMemcpyAsync(out, g.gtab[curve], words * 4, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return EB200_OK;
}

}  // extern "C"
