// eb200.cu -- CUDA kernels (sm_100a) and the C ABI of libelliptic_b200.so.
// See include/elliptic_b200.h for the boundary, ecdsa_k256_body.cuh (secp256k1) and
// ecdsa_sw_body.cuh (p256 / p384) for the algorithms.  No CPU fallback exists in this
// library by design: without a CUDA device every compute entry point fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>
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
#include "ed25519_ec.cuh"
#include "sw_runtime.cuh"

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
                       const u32* __restrict__ gtab, u32* __restrict__ ws, uint8_t* __restrict__ status,
                       const uint8_t* __restrict__ kgiven) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) k256_sign_nonce_item(i, N, e, priv, gtab, ws, status, kgiven);
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
#ifndef EB_SW_MINBLOCKS_BIG
#define EB_SW_MINBLOCKS_BIG 2     // 12- and 18-limb curves (p384, p521): 255 registers
#endif
#ifndef EB_SW_MINBLOCKS8
#define EB_SW_MINBLOCKS8 4        // 8-limb curves (p256, p224): four 128-thread blocks per SM (128 registers, 224 B of
#endif                            // spill): 51.1 ms against 52.2 ms with three blocks / 168 registers at N = 2^20 (r02)
template <class C>
__global__ void __launch_bounds__(128, (C::N <= 8) ? EB_SW_MINBLOCKS8 : EB_SW_MINBLOCKS_BIG)
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
                     const u32* __restrict__ gtab, u32* __restrict__ ws, uint8_t* __restrict__ status,
                     const uint8_t* __restrict__ kgiven) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) SG::nonce_item(i, N, e, priv, gtab, ws, status, kgiven);
}
// EC.sign with the `pers` option and EC.genKeyPair({entropy, pers}): literal per-item loops on the byte-stream DRBG
template <class SG>
__global__ void __launch_bounds__(128)
sw_sign_pers_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ priv, const uint8_t* __restrict__ pers,
                    u32 np, u32 canonical, const u32* __restrict__ gtab, uint8_t* __restrict__ r, uint8_t* __restrict__ s,
                    uint8_t* __restrict__ recid, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) status[i] = SG::slow_item_pers(i, e, priv, pers, (int)np, canonical, gtab, r, s, recid);
}
template <class SG>
__global__ void __launch_bounds__(128)
sw_keygen_kernel(size_t N, const uint8_t* __restrict__ entropy, u32 ne, const uint8_t* __restrict__ pers, u32 np,
                 uint8_t* __restrict__ out_priv, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) status[i] = SG::keygen_item(i, entropy, (int)ne, pers, (int)np, out_priv);
}
__global__ void __launch_bounds__(128)
k256_sign_pers_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ priv, const uint8_t* __restrict__ pers,
                      u32 np, u32 canonical, const u32* __restrict__ gtab, uint8_t* __restrict__ r, uint8_t* __restrict__ s,
                      uint8_t* __restrict__ recid, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) status[i] = k256_sign_item_pers(i, e, priv, pers, (int)np, canonical, gtab, r, s, recid);
}
__global__ void __launch_bounds__(128)
k256_keygen_kernel(size_t N, const uint8_t* __restrict__ entropy, u32 ne, const uint8_t* __restrict__ pers, u32 np,
                   uint8_t* __restrict__ out_priv, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) status[i] = k256_keygen_item(i, entropy, (int)ne, pers, (int)np, out_priv);
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
__global__ void __launch_bounds__(128)
ed25519_sign_kernel(size_t N, const uint8_t* __restrict__ secrets, const uint8_t* __restrict__ msgs,
                    const u64* __restrict__ msg_off, const u32* __restrict__ gtab, uint8_t* __restrict__ sig,
                    uint8_t* __restrict__ pub, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  status[i] = ed25519_sign_item(i, secrets, msgs, msg_off, gtab, sig, pub);
}
// the `ec` API over ed25519 (ed25519_ec.cuh)
__global__ void __launch_bounds__(128) ed_ec_decode_pub_kernel(size_t N, const uint8_t* __restrict__ in, u32 fmt,
                                                               uint8_t* __restrict__ xy, uint8_t* __restrict__ pre) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) pre[i] = ed_ec_decode_pub(in + (fmt == EB200_PUB_SEC1_65 ? 65 : 33) * i, fmt, xy + 64 * i);
}
__global__ void __launch_bounds__(128, 3)
ed_ec_verify_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ r, const uint8_t* __restrict__ s,
                    const uint8_t* __restrict__ xy, const uint8_t* __restrict__ pre, const u32* __restrict__ gtab,
                    u32* __restrict__ atab, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) status[i] = ed_ec_verify_item(i, e, r, s, xy, pre, gtab, atab);
}
__global__ void __launch_bounds__(128)
ed_ec_sign_kernel(size_t N, const uint8_t* __restrict__ e, const uint8_t* __restrict__ priv, const uint8_t* __restrict__ kgiven,
                  const uint8_t* __restrict__ pers, u32 np, u32 canonical, const u32* __restrict__ gtab,
                  uint8_t* __restrict__ r, uint8_t* __restrict__ s, uint8_t* __restrict__ recid, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) status[i] = ed_ec_sign_item(i, e, priv, kgiven, pers, (int)np, canonical, gtab, r, s, recid);
}
__global__ void __launch_bounds__(128)
ed_ec_keygen_kernel(size_t N, const uint8_t* __restrict__ entropy, u32 ne, const uint8_t* __restrict__ pers, u32 np,
                    uint8_t* __restrict__ out_priv, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) status[i] = ed_ec_keygen_item(i, entropy, (int)ne, pers, (int)np, out_priv);
}
__global__ void __launch_bounds__(128, 3)
ed_ec_mul_add_kernel(size_t N, const uint8_t* __restrict__ k1, const uint8_t* __restrict__ k2, const uint8_t* __restrict__ pts,
                     u32 derive, const u32* __restrict__ gtab, u32* __restrict__ atab, uint8_t* __restrict__ out,
                     uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) status[i] = ed_ec_mul_add_item(i, k1, k2, pts, derive != 0, gtab, atab, out);
}
// Montgomery-curve Point.mul (mont.js:130-153): the ladder alone, no validation (derive = validate + this)
__global__ void __launch_bounds__(128, 4)
x25519_mul_kernel(size_t N, const uint8_t* __restrict__ k, const uint8_t* __restrict__ px, uint8_t* __restrict__ out,
                  uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) status[i] = x25519_mul_item(i, k, px, out);
}
__global__ void __launch_bounds__(128, 4)
x25519_derive_kernel(size_t N, const uint8_t* __restrict__ priv, const uint8_t* __restrict__ pubx,
                     uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  status[i] = x25519_derive_item(i, priv, pubx, out);
}

// run-time short curves (sw_runtime.cuh): one thread per item; op 0 mul / mulAdd, 1 add, 2 dbl, 3 validate
template <int NL>
__global__ void __launch_bounds__(128)
rt_curve_kernel(int op, size_t N, RtCurve<NL> C, const uint8_t* __restrict__ k1, const uint8_t* __restrict__ p1,
                const uint8_t* __restrict__ k2, const uint8_t* __restrict__ p2, u32 klen, uint8_t* __restrict__ out,
                uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) status[i] = RtG<NL>::item(op, i, k1, p1, k2, p2, klen, out, C);
}

// ---------------------------------------------------------------------------
// contexts: one per CUDA device, created by eb200_init(devices, ndev, flags).  Host-pointer calls are
// sharded over the initialised devices in contiguous blocks, each block driven by its own host
// thread on its own device; device-pointer calls run on the device that owns the pointers.
namespace {
constexpr int MAX_CHUNKS = 16;
}  // namespace
#define EB_MAX_CHUNKS 16
#include "chunk_plan.h"
namespace {
constexpr int MAX_DEV = 16;
constexpr int STAGE_SLOTS = 8;                       // pinned staging ring for pageable caller buffers
constexpr size_t STAGE_BYTES = (size_t)4 << 20;
constexpr size_t SHARD_MIN_ITEMS = (size_t)1 << 14;  // below this per device a second GPU does not pay

// ---- parallel host memcpy (pageable caller buffers -> pinned staging slots) ----------------------------
struct CopyJob { void* dst; const void* src; size_t bytes; };
class CopyPool {
 public:
  void run(const CopyJob* j, int n) {
    if (n <= 0) return;
    std::lock_guard<std::mutex> one(call_mu_);
    if (n == 1) { memcpy(j[0].dst, j[0].src, j[0].bytes); return; }
    {
      std::lock_guard<std::mutex> lk(m_);
      if (workers_.empty()) for (int t = 0; t < 3; t++) workers_.emplace_back([this] { loop(); });
      jobs_.store(j); njobs_.store(n); next_.store(0); pending_ = n; gen_++;
    }
    cv_work_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
    cv_work_.notify_all();
    for (auto& t : workers_) t.join();
  }
 private:
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      { std::unique_lock<std::mutex> lk(m_); cv_work_.wait(lk, [&] { return stop_ || gen_ != seen; }); if (stop_) return; seen = gen_; }
      work();
    }
  }
  void work() {
    for (;;) {
      int i = next_.fetch_add(1);
      if (i >= njobs_.load()) break;
      const CopyJob* j = jobs_.load();
      memcpy(j[i].dst, j[i].src, j[i].bytes);
      std::lock_guard<std::mutex> lk(m_);
      if (--pending_ == 0) cv_done_.notify_all();
    }
  }
  std::mutex m_, call_mu_;
  std::condition_variable cv_work_, cv_done_;
  std::vector<std::thread> workers_;
  std::atomic<const CopyJob*> jobs_{nullptr};
  std::atomic<int> njobs_{0}, next_{0};
  int pending_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
};

struct Ctx {
  std::mutex mu;                      // serialises the calls that use this device's buffers / events
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
  uint8_t* h_stage[STAGE_SLOTS] = {};
  cudaEvent_t ev_stage[STAGE_SLOTS] = {};
  bool stage_used[STAGE_SLOTS] = {};
  unsigned stage_next = 0;
  CopyPool pool;                                       // this device's staging copies (one pool per device: no cross-device serialisation)
  eb200_timing timing = {};
};
Ctx g_ctx[MAX_DEV];
int g_devs[MAX_DEV];
int g_ndev = 0;
std::atomic<unsigned> g_rr{0};
std::mutex g_mu;                                     // init / shutdown / the device list
thread_local char g_err[256] = "";
thread_local eb200_timing t_timing = {};             // timing of this thread's last host-pointer call
thread_local Ctx* t_pending = nullptr;               // device-pointer call whose events have not been read yet
thread_local unsigned t_pending_launches = 0;

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
    case EB200_CURVE_SECP256K1: case EB200_CURVE_P256: case EB200_CURVE_ED25519: return 32;
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
  else if (curve == EB200_CURVE_ED25519) { prep_words = 0; scratch_words = 0; qtab_words = ED_ATAB_WORDS; }
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


bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

struct Seg { void* dst; const void* src; size_t bytes; };
// Host -> device copies of `k` segments on stream st.  Pinned (or tiny) sources are copied directly;
// pageable ones go through the context's pinned ring, four pieces at a time with a parallel memcpy, so that
// a caller who did not pin its buffers (a Node.js Buffer) gets the pinned transfer rate.
int h2d(Ctx& c, const Seg* seg, int k, cudaStream_t st) {
  size_t total = 0;
  for (int i = 0; i < k; i++) total += seg[i].bytes;
  bool direct = total < ((size_t)256 << 10);
  if (!direct) { direct = true; for (int i = 0; i < k; i++) if (seg[i].bytes && !is_pinned(seg[i].src)) { direct = false; break; } }
  if (direct) {
    for (int i = 0; i < k; i++) if (seg[i].bytes) CK(cudaMemcpyAsync(seg[i].dst, seg[i].src, seg[i].bytes, cudaMemcpyHostToDevice, st));
    return EB200_OK;
  }
  for (int s = 0; s < STAGE_SLOTS; s++) {
    if (!c.h_stage[s]) CK(cudaHostAlloc(&c.h_stage[s], STAGE_BYTES, cudaHostAllocDefault));
    if (!c.ev_stage[s]) CK(cudaEventCreateWithFlags(&c.ev_stage[s], cudaEventDisableTiming));
  }
  CopyJob jobs[4];
  void* dsts[4];
  int nj = 0;
  int slots[4];
  for (int i = 0; i < k; i++) {
    size_t off = 0;
    while (off < seg[i].bytes) {
      size_t m = seg[i].bytes - off < STAGE_BYTES ? seg[i].bytes - off : STAGE_BYTES;
      int slot = (int)(c.stage_next++ % STAGE_SLOTS);
      if (c.stage_used[slot]) CK(cudaEventSynchronize(c.ev_stage[slot]));
      jobs[nj] = CopyJob{c.h_stage[slot], (const uint8_t*)seg[i].src + off, m};
      dsts[nj] = (uint8_t*)seg[i].dst + off;
      slots[nj] = slot;
      nj++;
      off += m;
      if (nj == 4) {
        c.pool.run(jobs, nj);
        for (int j = 0; j < nj; j++) {
          CK(cudaMemcpyAsync(dsts[j], jobs[j].dst, jobs[j].bytes, cudaMemcpyHostToDevice, st));
          CK(cudaEventRecord(c.ev_stage[slots[j]], st));
          c.stage_used[slots[j]] = true;
        }
        nj = 0;
      }
    }
  }
  if (nj) {
    c.pool.run(jobs, nj);
    for (int j = 0; j < nj; j++) {
      CK(cudaMemcpyAsync(dsts[j], jobs[j].dst, jobs[j].bytes, cudaMemcpyHostToDevice, st));
      CK(cudaEventRecord(c.ev_stage[slots[j]], st));
      c.stage_used[slots[j]] = true;
    }
  }
  return EB200_OK;
}
int h2d1(Ctx& c, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  Seg s{dst, src, bytes};
  return h2d(c, &s, 1, st);
}

template <class C>
int sw_ensure_table(Ctx& c, int curve) {
  typedef SW<C> W;
  if (c.gtab[curve]) return EB200_OK;
  size_t entries = (size_t)W::GWINDOWS * W::GENTRIES;
  CK(cudaMalloc(&c.gtab[curve], entries * 2 * W::N * 4));
  sw_gtab_kernel<C><<<(unsigned)((entries + 127) / 128), 128, 0, c.stream>>>(c.gtab[curve]);
  CK(cudaGetLastError());
  CK(cudaMalloc(&c.sw_replay_tab[curve], (size_t)SWReplay<C>::TAB_WORDS * 4));
  sw_replay_tab_kernel<C><<<1, 128, 0, c.stream>>>(c.sw_replay_tab[curve]);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c.stream));
  return EB200_OK;
}
int ensure_table(Ctx& c, int curve) {
  if (curve == EB200_CURVE_SECP256K1) {
    if (c.gtab[curve]) return EB200_OK;
    size_t entries = (size_t)GTAB_WINDOWS * GTAB_ENTRIES;
    CK(cudaMalloc(&c.gtab[curve], entries * 16 * 4));
    k256_gtab_kernel<<<(unsigned)((entries + 127) / 128), 128, 0, c.stream>>>(c.gtab[curve]);
    CK(cudaGetLastError());
    CK(cudaMalloc(&c.replay_tab, (size_t)REPLAY_TAB_WORDS * 4));
    k256_replay_tab_kernel<<<2, 128, 0, c.stream>>>(c.replay_tab);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c.stream));
    return EB200_OK;
  }
  if (curve == EB200_CURVE_P256 || curve == EB200_CURVE_P384 || curve == EB200_CURVE_P521 || curve == EB200_CURVE_P192 ||
      curve == EB200_CURVE_P224) {
#define EB_ENS(C) sw_ensure_table<C>(c, curve)
    return SW_DISPATCH(curve, EB_ENS);
#undef EB_ENS
  }
  if (curve == EB200_CURVE_ED25519) {
    if (c.gtab[curve]) return EB200_OK;
    size_t entries = (size_t)ED_GWINDOWS * ED_GENTRIES;
    CK(cudaMalloc(&c.gtab[curve], entries * 24 * 4));
    ed_gtab_kernel<<<(unsigned)((entries + 127) / 128), 128, 0, c.stream>>>(c.gtab[curve]);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c.stream));
    return EB200_OK;
  }
  return EB200_ERR_UNSUPPORTED;
}

template <class C>
int sw_launch_verify(Ctx& c, int curve, size_t n, const uint8_t* d_e, const uint8_t* d_r, const uint8_t* d_s, const uint8_t* xy,
                     const uint8_t* pre, u32* ws, u32* scratch, u32* qtab, uint8_t* d_status, cudaStream_t st,
                     unsigned pb, unsigned nb, cudaEvent_t ev_main0, cudaEvent_t* ev_main1) {
  sw_prep_kernel<C><<<pb, 128, 0, st>>>(n, d_e, d_r, d_s, ws, scratch);
  CK(cudaGetLastError());
  if (ev_main0) CK(cudaEventRecord(ev_main0, st));
  sw_verify_kernel<C><<<nb, 128, 0, st>>>(n, xy, d_r, ws, c.gtab[curve], qtab, pre, d_status);
  CK(cudaGetLastError());
  if (*ev_main1) { CK(cudaEventRecord(*ev_main1, st)); *ev_main1 = nullptr; }
  sw_replay_kernel<C><<<nb, 128, 0, st>>>(n, d_e, d_r, d_s, xy, c.sw_replay_tab[curve], d_status);
  return EB200_OK;
}

// Launches decode (if needed) + prep + verify for n items on stream st.  All pointers are device pointers.
int launch_verify(Ctx& c, int curve, size_t n, const uint8_t* d_e, const uint8_t* d_r, const uint8_t* d_s,
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
  if (curve == EB200_CURVE_ED25519) {       // the `ec` API over the Edwards preset: one kernel, per-item scalar inversion
    if (pub_fmt != EB200_PUB_XY) {
      ed_ec_decode_pub_kernel<<<nb, 128, 0, st>>>(n, d_pub, pub_fmt, d_workspace + L.xy, d_workspace + L.pre);
      CK(cudaGetLastError());
      xy = d_workspace + L.xy; pre = d_workspace + L.pre; cnt++;
    }
    if (d_der) {
      der_decode_kernel<<<nb, 128, 0, st>>>(n, 32u, d_der, d_der_off, (uint8_t*)d_r, (uint8_t*)d_s, d_workspace + L.pre, pre != nullptr);
      CK(cudaGetLastError());
      pre = d_workspace + L.pre; cnt++;
    }
    if (ev_main0) CK(cudaEventRecord(ev_main0, st));
    ed_ec_verify_kernel<<<nb, 128, 0, st>>>(n, d_e, d_r, d_s, xy, pre, c.gtab[curve], qtab, d_status);
    CK(cudaGetLastError());
    if (ev_main1) CK(cudaEventRecord(ev_main1, st));
    if (launches) *launches += cnt + 1;
    return EB200_OK;
  }
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
    k256_verify_kernel<<<vb, EB_VERIFY_BLOCK, 0, st>>>(n, xy, d_r, ws, c.gtab[curve], qtab, pre, d_status);
    CK(cudaGetLastError());
    if (ev_main1) { CK(cudaEventRecord(ev_main1, st)); ev_main1 = nullptr; }
    k256_replay_kernel<<<nb, 128, 0, st>>>(n, d_e, d_r, d_s, xy, c.replay_tab, d_status);
    cnt++;
  } else {
#define EB_VER(C) sw_launch_verify<C>(c, curve, n, d_e, d_r, d_s, xy, pre, ws, scratch, qtab, d_status, st, pb, nb, ev_main0, &ev_main1)
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

int ctx_create(Ctx& c, int device) {
  CK(cudaSetDevice(device));
  c.device = device;
  CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
  for (int i = 0; i < 6; i++) CK(cudaEventCreate(&c.ev[i]));
  for (int i = 0; i < MAX_CHUNKS; i++) {
    CK(cudaEventCreate(&c.ev_in[i]));
    CK(cudaEventCreate(&c.ev_k0[i]));
    CK(cudaEventCreate(&c.ev_k1[i]));
    CK(cudaEventCreate(&c.ev_done[i]));
  }
  c.ready = true;
  return EB200_OK;
}
void ctx_destroy(Ctx& c) {
  if (c.device < 0) return;
  cudaSetDevice(c.device);
  if (c.stream) cudaStreamSynchronize(c.stream);
  for (int k = 0; k < 16; k++) if (c.gtab[k]) { cudaFree(c.gtab[k]); c.gtab[k] = nullptr; }
  for (int k = 0; k < 16; k++) if (c.sw_replay_tab[k]) { cudaFree(c.sw_replay_tab[k]); c.sw_replay_tab[k] = nullptr; }
  if (c.replay_tab) { cudaFree(c.replay_tab); c.replay_tab = nullptr; }
  // the staging buffers may hold private keys or nonces of a signing call: wipe before release
  if (c.d_in) { cudaMemset(c.d_in, 0, c.d_in_cap); cudaFree(c.d_in); } c.d_in = nullptr; c.d_in_cap = 0;
  if (c.d_ws) { cudaMemset(c.d_ws, 0, c.d_ws_cap); cudaFree(c.d_ws); } c.d_ws = nullptr; c.d_ws_cap = 0;
  cudaFree(c.d_status); c.d_status = nullptr; c.d_status_cap = 0;
  for (int i = 0; i < 6; i++) if (c.ev[i]) { cudaEventDestroy(c.ev[i]); c.ev[i] = nullptr; }
  for (int i = 0; i < MAX_CHUNKS; i++) {
    if (c.ev_in[i]) { cudaEventDestroy(c.ev_in[i]); c.ev_in[i] = nullptr; }
    if (c.ev_k0[i]) { cudaEventDestroy(c.ev_k0[i]); c.ev_k0[i] = nullptr; }
    if (c.ev_k1[i]) { cudaEventDestroy(c.ev_k1[i]); c.ev_k1[i] = nullptr; }
    if (c.ev_done[i]) { cudaEventDestroy(c.ev_done[i]); c.ev_done[i] = nullptr; }
  }
  for (int s = 0; s < STAGE_SLOTS; s++) {
    if (c.h_stage[s]) { memset(c.h_stage[s], 0, STAGE_BYTES); cudaFreeHost(c.h_stage[s]); c.h_stage[s] = nullptr; }
    if (c.ev_stage[s]) { cudaEventDestroy(c.ev_stage[s]); c.ev_stage[s] = nullptr; }
    c.stage_used[s] = false;
  }
  if (c.stream) { cudaStreamDestroy(c.stream); c.stream = nullptr; }
  if (c.copy_stream) { cudaStreamDestroy(c.copy_stream); c.copy_stream = nullptr; }
  if (c.stream2) { cudaStreamDestroy(c.stream2); c.stream2 = nullptr; }
  c.ready = false;
  c.device = -1;
}

// The context that owns a device pointer (device-pointer entry points).
Ctx* ctx_of(const void* dptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, dptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return nullptr;
  if (a.device < 0 || a.device >= MAX_DEV || !g_ctx[a.device].ready) return nullptr;
  return &g_ctx[a.device];
}

void merge_timing(eb200_timing& a, const eb200_timing& b) {
  if (b.h2d_ms > a.h2d_ms) a.h2d_ms = b.h2d_ms;
  if (b.kernel_ms > a.kernel_ms) a.kernel_ms = b.kernel_ms;
  if (b.d2h_ms > a.d2h_ms) a.d2h_ms = b.d2h_ms;
  if (b.main_kernel_ms > a.main_kernel_ms) a.main_kernel_ms = b.main_kernel_ms;
  a.launches += b.launches;
}

// Runs fn(ctx, lo, m) over contiguous blocks of [0, n): one block per initialised device when the batch is
// large enough, each on its own host thread (the blocks never exchange data; results land in the caller's
// buffers at their own offsets).  A single-block call rotates over the devices so that concurrent callers
// spread out.  Timing: the slowest block, launches summed.
template <class F>
int run_sharded(size_t n, F&& fn) {
  int devs[MAX_DEV], nd;
  { std::lock_guard<std::mutex> lk(g_mu); nd = g_ndev; for (int i = 0; i < nd; i++) devs[i] = g_devs[i]; }
  if (nd == 0) return EB200_ERR_NOT_INIT;
  int use = (int)(n / SHARD_MIN_ITEMS);
  if (use > nd) use = nd;
  if (use < 1) use = 1;
  t_pending = nullptr;
  if (use == 1) {
    Ctx& c = g_ctx[devs[nd > 1 ? g_rr.fetch_add(1) % (unsigned)nd : 0]];
    std::lock_guard<std::mutex> lk(c.mu);
    CK(cudaSetDevice(c.device));
    c.timing = eb200_timing{};
    int rc = fn(c, (size_t)0, n);
    t_timing = c.timing;
    return rc;
  }
  int rcs[MAX_DEV];
  eb200_timing tms[MAX_DEV];
  char errs[MAX_DEV][256];
  std::thread th[MAX_DEV];
  size_t per = ((n + use - 1) / use + 127) & ~(size_t)127;
  for (int k = 0; k < use; k++) {
    size_t lo = (size_t)k * per, m = lo >= n ? 0 : (lo + per <= n ? per : n - lo);
    th[k] = std::thread([&, k, lo, m] {
      errs[k][0] = 0;
      rcs[k] = EB200_OK;
      tms[k] = eb200_timing{};
      if (!m) return;
      Ctx& c = g_ctx[devs[k]];
      std::lock_guard<std::mutex> lk(c.mu);
      cudaError_t e = cudaSetDevice(c.device);
      if (e != cudaSuccess) { rcs[k] = cuda_fail(e, "cudaSetDevice"); }
      else { c.timing = eb200_timing{}; rcs[k] = fn(c, lo, m); tms[k] = c.timing; }
      if (rcs[k]) snprintf(errs[k], sizeof errs[k], "device %d: %s", c.device, g_err);
    });
  }
  int rc = EB200_OK;
  eb200_timing tm = {};
  for (int k = 0; k < use; k++) {
    th[k].join();
    if (rcs[k] && !rc) { rc = rcs[k]; snprintf(g_err, sizeof g_err, "%s", errs[k]); }
    merge_timing(tm, tms[k]);
  }
  t_timing = tm;
  return rc;
}
}  // namespace

extern "C" {

const char* eb200_strerror(int code) {
  switch (code) {
    case EB200_OK: return "ok";
    case EB200_ERR_NO_DEVICE: return "no CUDA device available (this library has no CPU fallback)";
    case EB200_ERR_CUDA: return "CUDA error (see eb200_last_error)";
    case EB200_ERR_ARG: return "invalid argument";
    case EB200_ERR_NOT_INIT: return "eb200_init has not been called (or not for the device that owns these pointers)";
    case EB200_ERR_UNSUPPORTED: return "curve or format not supported by this build";
    default: return "unknown error";
  }
}

const char* eb200_last_error(void) { return g_err; }

int eb200_init(const int* devices, int ndev, uint32_t flags) {
  std::lock_guard<std::mutex> lk(g_mu);
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess) { cuda_fail(e, "cudaGetDeviceCount"); return EB200_ERR_NO_DEVICE; }
  if (cnt == 0) { snprintf(g_err, sizeof g_err, "no CUDA devices"); return EB200_ERR_NO_DEVICE; }
  int all[MAX_DEV];
  if (!devices || ndev <= 0) {                       // NULL / 0: every visible device
    ndev = cnt < MAX_DEV ? cnt : MAX_DEV;
    for (int i = 0; i < ndev; i++) all[i] = i;
    devices = all;
  }
  if (ndev > MAX_DEV) return EB200_ERR_ARG;
  for (int i = 0; i < ndev; i++) if (devices[i] < 0 || devices[i] >= cnt || devices[i] >= MAX_DEV) return EB200_ERR_ARG;
  for (int i = 0; i < ndev; i++) {
    Ctx& c = g_ctx[devices[i]];
    std::lock_guard<std::mutex> lc(c.mu);
    if (!c.ready) {
      int rc = ctx_create(c, devices[i]);
      if (rc) { ctx_destroy(c); return rc; }
      g_devs[g_ndev++] = devices[i];
    }
    CK(cudaSetDevice(c.device));
    // the headline curve's table is built eagerly; the others on first use (or now, with EB200_INIT_ALL_TABLES)
    int rc = ensure_table(c, EB200_CURVE_SECP256K1);
    if (!rc && (flags & EB200_INIT_ALL_TABLES))
      for (int cv = EB200_CURVE_P256; cv <= EB200_CURVE_P224 && !rc; cv++)
        if (cv != EB200_CURVE_CURVE25519) rc = ensure_table(c, cv);
    if (rc) return rc;
  }
  return EB200_OK;
}

int eb200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < g_ndev; i++) {
    Ctx& c = g_ctx[g_devs[i]];
    std::lock_guard<std::mutex> lc(c.mu);
    ctx_destroy(c);
  }
  g_ndev = 0;
  t_pending = nullptr;
  return EB200_OK;
}

int eb200_device_count(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_ndev;
}

int eb200_last_timing(eb200_timing* out) {
  if (!out) return EB200_ERR_ARG;
  if (t_pending) {
    // device-pointer call made by this thread: the caller has synchronised its stream by now
    Ctx& c = *t_pending;
    std::lock_guard<std::mutex> lk(c.mu);
    t_timing = eb200_timing{};
    if (cudaEventElapsedTime(&t_timing.kernel_ms, c.ev[1], c.ev[2]) != cudaSuccess) return EB200_ERR_CUDA;
    if (cudaEventElapsedTime(&t_timing.main_kernel_ms, c.ev[4], c.ev[5]) != cudaSuccess) return EB200_ERR_CUDA;
    t_timing.launches = t_pending_launches;
    t_pending = nullptr;
  }
  *out = t_timing;
  return EB200_OK;
}

size_t eb200_ecdsa_verify_workspace_bytes(int curve, size_t n) {
  if (!curve_ok(curve)) return 0;
  return ws_layout(curve, n).total;
}

int eb200_ecdsa_verify_batch_dev(int curve, size_t n, const uint8_t* d_e, const uint8_t* d_r,
                                 const uint8_t* d_s, const uint8_t* d_pub, uint32_t pub_fmt,
                                 uint8_t* d_status, void* d_workspace, void* stream) {
  if (!curve_ok(curve) || !fmt_ok(pub_fmt)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return eb200_device_count() ? EB200_OK : EB200_ERR_NOT_INIT;
  if (!d_e || !d_r || !d_s || !d_pub || !d_status || !d_workspace) return EB200_ERR_ARG;
  Ctx* cp = ctx_of(d_status);
  if (!cp) return EB200_ERR_NOT_INIT;
  Ctx& c = *cp;
  std::lock_guard<std::mutex> lk(c.mu);
  CK(cudaSetDevice(c.device));
  int rc = ensure_table(c, curve);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;   // NULL is the CUDA default stream, as everywhere in CUDA
  // events on the caller's stream: eb200_last_timing() reports them once the stream has been synchronised
  CK(cudaEventRecord(c.ev[1], st));
  unsigned launches = 0;
  rc = launch_verify(c, curve, n, d_e, d_r, d_s, d_pub, pub_fmt, d_status, (uint8_t*)d_workspace, st, c.ev[4], c.ev[5], &launches);
  if (rc) return rc;
  CK(cudaEventRecord(c.ev[2], st));
  t_pending = &c;
  t_pending_launches = launches;
  return EB200_OK;
}
}  // extern "C"

// Host-pointer verify on one device.  Large batches are cut into chunks: chunk k+1 is copied host->device on a
// copy stream while chunk k is being verified, and results stream back as each chunk finishes.
// chunk boundaries: chunk_plan.h (make_plan), unit-tested on the CPU through tests/hostemu
static int verify_on(Ctx& c, int curve, size_t n, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* pub,
                     uint32_t pub_fmt, uint8_t* status) {
  int rc = ensure_table(c, curve);
  if (rc) return rc;
  const size_t len = curve_len(curve), pb = pub_item_bytes(len, pub_fmt);
  const ChunkPlan P = make_plan(n);
  const int chunks = P.chunks;
  const size_t per = P.max_m;
  size_t item_in = 3 * len + pb;
  if ((rc = grow(&c.d_in, &c.d_in_cap, align256(n * item_in) + 1024))) return rc;
  // two chunks in flight (alternating compute streams, so the grid tail of chunk k is filled by chunk k+1)
  const size_t ws_slot = ws_layout(curve, per).total;
  if ((rc = grow(&c.d_ws, &c.d_ws_cap, (chunks > 1 ? 2 : 1) * ws_slot))) return rc;
  if ((rc = grow(&c.d_status, &c.d_status_cap, n))) return rc;
  uint8_t* d_e = c.d_in;
  uint8_t* d_r = d_e + n * len;
  uint8_t* d_s = d_r + n * len;
  uint8_t* d_pub = d_s + n * len;
  cudaStream_t cs = c.copy_stream;
  unsigned launches = 0;
  CK(cudaEventRecord(c.ev[0], cs));
  int used = 0;
  for (int k = 0; k < chunks; k++) {
    size_t lo = P.lo[k];
    size_t m = P.lo[k + 1] - lo;
    used = k + 1;
    Seg seg[4] = {{d_e + lo * len, e + lo * len, m * len}, {d_r + lo * len, r + lo * len, m * len},
                  {d_s + lo * len, s + lo * len, m * len}, {d_pub + lo * pb, pub + lo * pb, m * pb}};
    if ((rc = h2d(c, seg, 4, cs))) return rc;
    CK(cudaEventRecord(c.ev_in[k], cs));
    cudaStream_t ks = (k & 1) ? c.stream2 : c.stream;
    CK(cudaStreamWaitEvent(ks, c.ev_in[k], 0));
    if ((rc = launch_verify(c, curve, m, d_e + lo * len, d_r + lo * len, d_s + lo * len, d_pub + lo * pb, pub_fmt,
                            c.d_status + lo, c.d_ws + (size_t)(k & 1) * ws_slot, ks, c.ev_k0[k], c.ev_k1[k], &launches))) return rc;
    CK(cudaEventRecord(c.ev_done[k], ks));
  }
  // results: one device->host copy per chunk, behind that chunk's kernels, on the copy stream
  for (int k = 0; k < used; k++) {
    size_t lo = P.lo[k];
    size_t m = P.lo[k + 1] - lo;
    CK(cudaStreamWaitEvent(cs, c.ev_done[k], 0));
    CK(cudaMemcpyAsync(status + lo, c.d_status + lo, m, cudaMemcpyDeviceToHost, cs));
  }
  CK(cudaEventRecord(c.ev[3], cs));
  CK(cudaStreamSynchronize(cs));
  CK(cudaStreamSynchronize(c.stream));
  CK(cudaStreamSynchronize(c.stream2));
  float total = 0, t = 0;
  cudaEventElapsedTime(&total, c.ev[0], c.ev[3]);
  cudaEventElapsedTime(&c.timing.h2d_ms, c.ev[0], c.ev_in[used - 1]);       // all inputs resident
  for (int k = 0; k < used; k++) {       // chunks overlap on two streams: report the span of the main kernels
    cudaEventElapsedTime(&t, c.ev_k0[0], c.ev_k1[k]);
    if (t > c.timing.main_kernel_ms) c.timing.main_kernel_ms = t;
  }
  cudaEventElapsedTime(&t, c.ev_done[used - 1], c.ev[3]);
  c.timing.d2h_ms = t;                                                       // exposed tail copy
  c.timing.kernel_ms = total;                                                // whole call on the GPU timeline
  c.timing.launches = launches;
  return EB200_OK;
}

// common tail of the single-stream calls: events ev[0..3] = start, inputs resident, kernels done, outputs home
static int finish_timing(Ctx& c, unsigned launches, bool main_is_total) {
  CK(cudaStreamSynchronize(c.stream));
  cudaEventElapsedTime(&c.timing.h2d_ms, c.ev[0], c.ev[1]);
  cudaEventElapsedTime(&c.timing.kernel_ms, c.ev[1], c.ev[2]);
  cudaEventElapsedTime(&c.timing.d2h_ms, c.ev[2], c.ev[3]);
  if (main_is_total) c.timing.main_kernel_ms = c.timing.kernel_ms;
  else cudaEventElapsedTime(&c.timing.main_kernel_ms, c.ev[4], c.ev[5]);
  c.timing.launches = launches;
  return EB200_OK;
}

// Generic two-stream chunk pipeline for the fixed-stride calls: chunk k's inputs go up on the copy stream, its
// kernels run on stream (k & 1) (so the grid tail of one chunk is filled by the next), its outputs come home on the
// copy stream behind them.  in(lo, m, seg) / out(lo, m, seg) fill up to 8 segments (out: dst = host, src = device);
// run(lo, m, stream, slot, k) launches the kernels and records ev_k0[k] / ev_k1[k] around the main one.
template <class In, class Run, class Out>
static int run_chunked(Ctx& c, const ChunkPlan& P, unsigned launches_per_chunk, In&& in, Run&& run, Out&& out) {
  int rc;
  cudaStream_t cs = c.copy_stream;
  CK(cudaEventRecord(c.ev[0], cs));
  int used = 0;
  for (int k = 0; k < P.chunks; k++) {
    size_t lo = P.lo[k];
    size_t m = P.lo[k + 1] - lo;
    used = k + 1;
    Seg seg[8];
    int cnt = in(lo, m, seg);
    if ((rc = h2d(c, seg, cnt, cs))) return rc;
    CK(cudaEventRecord(c.ev_in[k], cs));
    cudaStream_t ks = (k & 1) ? c.stream2 : c.stream;
    CK(cudaStreamWaitEvent(ks, c.ev_in[k], 0));
    if ((rc = run(lo, m, ks, k & 1, k))) return rc;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c.ev_done[k], ks));
  }
  for (int k = 0; k < used; k++) {
    size_t lo = P.lo[k];
    size_t m = P.lo[k + 1] - lo;
    CK(cudaStreamWaitEvent(cs, c.ev_done[k], 0));
    Seg seg[8];
    int cnt = out(lo, m, seg);
    for (int j = 0; j < cnt; j++)
      if (seg[j].dst && seg[j].bytes) CK(cudaMemcpyAsync(seg[j].dst, seg[j].src, seg[j].bytes, cudaMemcpyDeviceToHost, cs));
  }
  CK(cudaEventRecord(c.ev[3], cs));
  CK(cudaStreamSynchronize(cs));
  CK(cudaStreamSynchronize(c.stream));
  CK(cudaStreamSynchronize(c.stream2));
  float total = 0, t = 0;
  cudaEventElapsedTime(&total, c.ev[0], c.ev[3]);
  cudaEventElapsedTime(&c.timing.h2d_ms, c.ev[0], c.ev_in[used - 1]);
  for (int k = 0; k < used; k++) {
    cudaEventElapsedTime(&t, c.ev_k0[0], c.ev_k1[k]);
    if (t > c.timing.main_kernel_ms) c.timing.main_kernel_ms = t;
  }
  cudaEventElapsedTime(&t, c.ev_done[used - 1], c.ev[3]);
  c.timing.d2h_ms = t;
  c.timing.kernel_ms = total;
  c.timing.launches = launches_per_chunk * (unsigned)used;
  return EB200_OK;
}

// DER-encoded signatures, parsed on the GPU (variable length: concatenated bytes + offsets; sig_off points at
// this block's first offset, all offsets are absolute into `sigs`)
static int verify_der_on(Ctx& c, int curve, size_t n, const uint8_t* e, const uint8_t* sigs, const uint64_t* sig_off,
                         const uint8_t* pub, uint32_t pub_fmt, uint8_t* status) {
  int rc = ensure_table(c, curve);
  if (rc) return rc;
  const size_t len = curve_len(curve), pb = pub_item_bytes(len, pub_fmt);
  const size_t sig_bytes = (size_t)(sig_off[n] - sig_off[0]);
  const size_t off_bytes = align256((n + 1) * 8);
  if ((rc = grow(&c.d_in, &c.d_in_cap, off_bytes + align256(n * (3 * len + pb)) + align256(sig_bytes) + 1024))) return rc;
  if ((rc = grow(&c.d_ws, &c.d_ws_cap, ws_layout(curve, n).total))) return rc;
  if ((rc = grow(&c.d_status, &c.d_status_cap, n))) return rc;
  unsigned long long* d_off = (unsigned long long*)c.d_in;
  uint8_t* d_e = c.d_in + off_bytes;
  uint8_t* d_r = d_e + n * len;
  uint8_t* d_s = d_r + n * len;
  uint8_t* d_pub = d_s + n * len;
  uint8_t* d_sig = d_pub + align256(n * pb);
  cudaStream_t st = c.stream;
  unsigned launches = 0;
  CK(cudaEventRecord(c.ev[0], st));
  Seg seg[4] = {{d_off, sig_off, (n + 1) * 8}, {d_e, e, n * len}, {d_pub, pub, n * pb}, {d_sig, sigs + sig_off[0], sig_bytes}};
  if ((rc = h2d(c, seg, 4, st))) return rc;
  CK(cudaEventRecord(c.ev[1], st));
  // offsets are used relative to sig_off[0] on the device
  if ((rc = launch_verify(c, curve, n, d_e, d_r, d_s, d_pub, pub_fmt, c.d_status, c.d_ws, st, c.ev[4], c.ev[5], &launches,
                          d_sig - sig_off[0], d_off))) return rc;
  CK(cudaEventRecord(c.ev[2], st));
  CK(cudaMemcpyAsync(status, c.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(c.ev[3], st));
  return finish_timing(c, launches, false);
}

extern "C" {

int eb200_ecdsa_verify_batch(int curve, size_t n, const uint8_t* e, const uint8_t* r,
                             const uint8_t* s, const uint8_t* pub, uint32_t pub_fmt,
                             uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve) || !fmt_ok(pub_fmt)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!e || !r || !s || !pub || !status) return EB200_ERR_ARG;
  const size_t len = curve_len(curve), pb = pub_item_bytes(len, pub_fmt);
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return verify_on(c, curve, m, e + lo * len, r + lo * len, s + lo * len, pub + lo * pb, pub_fmt, status + lo);
  });
}

int eb200_ecdsa_verify_batch_der(int curve, size_t n, const uint8_t* e, const uint8_t* sigs, const uint64_t* sig_off,
                                 const uint8_t* pub, uint32_t pub_fmt, uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve) || !fmt_ok(pub_fmt)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!e || !sigs || !sig_off || !pub || !status) return EB200_ERR_ARG;
  for (size_t i = 0; i < n; i++) if (sig_off[i + 1] < sig_off[i]) return EB200_ERR_ARG;
  const size_t len = curve_len(curve), pb = pub_item_bytes(len, pub_fmt);
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return verify_der_on(c, curve, m, e + lo * len, sigs, sig_off + lo, pub + lo * pb, pub_fmt, status + lo);
  });
}

// ---- ECDSA public-key recovery ------------------------------------------------------------
}  // extern "C"

template <class C>
static int sw_recover_launch(Ctx& c, int curve, size_t n, const uint8_t* d_e, const uint8_t* d_r, const uint8_t* d_s, const uint8_t* d_id,
                             uint8_t* d_out, const WsLayout& L, cudaStream_t st) {
  unsigned nb = (unsigned)((n + 127) / 128);
  sw_prep_recover_kernel<C><<<nb, 128, 0, st>>>(n, d_e, d_r, d_s, (u32*)(c.d_ws + L.ws));
  CK(cudaGetLastError());
  CK(cudaEventRecord(c.ev[4], st));
  sw_recover_kernel<C><<<nb, 128, 0, st>>>(n, d_r, d_id, (u32*)(c.d_ws + L.ws), c.gtab[curve], (u32*)(c.d_ws + L.qtab), d_out, c.d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c.ev[5], st));
  return EB200_OK;
}

static int recover_on(Ctx& c, int curve, size_t n, const uint8_t* e, const uint8_t* r, const uint8_t* s,
                      const uint8_t* recid, uint8_t* out_xy, uint8_t* status) {
  int rc = ensure_table(c, curve);
  if (rc) return rc;
  const size_t len = curve_len(curve);
  WsLayout L = ws_layout(curve, n);
  if ((rc = grow(&c.d_in, &c.d_in_cap, n * (5 * len + 1) + 256))) return rc;
  if ((rc = grow(&c.d_ws, &c.d_ws_cap, L.total))) return rc;
  if ((rc = grow(&c.d_status, &c.d_status_cap, n))) return rc;
  uint8_t *d_e = c.d_in, *d_r = d_e + len * n, *d_s = d_r + len * n, *d_out = d_s + len * n, *d_id = d_out + 2 * len * n;
  cudaStream_t st = c.stream;
  CK(cudaEventRecord(c.ev[0], st));
  Seg seg[4] = {{d_e, e, len * n}, {d_r, r, len * n}, {d_s, s, len * n}, {d_id, recid, n}};
  if ((rc = h2d(c, seg, 4, st))) return rc;
  CK(cudaEventRecord(c.ev[1], st));
  if (curve == EB200_CURVE_SECP256K1) {
    size_t T = (n + PREP_BATCH - 1) / PREP_BATCH;
    k256_prep_recover_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(n, d_e, d_r, d_s, (u32*)(c.d_ws + L.ws), (u32*)(c.d_ws + L.scratch));
    CK(cudaGetLastError());
    CK(cudaEventRecord(c.ev[4], st));
    k256_recover_kernel<<<(unsigned)((n + EB_VERIFY_BLOCK - 1) / EB_VERIFY_BLOCK), EB_VERIFY_BLOCK, 0, st>>>(
        n, d_r, d_id, (u32*)(c.d_ws + L.ws), c.gtab[curve], (u32*)(c.d_ws + L.qtab), d_out, c.d_status);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c.ev[5], st));
  } else {
#define EB_REC(C) sw_recover_launch<C>(c, curve, n, d_e, d_r, d_s, d_id, d_out, L, st)
    rc = SW_DISPATCH(curve, EB_REC);
#undef EB_REC
    if (rc) return rc;
  }
  CK(cudaEventRecord(c.ev[2], st));
  CK(cudaMemcpyAsync(out_xy, d_out, 2 * len * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(status, c.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(c.ev[3], st));
  return finish_timing(c, 2, false);
}

// ---- Point.mul / Point.mulAdd batches ---------------------------------------------------------------
// k1 == NULL: k2*P;  pts == NULL: k2*G;  both given: k1*G + k2*P.
template <class C>
static int sw_mul_add_launch(Ctx& c, int curve, size_t n, const uint8_t* d_k1, const uint8_t* d_k2, const uint8_t* d_pts,
                             uint8_t* d_out, const WsLayout& L, cudaStream_t st, unsigned* launches, bool derive) {
  unsigned nb = (unsigned)((n + 127) / 128);
  if (!d_pts) {
    CK(cudaEventRecord(c.ev[4], st));
    sw_mul_g_kernel<C><<<nb, 128, 0, st>>>(n, d_k2, c.gtab[curve], d_out, c.d_status);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c.ev[5], st));
    *launches = 1;
    return EB200_OK;
  }
  sw_prep_scalars_kernel<C><<<nb, 128, 0, st>>>(n, d_k1, d_k2, (u32*)(c.d_ws + L.ws));
  CK(cudaGetLastError());
  CK(cudaEventRecord(c.ev[4], st));
  sw_mul_add_kernel<C><<<nb, 128, 0, st>>>(n, d_pts, (u32*)(c.d_ws + L.ws), c.gtab[curve], (u32*)(c.d_ws + L.qtab), d_out, c.d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c.ev[5], st));
  if (derive) status_map_kernel<<<nb, 128, 0, st>>>(n, c.d_status, ST_NEEDS_HOST, ST_THROW_NOT_VALIDATED);
  else sw_mul_add_replay_kernel<C><<<nb, 128, 0, st>>>(n, d_k1, d_k2, d_pts, c.sw_replay_tab[curve], d_out, c.d_status);
  CK(cudaGetLastError());
  *launches = 3;
  return EB200_OK;
}

// derive: KeyPair.derive (ec/key.js:102-107) -- an off-curve point is the reference's
// 'public point not validated' throw instead of a replayed multiplication, and only x is returned.
static int mul_add_on(Ctx& c, int curve, size_t n, const uint8_t* k1, const uint8_t* k2, const uint8_t* pts,
                      uint8_t* out_xy, uint8_t* status, bool derive) {
  int rc = ensure_table(c, curve);
  if (rc) return rc;
  const size_t len = curve_len(curve);
  WsLayout L = ws_layout(curve, n);
  if ((rc = grow(&c.d_in, &c.d_in_cap, n * 6 * len + 256))) return rc;
  if ((rc = grow(&c.d_ws, &c.d_ws_cap, L.total))) return rc;
  if ((rc = grow(&c.d_status, &c.d_status_cap, n))) return rc;
  uint8_t *d_k1 = c.d_in, *d_k2 = d_k1 + len * n, *d_pts = d_k2 + len * n, *d_out = d_pts + 2 * len * n;
  cudaStream_t st = c.stream;
  unsigned nb = (unsigned)((n + 127) / 128), launches = 0;
  CK(cudaEventRecord(c.ev[0], st));
  Seg seg[3] = {{d_k1, k1, k1 ? len * n : 0}, {d_k2, k2, len * n}, {d_pts, pts, pts ? 2 * len * n : 0}};
  if ((rc = h2d(c, seg, 3, st))) return rc;
  CK(cudaEventRecord(c.ev[1], st));
  if (curve == EB200_CURVE_ED25519) {
    CK(cudaEventRecord(c.ev[4], st));
    ed_ec_mul_add_kernel<<<nb, 128, 0, st>>>(n, k1 ? d_k1 : nullptr, d_k2, pts ? d_pts : nullptr, derive ? 1u : 0u, c.gtab[curve],
                                             (u32*)(c.d_ws + L.qtab), d_out, c.d_status);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c.ev[5], st));
    launches = 1;
  } else if (curve != EB200_CURVE_SECP256K1) {
#define EB_MA(C) sw_mul_add_launch<C>(c, curve, n, k1 ? d_k1 : nullptr, d_k2, pts ? d_pts : nullptr, d_out, L, st, &launches, derive)
    if ((rc = SW_DISPATCH(curve, EB_MA))) return rc;
#undef EB_MA
  } else if (!pts) {
    CK(cudaEventRecord(c.ev[4], st));
    k256_mul_g_kernel<<<nb, 128, 0, st>>>(n, d_k2, c.gtab[curve], d_out, c.d_status);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c.ev[5], st));
    launches = 1;
  } else {
    k256_prep_scalars_kernel<<<nb, 128, 0, st>>>(n, k1 ? d_k1 : nullptr, d_k2, (u32*)(c.d_ws + L.ws));
    CK(cudaGetLastError());
    CK(cudaEventRecord(c.ev[4], st));
    k256_mul_add_kernel<<<(unsigned)((n + EB_VERIFY_BLOCK - 1) / EB_VERIFY_BLOCK), EB_VERIFY_BLOCK, 0, st>>>(
        n, d_pts, (u32*)(c.d_ws + L.ws), c.gtab[curve], (u32*)(c.d_ws + L.qtab), d_out, c.d_status);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c.ev[5], st));
    if (derive) status_map_kernel<<<nb, 128, 0, st>>>(n, c.d_status, ST_NEEDS_HOST, ST_THROW_NOT_VALIDATED);
    else k256_mul_add_replay_kernel<<<nb, 128, 0, st>>>(n, k1 ? d_k1 : nullptr, d_k2, d_pts, c.replay_tab, d_out, c.d_status);
    CK(cudaGetLastError());
    launches = 3;
  }
  CK(cudaEventRecord(c.ev[2], st));
  if (derive) {
    CK(cudaMemcpy2DAsync(out_xy, len, d_out, 2 * len, len, n, cudaMemcpyDeviceToHost, st));   // x only
    // the scalars of an ECDH call are private keys: do not leave them in the shared staging buffer
    CK(cudaMemsetAsync(d_k2, 0, len * n, st));
  } else {
    CK(cudaMemcpyAsync(out_xy, d_out, 2 * len * n, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaMemcpyAsync(status, c.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(c.ev[3], st));
  return finish_timing(c, launches, false);
}

static int mul_add_common(int curve, size_t n, const uint8_t* k1, const uint8_t* k2, const uint8_t* pts,
                          uint8_t* out_xy, uint8_t* status, bool derive = false) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!k2 || !out_xy || !status || (k1 && !pts)) return EB200_ERR_ARG;
  const size_t len = curve_len(curve), ol = derive ? len : 2 * len;
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return mul_add_on(c, curve, m, k1 ? k1 + lo * len : nullptr, k2 + lo * len, pts ? pts + lo * 2 * len : nullptr,
                      out_xy + lo * ol, status + lo, derive);
  });
}

// ---- ECDSA sign (RFC 6979 nonces on the GPU) ---------------------------------------------------------
// mode: kgiven != NULL -> the caller's nonces, one attempt (items the reference would `continue` on come back as
// EB200_ST_RETRY); pers != NULL -> the literal loop on the byte-stream DRBG; neither -> RFC 6979 fast pipeline.
template <class SG>
static int sw_sign_launch(Ctx& c, int curve, size_t n, const uint8_t* d_e, const uint8_t* d_k, u32 canonical, u32* d_sws, u32* d_scr,
                          uint8_t* d_r, uint8_t* d_s, uint8_t* d_id, cudaStream_t st, const uint8_t* d_kgiven,
                          const uint8_t* d_pers, u32 np) {
  unsigned nb = (unsigned)((n + 127) / 128);
  if (d_pers) {
    sw_sign_pers_kernel<SG><<<nb, 128, 0, st>>>(n, d_e, d_k, d_pers, np, canonical, c.gtab[curve], d_r, d_s, d_id, c.d_status);
    CK(cudaGetLastError());
    return EB200_OK;
  }
  size_t T = (n + SG::BATCH - 1) / SG::BATCH;
  sw_sign_nonce_kernel<SG><<<nb, 128, 0, st>>>(n, d_e, d_k, c.gtab[curve], d_sws, c.d_status, d_kgiven);
  CK(cudaGetLastError());
  sw_sign_finish_kernel<SG><<<(unsigned)((T + 127) / 128), 128, 0, st>>>(n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, c.d_status);
  CK(cudaGetLastError());
  if (d_kgiven) status_map_kernel<<<nb, 128, 0, st>>>(n, c.d_status, ST_NEEDS_HOST, EB200_ST_RETRY);
  else sw_sign_slow_kernel<SG><<<nb, 128, 0, st>>>(n, d_e, d_k, canonical, c.gtab[curve], d_r, d_s, d_id, c.d_status);
  CK(cudaGetLastError());
  return EB200_OK;
}

static int sign_on(Ctx& c, int curve, size_t n, const uint8_t* e, const uint8_t* priv, uint32_t flags,
                   uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid, uint8_t* status,
                   const uint8_t* kgiven = nullptr, const uint8_t* pers = nullptr, size_t np = 0) {
  int rc = ensure_table(c, curve);
  if (rc) return rc;
  const size_t len = curve_len(curve), limbs = fe_len(curve) / 4;
  if ((rc = grow(&c.d_in, &c.d_in_cap, n * (5 * len + 1) + align256(np + 1) + 512))) return rc;
  if ((rc = grow(&c.d_status, &c.d_status_cap, n))) return rc;
  const size_t ws_bytes = align256(4 * limbs * 4 * n);                 // X, Y, Z, k
  const size_t scr_bytes = 2 * limbs * 4 * n;
  if ((rc = grow(&c.d_ws, &c.d_ws_cap, ws_bytes + scr_bytes))) return rc;
  uint8_t *d_e = c.d_in, *d_k = d_e + len * n, *d_r = d_k + len * n, *d_s = d_r + len * n, *d_kg = d_s + len * n, *d_id = d_kg + len * n;
  uint8_t* d_pers = (uint8_t*)(((uintptr_t)(d_id + n) + 255) & ~(uintptr_t)255);
  u32 *d_sws = (u32*)c.d_ws, *d_scr = (u32*)(c.d_ws + ws_bytes);
  const u32 canonical = flags & EB200_SIGN_CANONICAL;
  cudaStream_t st = c.stream;
  CK(cudaEventRecord(c.ev[0], st));
  Seg seg[4] = {{d_e, e, len * n}, {d_k, priv, len * n}, {d_kg, kgiven, kgiven ? len * n : 0}, {d_pers, pers, pers ? np : 0}};
  if ((rc = h2d(c, seg, 4, st))) return rc;
  CK(cudaEventRecord(c.ev[1], st));
  const uint8_t* kg = kgiven ? d_kg : nullptr;
  const uint8_t* pp = pers ? d_pers : nullptr;
  if (curve == EB200_CURVE_ED25519) {
    static const uint8_t* none = nullptr;
    (void)none;
    ed_ec_sign_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, d_e, d_k, kg, pp, (u32)np, canonical, c.gtab[curve], d_r, d_s, d_id, c.d_status);
    CK(cudaGetLastError());
  } else if (curve == EB200_CURVE_P256) {
    if ((rc = sw_sign_launch<SWSign<P256, Sha256W>>(c, curve, n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, st, kg, pp, (u32)np))) return rc;
  } else if (curve == EB200_CURVE_P384) {
    if ((rc = sw_sign_launch<SWSign<P384, Sha384W>>(c, curve, n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, st, kg, pp, (u32)np))) return rc;
  } else if (curve == EB200_CURVE_P521) {     // curves.js:124, 50, 65: sha512, sha256, sha256
    if ((rc = sw_sign_launch<SWSign<P521, Sha512W>>(c, curve, n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, st, kg, pp, (u32)np))) return rc;
  } else if (curve == EB200_CURVE_P192) {
    if ((rc = sw_sign_launch<SWSign<P192, Sha256W>>(c, curve, n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, st, kg, pp, (u32)np))) return rc;
  } else if (curve == EB200_CURVE_P224) {
    if ((rc = sw_sign_launch<SWSign<P224, Sha256W>>(c, curve, n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, st, kg, pp, (u32)np))) return rc;
  } else {
    unsigned nb = (unsigned)((n + 127) / 128);
    size_t T = (n + PREP_BATCH - 1) / PREP_BATCH;
    if (pp) {
      k256_sign_pers_kernel<<<nb, 128, 0, st>>>(n, d_e, d_k, pp, (u32)np, canonical, c.gtab[curve], d_r, d_s, d_id, c.d_status);
      CK(cudaGetLastError());
    } else {
      k256_sign_nonce_kernel<<<nb, 128, 0, st>>>(n, d_e, d_k, c.gtab[curve], d_sws, c.d_status, kg);
      CK(cudaGetLastError());
      k256_sign_finish_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(n, d_e, d_k, canonical, d_sws, d_scr, d_r, d_s, d_id, c.d_status);
      CK(cudaGetLastError());
      if (kg) status_map_kernel<<<nb, 128, 0, st>>>(n, c.d_status, ST_NEEDS_HOST, EB200_ST_RETRY);
      else k256_sign_slow_kernel<<<nb, 128, 0, st>>>(n, d_e, d_k, canonical, c.gtab[curve], d_r, d_s, d_id, c.d_status);
      CK(cudaGetLastError());
    }
  }
  CK(cudaEventRecord(c.ev[2], st));
  CK(cudaMemcpyAsync(out_r, d_r, len * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out_s, d_s, len * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out_recid, d_id, n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(status, c.d_status, n, cudaMemcpyDeviceToHost, st));
  // private keys, nonces k and k*G live in buffers that later calls reuse: wipe them before returning
  CK(cudaMemsetAsync(d_k, 0, len * n, st));
  if (kgiven) CK(cudaMemsetAsync(d_kg, 0, len * n, st));
  CK(cudaMemsetAsync(c.d_ws, 0, ws_bytes + scr_bytes, st));
  CK(cudaEventRecord(c.ev[3], st));
  return finish_timing(c, pp ? 1 : 3, true);
}

// EC.genKeyPair({entropy, pers}) (ec/index.js:55-79): private keys from HMAC-DRBG(entropy_i, nonce = n, pers), then
// the public points k*G.  entropy: n x ne bytes.
static int keygen_on(Ctx& c, int curve, size_t n, const uint8_t* entropy, size_t ne, const uint8_t* pers, size_t np,
                     uint8_t* out_priv, uint8_t* out_pub, uint8_t* status) {
  int rc = ensure_table(c, curve);
  if (rc) return rc;
  const size_t len = curve_len(curve);
  if ((rc = grow(&c.d_in, &c.d_in_cap, align256(n * ne) + align256(np + 1) + n * 3 * len + 512))) return rc;
  if ((rc = grow(&c.d_status, &c.d_status_cap, n))) return rc;
  uint8_t* d_ent = c.d_in;
  uint8_t* d_pers = d_ent + align256(n * ne);
  uint8_t* d_priv = d_pers + align256(np + 1);
  uint8_t* d_pub = d_priv + len * n;
  cudaStream_t st = c.stream;
  unsigned nb = (unsigned)((n + 127) / 128);
  CK(cudaEventRecord(c.ev[0], st));
  Seg seg[2] = {{d_ent, entropy, n * ne}, {d_pers, pers, pers ? np : 0}};
  if ((rc = h2d(c, seg, 2, st))) return rc;
  CK(cudaEventRecord(c.ev[1], st));
  const uint8_t* pp = pers ? d_pers : nullptr;
  switch (curve) {
    case EB200_CURVE_ED25519: ed_ec_keygen_kernel<<<nb, 128, 0, st>>>(n, d_ent, (u32)ne, pp, (u32)np, d_priv, c.d_status); break;
    case EB200_CURVE_SECP256K1: k256_keygen_kernel<<<nb, 128, 0, st>>>(n, d_ent, (u32)ne, pp, (u32)np, d_priv, c.d_status); break;
    case EB200_CURVE_P256: sw_keygen_kernel<SWSign<P256, Sha256W>><<<nb, 128, 0, st>>>(n, d_ent, (u32)ne, pp, (u32)np, d_priv, c.d_status); break;
    case EB200_CURVE_P384: sw_keygen_kernel<SWSign<P384, Sha384W>><<<nb, 128, 0, st>>>(n, d_ent, (u32)ne, pp, (u32)np, d_priv, c.d_status); break;
    case EB200_CURVE_P521: sw_keygen_kernel<SWSign<P521, Sha512W>><<<nb, 128, 0, st>>>(n, d_ent, (u32)ne, pp, (u32)np, d_priv, c.d_status); break;
    case EB200_CURVE_P192: sw_keygen_kernel<SWSign<P192, Sha256W>><<<nb, 128, 0, st>>>(n, d_ent, (u32)ne, pp, (u32)np, d_priv, c.d_status); break;
    default: sw_keygen_kernel<SWSign<P224, Sha256W>><<<nb, 128, 0, st>>>(n, d_ent, (u32)ne, pp, (u32)np, d_priv, c.d_status); break;
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(status, c.d_status, n, cudaMemcpyDeviceToHost, st));      // the keygen verdicts (the mul below reuses d_status)
  if (curve == EB200_CURVE_SECP256K1) k256_mul_g_kernel<<<nb, 128, 0, st>>>(n, d_priv, c.gtab[curve], d_pub, c.d_status);
  else if (curve == EB200_CURVE_ED25519) ed_ec_mul_add_kernel<<<nb, 128, 0, st>>>(n, nullptr, d_priv, nullptr, 0u, c.gtab[curve], nullptr, d_pub, c.d_status);
  else {
#define EB_KG(C) (sw_mul_g_kernel<C><<<nb, 128, 0, st>>>(n, d_priv, c.gtab[curve], d_pub, c.d_status), 0)
    (void)SW_DISPATCH(curve, EB_KG);
#undef EB_KG
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(c.ev[2], st));
  CK(cudaMemcpyAsync(out_priv, d_priv, len * n, cudaMemcpyDeviceToHost, st));
  if (out_pub) CK(cudaMemcpyAsync(out_pub, d_pub, 2 * len * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemsetAsync(d_ent, 0, n * ne, st));
  CK(cudaMemsetAsync(d_priv, 0, len * n, st));
  CK(cudaEventRecord(c.ev[3], st));
  return finish_timing(c, 2, true);
}

extern "C" {

int eb200_ecdsa_recover_batch(int curve, size_t n, const uint8_t* e, const uint8_t* r, const uint8_t* s,
                              const uint8_t* recid, uint8_t* out_xy, uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!e || !r || !s || !recid || !out_xy || !status) return EB200_ERR_ARG;
  if (curve == EB200_CURVE_ED25519) return EB200_ERR_UNSUPPORTED;      // recoverPubKey over the Edwards preset is not accelerated
  const size_t len = curve_len(curve);
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return recover_on(c, curve, m, e + lo * len, r + lo * len, s + lo * len, recid + lo, out_xy + lo * 2 * len, status + lo);
  });
}

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

int eb200_ecdsa_sign_batch(int curve, size_t n, const uint8_t* e, const uint8_t* priv, uint32_t flags,
                           uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid, uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!e || !priv || !out_r || !out_s || !out_recid || !status) return EB200_ERR_ARG;
  const size_t len = curve_len(curve);
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return sign_on(c, curve, m, e + lo * len, priv + lo * len, flags, out_r + lo * len, out_s + lo * len, out_recid + lo, status + lo);
  });
}

// EC.sign with options.k (ec/index.js:154-157): one attempt with the caller's nonces
int eb200_ecdsa_sign_batch_k(int curve, size_t n, const uint8_t* e, const uint8_t* priv, const uint8_t* k, uint32_t flags,
                             uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid, uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!e || !priv || !k || !out_r || !out_s || !out_recid || !status) return EB200_ERR_ARG;
  const size_t len = curve_len(curve);
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return sign_on(c, curve, m, e + lo * len, priv + lo * len, flags, out_r + lo * len, out_s + lo * len, out_recid + lo, status + lo,
                   k + lo * len);
  });
}

// EC.sign with options.pers (ec/index.js:143-151; the bytes after persEnc decoding, shared by the batch)
int eb200_ecdsa_sign_batch_pers(int curve, size_t n, const uint8_t* e, const uint8_t* priv, const uint8_t* pers, size_t pers_len,
                                uint32_t flags, uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid, uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!e || !priv || !out_r || !out_s || !out_recid || !status || (pers_len && !pers) || pers_len > (1u << 20)) return EB200_ERR_ARG;
  static const uint8_t none = 0;
  const uint8_t* pp = pers_len ? pers : &none;
  const size_t len = curve_len(curve);
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return sign_on(c, curve, m, e + lo * len, priv + lo * len, flags, out_r + lo * len, out_s + lo * len, out_recid + lo, status + lo,
                   nullptr, pp, pers_len);
  });
}

// EC.genKeyPair({entropy, pers}) (ec/index.js:55-79)
int eb200_ec_keygen_batch(int curve, size_t n, const uint8_t* entropy, size_t entropy_len, const uint8_t* pers, size_t pers_len,
                          uint8_t* out_priv, uint8_t* out_pub_xy, uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (!curve_ok(curve)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!entropy || !entropy_len || !out_priv || !status || (pers_len && !pers) || pers_len > (1u << 20) || entropy_len > (1u << 16)) return EB200_ERR_ARG;
  const size_t len = curve_len(curve);
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return keygen_on(c, curve, m, entropy + lo * entropy_len, entropy_len, pers_len ? pers : nullptr, pers_len,
                     out_priv + lo * len, out_pub_xy ? out_pub_xy + lo * 2 * len : nullptr, status + lo);
  });
}

// ---- EdDSA (ed25519) verify ---------------------------------------------------------------
size_t eb200_eddsa_verify_workspace_bytes(size_t n) { return align256((size_t)ED_ATAB_WORDS * 4 * n); }

int eb200_eddsa_verify_batch_dev(size_t n, const uint8_t* d_R, const uint8_t* d_S, const uint8_t* d_A,
                                 const uint8_t* d_h, uint8_t* d_status, void* d_workspace, void* stream) {
  if (n == 0) return eb200_device_count() ? EB200_OK : EB200_ERR_NOT_INIT;
  if (!d_R || !d_S || !d_A || !d_h || !d_status || !d_workspace) return EB200_ERR_ARG;
  Ctx* cp = ctx_of(d_status);
  if (!cp) return EB200_ERR_NOT_INIT;
  Ctx& c = *cp;
  std::lock_guard<std::mutex> lk(c.mu);
  CK(cudaSetDevice(c.device));
  int rc = ensure_table(c, EB200_CURVE_ED25519);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaEventRecord(c.ev[1], st));
  CK(cudaEventRecord(c.ev[4], st));
  ed25519_verify_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, d_R, d_S, d_A, d_h, c.gtab[EB200_CURVE_ED25519],
                                                                   (u32*)d_workspace, d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c.ev[5], st));
  CK(cudaEventRecord(c.ev[2], st));
  t_pending = &c;
  t_pending_launches = 1;
  return EB200_OK;
}
}  // extern "C"

// h == NULL: raw messages (msgs + offsets; msg_off points at this block's first offset, offsets absolute), SHA-512 on the GPU
static int eddsa_on(Ctx& c, size_t n, const uint8_t* R, const uint8_t* S, const uint8_t* A, const uint8_t* h,
                    const uint8_t* msgs, const uint64_t* msg_off, uint8_t* status) {
  int rc = ensure_table(c, EB200_CURVE_ED25519);
  if (rc) return rc;
  const ChunkPlan P = make_plan(n);
  const int chunks = P.chunks;
  const size_t per = P.max_m;
  size_t mbytes = h ? 0 : (size_t)(msg_off[n] - msg_off[0]);
  size_t off_bytes = h ? 0 : (n + 1) * sizeof(uint64_t);
  size_t base = align256(n * 128);
  const size_t ws_slot = eb200_eddsa_verify_workspace_bytes(per < n ? per : n);
  if ((rc = grow(&c.d_in, &c.d_in_cap, base + align256(off_bytes) + align256(mbytes + 1)))) return rc;
  if ((rc = grow(&c.d_ws, &c.d_ws_cap, (chunks > 1 ? 2 : 1) * ws_slot))) return rc;
  if ((rc = grow(&c.d_status, &c.d_status_cap, n))) return rc;
  uint8_t *dR = c.d_in, *dS = dR + 32 * n, *dA = dS + 32 * n, *dh = dA + 32 * n;
  uint64_t* doff = (uint64_t*)(c.d_in + base);
  uint8_t* dm = c.d_in + base + align256(off_bytes);
  const u32* gt = c.gtab[EB200_CURVE_ED25519];
  return run_chunked(c, P, h ? 1u : 2u,
    [&](size_t lo, size_t m, Seg* seg) {
      seg[0] = {dR + 32 * lo, R + 32 * lo, 32 * m};
      seg[1] = {dS + 32 * lo, S + 32 * lo, 32 * m};
      seg[2] = {dA + 32 * lo, A + 32 * lo, 32 * m};
      if (h) { seg[3] = {dh + 32 * lo, h + 32 * lo, 32 * m}; return 4; }
      seg[3] = {doff + lo, msg_off + lo, (m + 1) * sizeof(uint64_t)};
      seg[4] = {dm + (msg_off[lo] - msg_off[0]), msgs + msg_off[lo], (size_t)(msg_off[lo + m] - msg_off[lo])};
      return 5;
    },
    [&](size_t lo, size_t m, cudaStream_t ks, int slot, int k) {
      unsigned nb = (unsigned)((m + 127) / 128);
      if (!h) ed25519_hash_kernel<<<nb, 128, 0, ks>>>(m, dR + 32 * lo, dA + 32 * lo, dm - msg_off[0], doff + lo, dh + 32 * lo);
      CK(cudaEventRecord(c.ev_k0[k], ks));
      ed25519_verify_kernel<<<nb, 128, 0, ks>>>(m, dR + 32 * lo, dS + 32 * lo, dA + 32 * lo, dh + 32 * lo, gt,
                                                (u32*)(c.d_ws + (size_t)slot * ws_slot), c.d_status + lo);
      CK(cudaEventRecord(c.ev_k1[k], ks));
      return EB200_OK;
    },
    [&](size_t lo, size_t m, Seg* seg) {
      seg[0] = {status + lo, c.d_status + lo, m};
      return 1;
    });
}

// EDDSA.sign batch: secrets n x 32, raw messages (offsets absolute, msg_off points at this block's first one)
static int eddsa_sign_on(Ctx& c, size_t n, const uint8_t* secrets, const uint8_t* msgs, const uint64_t* msg_off,
                         uint8_t* sig, uint8_t* pub, uint8_t* status) {
  int rc = ensure_table(c, EB200_CURVE_ED25519);
  if (rc) return rc;
  size_t mbytes = (size_t)(msg_off[n] - msg_off[0]);
  size_t off_bytes = (n + 1) * sizeof(uint64_t);
  size_t base = align256(n * 128);                       // secrets | sig | pub
  if ((rc = grow(&c.d_in, &c.d_in_cap, base + align256(off_bytes) + align256(mbytes + 1)))) return rc;
  if ((rc = grow(&c.d_status, &c.d_status_cap, n))) return rc;
  uint8_t *dsec = c.d_in, *dsig = dsec + 32 * n, *dpub = dsig + 64 * n;
  uint64_t* doff = (uint64_t*)(c.d_in + base);
  uint8_t* dm = c.d_in + base + align256(off_bytes);
  cudaStream_t st = c.stream;
  CK(cudaEventRecord(c.ev[0], st));
  Seg seg[3] = {{dsec, secrets, 32 * n}, {doff, msg_off, off_bytes}, {dm, msgs + msg_off[0], mbytes}};
  if ((rc = h2d(c, seg, 3, st))) return rc;
  CK(cudaEventRecord(c.ev[1], st));
  ed25519_sign_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, dsec, dm - msg_off[0], doff, c.gtab[EB200_CURVE_ED25519], dsig, dpub, c.d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c.ev[2], st));
  CK(cudaMemcpyAsync(sig, dsig, 64 * n, cudaMemcpyDeviceToHost, st));
  if (pub) CK(cudaMemcpyAsync(pub, dpub, 32 * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(status, c.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemsetAsync(dsec, 0, 32 * n, st));             // secrets do not stay in the shared buffer
  CK(cudaEventRecord(c.ev[3], st));
  return finish_timing(c, 1, true);
}

static int x25519_on(Ctx& c, size_t n, const uint8_t* priv, const uint8_t* pubx, uint8_t* out, uint8_t* status, bool validate) {
  int rc;
  const ChunkPlan P = make_plan(n);
  if ((rc = grow(&c.d_in, &c.d_in_cap, n * 96))) return rc;
  if ((rc = grow(&c.d_status, &c.d_status_cap, n))) return rc;
  uint8_t *dk = c.d_in, *dx = dk + 32 * n, *dout = dx + 32 * n;
  return run_chunked(c, P, 1u,
    [&](size_t lo, size_t m, Seg* seg) {
      seg[0] = {dk + 32 * lo, priv + 32 * lo, 32 * m};
      seg[1] = {dx + 32 * lo, pubx + 32 * lo, 32 * m};
      return 2;
    },
    [&](size_t lo, size_t m, cudaStream_t ks, int, int k) {
      unsigned nb = (unsigned)((m + 127) / 128);
      CK(cudaEventRecord(c.ev_k0[k], ks));
      if (validate) x25519_derive_kernel<<<nb, 128, 0, ks>>>(m, dk + 32 * lo, dx + 32 * lo, dout + 32 * lo, c.d_status + lo);
      else x25519_mul_kernel<<<nb, 128, 0, ks>>>(m, dk + 32 * lo, dx + 32 * lo, dout + 32 * lo, c.d_status + lo);
      CK(cudaEventRecord(c.ev_k1[k], ks));
      CK(cudaMemsetAsync(dk + 32 * lo, 0, 32 * m, ks));   // private scalars do not stay in the shared buffer
      return EB200_OK;
    },
    [&](size_t lo, size_t m, Seg* seg) {
      seg[0] = {out + 32 * lo, dout + 32 * lo, 32 * m};
      seg[1] = {status + lo, c.d_status + lo, m};
      return 2;
    });
}

extern "C" {

int eb200_eddsa_verify_batch(size_t n, const uint8_t* R, const uint8_t* S, const uint8_t* A, const uint8_t* h,
                             uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (n == 0) return EB200_OK;
  if (!R || !S || !A || !h || !status) return EB200_ERR_ARG;
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return eddsa_on(c, m, R + 32 * lo, S + 32 * lo, A + 32 * lo, h + 32 * lo, nullptr, nullptr, status + lo);
  });
}

// EdDSA verify from raw messages: SHA-512 on the GPU, then the same verify kernel.
int eb200_eddsa_verify_batch_msgs(size_t n, const uint8_t* R, const uint8_t* S, const uint8_t* A,
                                  const uint8_t* msgs, const uint64_t* msg_off, uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (n == 0) return EB200_OK;
  if (!R || !S || !A || !msg_off || !status || (!msgs && msg_off[n])) return EB200_ERR_ARG;
  for (size_t i = 0; i < n; i++) if (msg_off[i + 1] < msg_off[i]) return EB200_ERR_ARG;
  static const uint8_t none = 0;
  const uint8_t* mp = msgs ? msgs : &none;
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return eddsa_on(c, m, R + 32 * lo, S + 32 * lo, A + 32 * lo, nullptr, mp, msg_off + lo, status + lo);
  });
}

// EDDSA.prototype.sign (eddsa/index.js:34-44) for keys given as 32-byte secrets (eddsa.keyFromSecret)
int eb200_eddsa_sign_batch(size_t n, const uint8_t* secrets, const uint8_t* msgs, const uint64_t* msg_off,
                           uint8_t* out_sig, uint8_t* out_pub, uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (n == 0) return EB200_OK;
  if (!secrets || !msg_off || !out_sig || !status || (!msgs && msg_off[n])) return EB200_ERR_ARG;
  for (size_t i = 0; i < n; i++) if (msg_off[i + 1] < msg_off[i]) return EB200_ERR_ARG;
  static const uint8_t none = 0;
  const uint8_t* mp = msgs ? msgs : &none;
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return eddsa_sign_on(c, m, secrets + 32 * lo, mp, msg_off + lo, out_sig + 64 * lo, out_pub ? out_pub + 32 * lo : nullptr, status + lo);
  });
}

// ---- curve25519 ECDH derive -------------------------------------------------------------------
int eb200_x25519_derive_batch_dev(size_t n, const uint8_t* d_priv, const uint8_t* d_pubx, uint8_t* d_out,
                                  uint8_t* d_status, void* stream) {
  if (n == 0) return eb200_device_count() ? EB200_OK : EB200_ERR_NOT_INIT;
  if (!d_priv || !d_pubx || !d_out || !d_status) return EB200_ERR_ARG;
  Ctx* cp = ctx_of(d_status);
  if (!cp) return EB200_ERR_NOT_INIT;
  Ctx& c = *cp;
  std::lock_guard<std::mutex> lk(c.mu);
  CK(cudaSetDevice(c.device));
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaEventRecord(c.ev[1], st));
  CK(cudaEventRecord(c.ev[4], st));
  x25519_derive_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, d_priv, d_pubx, d_out, d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c.ev[5], st));
  CK(cudaEventRecord(c.ev[2], st));
  t_pending = &c;
  t_pending_launches = 1;
  return EB200_OK;
}

int eb200_x25519_derive_batch(size_t n, const uint8_t* priv, const uint8_t* pubx, uint8_t* out, uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (n == 0) return EB200_OK;
  if (!priv || !pubx || !out || !status) return EB200_ERR_ARG;
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return x25519_on(c, m, priv + 32 * lo, pubx + 32 * lo, out + 32 * lo, status + lo, true);
  });
}

// Montgomery-curve Point.mul (mont.js:130-153): x(k * P) for x-only points, k any 256-bit integer, no validation
int eb200_x25519_mul_batch(size_t n, const uint8_t* k, const uint8_t* px, uint8_t* out_x, uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (n == 0) return EB200_OK;
  if (!k || !px || !out_x || !status) return EB200_ERR_ARG;
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return x25519_on(c, m, k + 32 * lo, px + 32 * lo, out_x + 32 * lo, status + lo, false);
  });
}

// ---- run-time short curves (the generic .curve API) ------------------------------------------------
}  // extern "C"

namespace {
// big-endian bytes -> little-endian limbs (zero-extended); false if the value does not fit
template <int NL> bool rt_limbs(u32* r, const uint8_t* be, size_t len) {
  for (int i = 0; i < NL; i++) r[i] = 0;
  for (size_t b = 0; b < len; b++) {
    size_t bit = 8 * (len - 1 - b);
    if (bit / 32 >= (size_t)NL) { if (be[b]) return false; continue; }
    r[bit / 32] |= (u32)be[b] << (bit % 32);
  }
  return true;
}
// r = 2 r mod p
template <int NL> void rt_dbl_mod(u32* r, const u32* p) {
  u32 hi = r[NL - 1] >> 31;
  for (int i = NL - 1; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 31);
  r[0] <<= 1;
  u32 d[NL];
  u64 bw = 0;
  for (int i = 0; i < NL; i++) { u64 t = (u64)r[i] - p[i] - bw; d[i] = (u32)t; bw = (t >> 32) & 1; }
  if (hi || !bw) for (int i = 0; i < NL; i++) r[i] = d[i];
}
template <int NL> int rt_make(RtCurve<NL>& C, const eb200_short_curve* c) {
  if (!rt_limbs<NL>(C.p, c->p, c->len)) return EB200_ERR_ARG;
  u32 a[NL], b[NL];
  if (!rt_limbs<NL>(a, c->a, c->len) || !rt_limbs<NL>(b, c->b, c->len)) return EB200_ERR_ARG;
  if (!(C.p[0] & 1)) return EB200_ERR_ARG;                        // Montgomery arithmetic needs an odd modulus
  bool gt3 = false;
  for (int i = 1; i < NL; i++) gt3 = gt3 || C.p[i];
  if (!gt3 && C.p[0] <= 3) return EB200_ERR_ARG;
  u32 inv = 1;                                                     // -p^-1 mod 2^32 by Newton iteration
  for (int i = 0; i < 5; i++) inv *= 2 - C.p[0] * inv;
  C.n0inv = 0u - inv;
  // R mod p and R^2 mod p by repeated doubling of 1
  u32 r[NL];
  for (int i = 0; i < NL; i++) r[i] = i == 0;
  for (int i = 0; i < 32 * NL; i++) rt_dbl_mod<NL>(r, C.p);
  for (int i = 0; i < NL; i++) C.r1[i] = r[i];
  for (int i = 0; i < 32 * NL; i++) rt_dbl_mod<NL>(r, C.p);
  for (int i = 0; i < NL; i++) C.r2[i] = r[i];
  // a, b reduced and in Montgomery form: x R mod p by doubling x 32 NL times (host side, once per call)
  auto to_mont = [&](u32* x) {
    for (;;) {                                                    // reduce x below p first
      u32 d[NL]; u64 bw = 0;
      for (int i = 0; i < NL; i++) { u64 t = (u64)x[i] - C.p[i] - bw; d[i] = (u32)t; bw = (t >> 32) & 1; }
      if (bw) break;
      for (int i = 0; i < NL; i++) x[i] = d[i];
    }
    for (int i = 0; i < 32 * NL; i++) rt_dbl_mod<NL>(x, C.p);
  };
  bool az = true;
  to_mont(a); to_mont(b);
  for (int i = 0; i < NL; i++) { C.a[i] = a[i]; C.b[i] = b[i]; az = az && a[i] == 0; }
  C.a_is_zero = az;
  C.len = c->len;
  return EB200_OK;
}
template <int NL>
int rt_on(Ctx& c, const RtCurve<NL>& C, int op, size_t n, const uint8_t* k1, const uint8_t* p1, const uint8_t* k2, const uint8_t* p2,
          size_t klen, uint8_t* out, uint8_t* status) {
  int rc;
  const size_t pl = 2 * (size_t)C.len;
  if ((rc = grow(&c.d_in, &c.d_in_cap, n * (2 * klen + 3 * pl) + 1024))) return rc;
  if ((rc = grow(&c.d_status, &c.d_status_cap, n))) return rc;
  uint8_t *dk1 = c.d_in, *dk2 = dk1 + klen * n, *dp1 = dk2 + klen * n, *dp2 = dp1 + pl * n, *dout = dp2 + pl * n;
  cudaStream_t st = c.stream;
  CK(cudaEventRecord(c.ev[0], st));
  Seg seg[4] = {{dk1, k1, k1 ? klen * n : 0}, {dk2, k2, k2 ? klen * n : 0}, {dp1, p1, pl * n}, {dp2, p2, p2 ? pl * n : 0}};
  if ((rc = h2d(c, seg, 4, st))) return rc;
  CK(cudaEventRecord(c.ev[1], st));
  rt_curve_kernel<NL><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(op, n, C, k1 ? dk1 : nullptr, dp1, k2 ? dk2 : nullptr, p2 ? dp2 : nullptr,
                                                                   (u32)klen, dout, c.d_status);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c.ev[2], st));
  if (op != 3) CK(cudaMemcpyAsync(out, dout, pl * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(status, c.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(c.ev[3], st));
  return finish_timing(c, 1, true);
}
template <int NL>
int rt_call(const eb200_short_curve* cv, int op, size_t n, const uint8_t* k1, const uint8_t* p1, const uint8_t* k2, const uint8_t* p2,
            size_t klen, uint8_t* out, uint8_t* status) {
  RtCurve<NL> C;
  int rc = rt_make<NL>(C, cv);
  if (rc) return rc;
  const size_t pl = 2 * (size_t)cv->len;
  return run_sharded(n, [&](Ctx& c, size_t lo, size_t m) {
    return rt_on<NL>(c, C, op, m, k1 ? k1 + lo * klen : nullptr, p1 + lo * pl, k2 ? k2 + lo * klen : nullptr, p2 ? p2 + lo * pl : nullptr,
                     klen, out ? out + lo * pl : nullptr, status + lo);
  });
}
int rt_dispatch(const eb200_short_curve* cv, int op, size_t n, const uint8_t* k1, const uint8_t* p1, const uint8_t* k2, const uint8_t* p2,
                size_t klen, uint8_t* out, uint8_t* status) {
  if (!eb200_device_count()) return EB200_ERR_NOT_INIT;
  if (!cv || !cv->p || !cv->a || !cv->b || cv->len == 0 || cv->len > 72 || klen > 128) return EB200_ERR_ARG;
  if (n == 0) return EB200_OK;
  if (!p1 || !status || (op != 3 && !out)) return EB200_ERR_ARG;
  if (cv->len <= 32) return rt_call<8>(cv, op, n, k1, p1, k2, p2, klen, out, status);
  if (cv->len <= 48) return rt_call<12>(cv, op, n, k1, p1, k2, p2, klen, out, status);
  return rt_call<18>(cv, op, n, k1, p1, k2, p2, klen, out, status);
}
}  // namespace

extern "C" {

int eb200_curve_mul_batch(const eb200_short_curve* curve, size_t n, const uint8_t* k, size_t klen, const uint8_t* points_xy,
                          uint8_t* out_xy, uint8_t* status) {
  if (n && (!k || !klen)) return EB200_ERR_ARG;
  return rt_dispatch(curve, 0, n, k, points_xy, nullptr, nullptr, klen, out_xy, status);
}
int eb200_curve_mul_add_batch(const eb200_short_curve* curve, size_t n, const uint8_t* k1, const uint8_t* p1_xy, const uint8_t* k2,
                              const uint8_t* p2_xy, size_t klen, uint8_t* out_xy, uint8_t* status) {
  if (n && (!k1 || !k2 || !p2_xy || !klen)) return EB200_ERR_ARG;
  return rt_dispatch(curve, 0, n, k1, p1_xy, k2, p2_xy, klen, out_xy, status);
}
int eb200_curve_add_batch(const eb200_short_curve* curve, size_t n, const uint8_t* p1_xy, const uint8_t* p2_xy, uint8_t* out_xy,
                          uint8_t* status) {
  if (n && !p2_xy) return EB200_ERR_ARG;
  return rt_dispatch(curve, 1, n, nullptr, p1_xy, nullptr, p2_xy, 0, out_xy, status);
}
int eb200_curve_dbl_batch(const eb200_short_curve* curve, size_t n, const uint8_t* p_xy, uint8_t* out_xy, uint8_t* status) {
  return rt_dispatch(curve, 2, n, nullptr, p_xy, nullptr, nullptr, 0, out_xy, status);
}
int eb200_curve_validate_batch(const eb200_short_curve* curve, size_t n, const uint8_t* p_xy, uint8_t* status) {
  return rt_dispatch(curve, 3, n, nullptr, p_xy, nullptr, nullptr, 0, nullptr, status);
}

// ---- self-test hooks (first initialised device) ------------------------------------------------------
static Ctx* first_ctx() {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_ndev ? &g_ctx[g_devs[0]] : nullptr;
}

int eb200_selftest_fe(int curve, int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Ctx* cp = first_ctx();
  if (!cp) return EB200_ERR_NOT_INIT;
  if (!fe_len(curve)) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  Ctx& c = *cp;
  std::lock_guard<std::mutex> lk(c.mu);
  CK(cudaSetDevice(c.device));
  size_t bytes = n * fe_len(curve);
  u32 *da, *db, *dout;
  CK(cudaMalloc(&da, bytes)); CK(cudaMalloc(&db, bytes)); CK(cudaMalloc(&dout, bytes));
  // stream-ordered copies: a synchronous cudaMemcpy from pageable memory may return before its DMA lands,
  // and c.stream (non-blocking) does not wait for the legacy default stream
  CK(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, c.stream));
  CK(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, c.stream));
  unsigned nb = (unsigned)((n + 127) / 128);
  if (curve == EB200_CURVE_SECP256K1) k256_selftest_fe_kernel<<<nb, 128, 0, c.stream>>>(op, n, da, db, dout);
  else if (curve == EB200_CURVE_ED25519 || curve == EB200_CURVE_CURVE25519) f25_selftest_kernel<<<nb, 128, 0, c.stream>>>(op, n, da, db, dout);
  else {
#define EB_ST(C) (sw_selftest_fe_kernel<C><<<nb, 128, 0, c.stream>>>(op, n, da, db, dout), 0)
    (void)SW_DISPATCH(curve, EB_ST);
#undef EB_ST
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
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
  Ctx* cp = first_ctx();
  if (!cp) return EB200_ERR_NOT_INIT;
  int w, en, b;
  int rc = eb200_selftest_gtab_dims(curve, &w, &en, &b);
  if (rc) return rc;
  size_t words = (size_t)w * en * 2 * (fe_len(curve) / 4);
  if (!out || n_words < words) return EB200_ERR_ARG;
  Ctx& c = *cp;
  std::lock_guard<std::mutex> lk(c.mu);
  CK(cudaSetDevice(c.device));
  if ((rc = ensure_table(c, curve))) return rc;
  CK(cudaMemcpyAsync(out, c.gtab[curve], words * 4, cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  return EB200_OK;
}

}  // extern "C"
