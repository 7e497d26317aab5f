// eb200.cu -- CUDA kernels (sm_100a) and the C ABI of libelliptic_b200.so.
// See include/elliptic_b200.h for the boundary and ecdsa_k256_body.cuh for the
// algorithm.  No CPU fallback exists in this library by design.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include "../../include/elliptic_b200.h"
#include "ecdsa_k256_body.cuh"

using namespace eb;

// ---------------------------------------------------------------------------
// kernels
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

#ifndef EB_VERIFY_BLOCK
#define EB_VERIFY_BLOCK 128
#endif
#ifndef EB_VERIFY_MINBLOCKS
#define EB_VERIFY_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(EB_VERIFY_BLOCK, EB_VERIFY_MINBLOCKS)
k256_verify_kernel(size_t N, const uint8_t* __restrict__ pub, const uint8_t* __restrict__ r,
                   const u32* __restrict__ ws, const u32* __restrict__ gtab,
                   u32* __restrict__ qtab, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  status[i] = verify_item(i, N, pub, r, ws, gtab, qtab);
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
// context
namespace {
struct Ctx {
  bool ready = false;
  int device = -1;
  cudaStream_t stream = nullptr;
  u32* gtab_k256 = nullptr;
  // grow-on-demand device staging for the host-pointer API
  uint8_t* d_in = nullptr; size_t d_in_cap = 0;
  uint8_t* d_ws = nullptr; size_t d_ws_cap = 0;
  uint8_t* d_status = nullptr; size_t d_status_cap = 0;
  cudaEvent_t ev[6] = {};
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

struct WsLayout { size_t ws, scratch, qtab, total; };
WsLayout ws_layout(size_t n) {
  WsLayout L;
  L.ws = 0;
  L.scratch = align256(L.ws + (size_t)PREP_WORDS * n * 4);
  L.qtab = align256(L.scratch + (size_t)8 * n * 4);
  L.total = align256(L.qtab + (size_t)QTAB_WORDS * n * 4);
  return L;
}

int launch_k256_verify(size_t n, const uint8_t* d_e, const uint8_t* d_r, const uint8_t* d_s,
                       const uint8_t* d_pub, uint8_t* d_status, uint8_t* d_workspace,
                       cudaStream_t st, cudaEvent_t ev_main0, cudaEvent_t ev_main1) {
  if (n == 0) return EB200_OK;
  WsLayout L = ws_layout(n);
  u32* ws = (u32*)(d_workspace + L.ws);
  u32* scratch = (u32*)(d_workspace + L.scratch);
  u32* qtab = (u32*)(d_workspace + L.qtab);
  size_t T = (n + PREP_BATCH - 1) / PREP_BATCH;
  unsigned pb = (unsigned)((T + 127) / 128);
  k256_prep_kernel<<<pb, 128, 0, st>>>(n, d_e, d_r, d_s, ws, scratch);
  CK(cudaGetLastError());
  unsigned vb = (unsigned)((n + EB_VERIFY_BLOCK - 1) / EB_VERIFY_BLOCK);
  if (ev_main0) CK(cudaEventRecord(ev_main0, st));
  k256_verify_kernel<<<vb, EB_VERIFY_BLOCK, 0, st>>>(n, d_pub, d_r, ws, g.gtab_k256, qtab, d_status);
  CK(cudaGetLastError());
  if (ev_main1) CK(cudaEventRecord(ev_main1, st));
  return EB200_OK;
}
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
  for (int i = 0; i < 6; i++) if (!g.ev[i]) CK(cudaEventCreate(&g.ev[i]));
  if (g.gtab_k256) { cudaFree(g.gtab_k256); g.gtab_k256 = nullptr; }
  CK(cudaMalloc(&g.gtab_k256, (size_t)GTAB_WINDOWS * GTAB_ENTRIES * 16 * 4));
  k256_gtab_kernel<<<(unsigned)(((size_t)GTAB_WINDOWS * GTAB_ENTRIES + 127) / 128), 128, 0, g.stream>>>(g.gtab_k256);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g.stream));
  g.device = device;
  g.ready = true;
  return EB200_OK;
}

int eb200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g.ready) return EB200_OK;
  cudaSetDevice(g.device);
  cudaFree(g.gtab_k256); g.gtab_k256 = nullptr;
  cudaFree(g.d_in); g.d_in = nullptr; g.d_in_cap = 0;
  cudaFree(g.d_ws); g.d_ws = nullptr; g.d_ws_cap = 0;
  cudaFree(g.d_status); g.d_status = nullptr; g.d_status_cap = 0;
  for (int i = 0; i < 6; i++) if (g.ev[i]) { cudaEventDestroy(g.ev[i]); g.ev[i] = nullptr; }
  if (g.stream) { cudaStreamDestroy(g.stream); g.stream = nullptr; }
  g.ready = false;
  return EB200_OK;
}

int eb200_last_timing(eb200_timing* out) {
  if (!out) return EB200_ERR_ARG;
  if (g.dev_timing_pending) {
    // device-pointer call: the caller has synchronised its stream by now
    g.timing = eb200_timing{};
    if (cudaEventElapsedTime(&g.timing.kernel_ms, g.ev[1], g.ev[2]) != cudaSuccess) return EB200_ERR_CUDA;
    if (cudaEventElapsedTime(&g.timing.main_kernel_ms, g.ev[4], g.ev[5]) != cudaSuccess) return EB200_ERR_CUDA;
    g.timing.launches = 2;
    g.dev_timing_pending = false;
  }
  *out = g.timing;
  return EB200_OK;
}

size_t eb200_ecdsa_verify_workspace_bytes(int curve, size_t n) {
  if (curve != EB200_CURVE_SECP256K1) return 0;
  return ws_layout(n).total;
}

int eb200_ecdsa_verify_batch_dev(int curve, size_t n, const uint8_t* d_e, const uint8_t* d_r,
                                 const uint8_t* d_s, const uint8_t* d_pub, uint32_t pub_fmt,
                                 uint8_t* d_status, void* d_workspace, void* stream) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (curve != EB200_CURVE_SECP256K1) return EB200_ERR_UNSUPPORTED;
  if (pub_fmt != EB200_PUB_XY) return EB200_ERR_UNSUPPORTED;
  if (n && (!d_e || !d_r || !d_s || !d_pub || !d_status || !d_workspace)) return EB200_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;   // NULL is the CUDA default stream, as everywhere in CUDA
  // events on the caller's stream: eb200_last_timing() reports them once the stream has been synchronised
  CK(cudaEventRecord(g.ev[1], st));
  int rc = launch_k256_verify(n, d_e, d_r, d_s, d_pub, d_status, (uint8_t*)d_workspace, st, g.ev[4], g.ev[5]);
  if (rc) return rc;
  CK(cudaEventRecord(g.ev[2], st));
  g.dev_timing_pending = true;
  return EB200_OK;
}

int eb200_ecdsa_verify_batch(int curve, size_t n, const uint8_t* e, const uint8_t* r,
                             const uint8_t* s, const uint8_t* pub, uint32_t pub_fmt,
                             uint8_t* status) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (curve != EB200_CURVE_SECP256K1) return EB200_ERR_UNSUPPORTED;
  if (pub_fmt != EB200_PUB_XY) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  if (!e || !r || !s || !pub || !status) return EB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  const size_t len = 32;
  size_t in_bytes = n * (3 * len + 2 * len);
  int rc;
  if ((rc = grow(&g.d_in, &g.d_in_cap, in_bytes))) return rc;
  if ((rc = grow(&g.d_ws, &g.d_ws_cap, ws_layout(n).total))) return rc;
  if ((rc = grow(&g.d_status, &g.d_status_cap, n))) return rc;
  uint8_t* d_e = g.d_in;
  uint8_t* d_r = d_e + n * len;
  uint8_t* d_s = d_r + n * len;
  uint8_t* d_pub = d_s + n * len;
  cudaStream_t st = g.stream;
  CK(cudaEventRecord(g.ev[0], st));
  CK(cudaMemcpyAsync(d_e, e, n * len, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_r, r, n * len, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_s, s, n * len, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_pub, pub, n * 2 * len, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(g.ev[1], st));
  if ((rc = launch_k256_verify(n, d_e, d_r, d_s, d_pub, g.d_status, g.d_ws, st, g.ev[4], g.ev[5]))) return rc;
  CK(cudaEventRecord(g.ev[2], st));
  CK(cudaMemcpyAsync(status, g.d_status, n, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(g.ev[3], st));
  CK(cudaStreamSynchronize(st));
  g.dev_timing_pending = false;
  cudaEventElapsedTime(&g.timing.h2d_ms, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&g.timing.kernel_ms, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&g.timing.d2h_ms, g.ev[2], g.ev[3]);
  cudaEventElapsedTime(&g.timing.main_kernel_ms, g.ev[4], g.ev[5]);
  g.timing.launches = 2;
  return EB200_OK;
}

int eb200_selftest_fe(int curve, int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (curve != EB200_CURVE_SECP256K1) return EB200_ERR_UNSUPPORTED;
  if (n == 0) return EB200_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  u32 *da, *db, *dout;
  CK(cudaMalloc(&da, n * 32)); CK(cudaMalloc(&db, n * 32)); CK(cudaMalloc(&dout, n * 32));
  CK(cudaMemcpy(da, a, n * 32, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b, n * 32, cudaMemcpyHostToDevice));
  k256_selftest_fe_kernel<<<(unsigned)((n + 127) / 128), 128, 0, g.stream>>>(op, n, da, db, dout);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaMemcpy(out, dout, n * 32, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return EB200_OK;
}

int eb200_selftest_gtab_dims(int curve, int* windows, int* entries, int* wbits) {
  if (curve != EB200_CURVE_SECP256K1) return EB200_ERR_UNSUPPORTED;
  if (!windows || !entries || !wbits) return EB200_ERR_ARG;
  *windows = GTAB_WINDOWS; *entries = GTAB_ENTRIES; *wbits = GTAB_W;
  return EB200_OK;
}

int eb200_selftest_gtab(int curve, uint32_t* out, size_t n_words) {
  if (!g.ready) return EB200_ERR_NOT_INIT;
  if (curve != EB200_CURVE_SECP256K1) return EB200_ERR_UNSUPPORTED;
  size_t words = (size_t)GTAB_WINDOWS * GTAB_ENTRIES * 16;
  if (!out || n_words < words) return EB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  CK(cudaSetDevice(g.device));
  CK(cudaMemcpy(out, g.gtab_k256, words * 4, cudaMemcpyDeviceToHost));
  return EB200_OK;
}

}  // extern "C"
