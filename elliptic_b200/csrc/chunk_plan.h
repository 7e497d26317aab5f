// chunk_plan.h -- how a pipelined host call (eb200_ecdsa_verify_batch, eb200_eddsa_verify_batch*, eb200_x25519_*)
// cuts its batch into chunks.  Plain C++ (no CUDA): included by eb200.cu and by the host test harness.
#pragma once
#include <stddef.h>
#include <stdlib.h>
#ifndef EB_MAX_CHUNKS
#define EB_MAX_CHUNKS 16
#endif

// Chunk boundaries of a pipelined host call.  Equal chunks, except that the FIRST one is cut short (1/4 of a
// regular chunk): its host->device copy is the only one no kernel hides, so the shorter it is the sooner the GPU
// starts (EB200_LEAD=0 restores equal chunks; EB200_CHUNKS overrides the count).
struct ChunkPlan { int chunks; size_t lo[EB_MAX_CHUNKS + 2]; size_t max_m; };
static inline ChunkPlan make_plan(size_t n) {
  int ch = 1;
  if (n >= ((size_t)1 << 18)) ch = 4;         // 2^18-item chunks keep the grid tail small (r01: 8 chunks cost 10%)
  if (n >= ((size_t)1 << 22)) ch = EB_MAX_CHUNKS;
  if (const char* ev = getenv("EB200_CHUNKS")) { int k = atoi(ev); if (k >= 1 && k <= EB_MAX_CHUNKS) ch = k; }   // tuning knob
  bool lead = ch > 1 && ch < EB_MAX_CHUNKS;
  if (const char* ev = getenv("EB200_LEAD")) lead = lead && atoi(ev) != 0;
  ChunkPlan P;
  size_t per = (n + ch - 1) / ch;
  per = (per + 127) & ~(size_t)127;
  int k = 0;
  size_t pos = 0;
  P.lo[0] = 0;
  if (lead) {
    size_t first = ((per / 4) + 127) & ~(size_t)127;
    if (first < n) {
      pos = first;
      P.lo[++k] = pos;
      per = (n - first + ch - 1) / ch;
      per = (per + 127) & ~(size_t)127;
    }
  }
  while (pos < n) {
    pos = pos + per < n ? pos + per : n;
    P.lo[++k] = pos;
  }
  P.chunks = k;
  P.max_m = 0;
  for (int i = 0; i < k; i++) if (P.lo[i + 1] - P.lo[i] > P.max_m) P.max_m = P.lo[i + 1] - P.lo[i];
  return P;
}

