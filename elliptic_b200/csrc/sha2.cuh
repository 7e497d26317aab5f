// sha2.cuh -- SHA-256 / SHA-512 / HMAC-SHA-256 for the wire-format rows of the path:
//   * EdDSA  h = SHA512(R || A || M) mod n          (lib/elliptic/eddsa/index.js:65-70; hash.js
//                                                     sha512 vendored at dist/elliptic.js:8023-8354)
//   * RFC 6979 nonces via HMAC-DRBG(SHA-256)         (ec/index.js:142-157; hmac-drbg dist:8686-8800;
//                                                     hash.js sha256 dist:7879-7985, hmac dist:7612-7660)
// FIPS 180-4 restated from the standard; one thread hashes one message (the messages on this path
// are tens of bytes, so there is nothing to parallelise inside a hash).
#pragma once
#include "limbs.cuh"

namespace eb {

// ----------------------------------------------------------------------------- SHA-512
struct sha512_ctx { u64 h[8]; uint8_t buf[128]; u32 fill; u64 total; };

EB_HD u64 rotr64(u64 x, int n) { return (x >> n) | (x << (64 - n)); }

#define EB_SHA512_K_INIT       0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL, 0x3956c25bf348b538ULL, \
      0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL, 0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, \
      0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL, 0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, \
      0xc19bf174cf692694ULL, 0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL, \
      0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL, 0x983e5152ee66dfabULL, \
      0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL, 0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL, \
      0x06ca6351e003826fULL, 0x142929670a0e6e70ULL, 0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, \
      0x53380d139d95b3dfULL, 0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL, \
      0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL, 0xd192e819d6ef5218ULL, \
      0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL, 0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, \
      0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL, 0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL, \
      0x682e6ff3d6b2b8a3ULL, 0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL, \
      0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL, 0xca273eceea26619cULL, \
      0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL, 0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, \
      0x113f9804bef90daeULL, 0x1b710b35131c471bULL, 0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, \
      0x431d67c49c100d4cULL, 0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL
#if defined(__CUDACC__)
__device__ __constant__ u64 SHA512_K_DEV[80] = {EB_SHA512_K_INIT};
#endif
static const u64 SHA512_K_HOST[80] = {EB_SHA512_K_INIT};
EB_HD u64 sha512_k(int i) {
#if defined(__CUDA_ARCH__)
  return SHA512_K_DEV[i];
#else
  return SHA512_K_HOST[i];
#endif
}

EB_HD void sha512_block(u64* h, const uint8_t* p) {
  u64 w[16];
  for (int i = 0; i < 16; i++) {
    u64 v = 0;
    for (int k = 0; k < 8; k++) v = (v << 8) | p[8 * i + k];
    w[i] = v;
  }
  u64 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 80; i++) {
    u64 wi;
    if (i < 16) wi = w[i];
    else {
      u64 w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      u64 s0 = rotr64(w15, 1) ^ rotr64(w15, 8) ^ (w15 >> 7);
      u64 s1 = rotr64(w2, 19) ^ rotr64(w2, 61) ^ (w2 >> 6);
      wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
      w[i & 15] = wi;
    }
    u64 S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
    u64 ch = (e & f) ^ (~e & g);
    u64 t1 = hh + S1 + ch + sha512_k(i) + wi;
    u64 S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
    u64 mj = (a & b) ^ (a & c) ^ (b & c);
    u64 t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

EB_HD void sha512_init(sha512_ctx* c) {
  const u64 iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                     0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
  for (int i = 0; i < 8; i++) c->h[i] = iv[i];
  c->fill = 0; c->total = 0;
}
EB_HD void sha512_update(sha512_ctx* c, const uint8_t* p, size_t n) {
  c->total += n;
  for (size_t i = 0; i < n; i++) {
    c->buf[c->fill++] = p[i];
    if (c->fill == 128) { sha512_block(c->h, c->buf); c->fill = 0; }
  }
}
EB_HD void sha512_final(sha512_ctx* c, uint8_t* out64) {
  u64 bits = c->total * 8;
  c->buf[c->fill++] = 0x80;
  if (c->fill > 112) { while (c->fill < 128) c->buf[c->fill++] = 0; sha512_block(c->h, c->buf); c->fill = 0; }
  while (c->fill < 120) c->buf[c->fill++] = 0;
  for (int k = 0; k < 8; k++) c->buf[120 + k] = (uint8_t)(bits >> (56 - 8 * k));
  sha512_block(c->h, c->buf);
  for (int i = 0; i < 8; i++)
    for (int k = 0; k < 8; k++) out64[8 * i + k] = (uint8_t)(c->h[i] >> (56 - 8 * k));
}

// ----------------------------------------------------------------------------- SHA-256
struct sha256_ctx { u32 h[8]; uint8_t buf[64]; u32 fill; u64 total; };

EB_HD u32 rotr32(u32 x, int n) { return (x >> n) | (x << (32 - n)); }

#define EB_SHA256_K_INIT       0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, \
      0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, \
      0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, \
      0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, \
      0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, \
      0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, \
      0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, \
      0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u
#if defined(__CUDACC__)
__device__ __constant__ u32 SHA256_K_DEV[64] = {EB_SHA256_K_INIT};
#endif
static const u32 SHA256_K_HOST[64] = {EB_SHA256_K_INIT};
EB_HD u32 sha256_k(int i) {
#if defined(__CUDA_ARCH__)
  return SHA256_K_DEV[i];
#else
  return SHA256_K_HOST[i];
#endif
}
EB_HD void sha256_block(u32* h, const uint8_t* p) {
  u32 w[16];
  for (int i = 0; i < 16; i++) w[i] = ((u32)p[4 * i] << 24) | ((u32)p[4 * i + 1] << 16) | ((u32)p[4 * i + 2] << 8) | p[4 * i + 3];
  u32 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; i++) {
    u32 wi;
    if (i < 16) wi = w[i];
    else {
      u32 w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      u32 s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      u32 s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
      w[i & 15] = wi;
    }
    u32 S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    u32 ch = (e & f) ^ (~e & g);
    u32 t1 = hh + S1 + ch + sha256_k(i) + wi;
    u32 S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    u32 mj = (a & b) ^ (a & c) ^ (b & c);
    u32 t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
EB_HD void sha256_init(sha256_ctx* c) {
  const u32 iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
  for (int i = 0; i < 8; i++) c->h[i] = iv[i];
  c->fill = 0; c->total = 0;
}
EB_HD void sha256_update(sha256_ctx* c, const uint8_t* p, size_t n) {
  c->total += n;
  for (size_t i = 0; i < n; i++) {
    c->buf[c->fill++] = p[i];
    if (c->fill == 64) { sha256_block(c->h, c->buf); c->fill = 0; }
  }
}
EB_HD void sha256_final(sha256_ctx* c, uint8_t* out32) {
  u64 bits = c->total * 8;
  c->buf[c->fill++] = 0x80;
  if (c->fill > 56) { while (c->fill < 64) c->buf[c->fill++] = 0; sha256_block(c->h, c->buf); c->fill = 0; }
  while (c->fill < 56) c->buf[c->fill++] = 0;
  for (int k = 0; k < 8; k++) c->buf[56 + k] = (uint8_t)(bits >> (56 - 8 * k));
  sha256_block(c->h, c->buf);
  for (int i = 0; i < 8; i++)
    for (int k = 0; k < 4; k++) out32[4 * i + k] = (uint8_t)(c->h[i] >> (24 - 8 * k));
}

// HMAC-SHA-256 with a 32-byte key over up to three pieces
EB_HD void hmac_sha256(const uint8_t* key32, const uint8_t* a, size_t na, const uint8_t* b, size_t nb,
                       const uint8_t* c, size_t nc, uint8_t* out32) {
  uint8_t pad[64];
  sha256_ctx ctx;
  for (int i = 0; i < 64; i++) pad[i] = (i < 32 ? key32[i] : 0) ^ 0x36;
  sha256_init(&ctx);
  sha256_update(&ctx, pad, 64);
  if (na) sha256_update(&ctx, a, na);
  if (nb) sha256_update(&ctx, b, nb);
  if (nc) sha256_update(&ctx, c, nc);
  uint8_t inner[32];
  sha256_final(&ctx, inner);
  for (int i = 0; i < 64; i++) pad[i] = (i < 32 ? key32[i] : 0) ^ 0x5c;
  sha256_init(&ctx);
  sha256_update(&ctx, pad, 64);
  sha256_update(&ctx, inner, 32);
  sha256_final(&ctx, out32);
}

}  // namespace eb
