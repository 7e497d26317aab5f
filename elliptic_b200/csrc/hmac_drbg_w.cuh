// hmac_drbg_w.cuh -- HmacDRBG (hmac-drbg 1.0.1, dist/elliptic.js:8686-8800) over SHA-256 or SHA-384, word-oriented
// and entirely in registers, for the one shape EC.prototype.sign uses it in (lib/elliptic/ec/index.js:135-160):
// entropy = the private key, nonce = the truncated message, both exactly one digest long (32 bytes with SHA-256
// on a 256-bit curve, 48 bytes with SHA-384 on p384), no personalisation string, generate(digest length).
//
// HMAC pad states are computed once per key and reused by the calls that share it; every message fed to the
// hash is built directly as big-endian words (the seed material sits one byte off the word grid behind the
// 0x00 / 0x01 separator, hence the funnel shifts).
#pragma once
#include "sha2.cuh"

namespace eb {

struct Sha256W {
  typedef u32 W;
  static constexpr int D = 8;              // digest words
  static constexpr int WB = 32;            // bits per word
  static EB_HD void iv(W* s) {
    const W v[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    for (int i = 0; i < 8; i++) s[i] = v[i];
  }
  static EB_HD void compress(W* st, const W* win) {
    W w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = win[i];
    W a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
      W wi;
      if (i < 16) wi = w[i];
      else {
        W w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
        W s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
        W s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
        wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        w[i & 15] = wi;
      }
      W S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
      W ch = (e & f) ^ (~e & g);
      W t1 = h + S1 + ch + sha256_k(i) + wi;
      W S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
      W mj = (a & b) ^ (a & c) ^ (b & c);
      h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + S0 + mj;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
  }
};

struct Sha384W {
  typedef u64 W;
  static constexpr int D = 6;
  static constexpr int WB = 64;
  static EB_HD void iv(W* s) {
    const W v[8] = {0xcbbb9d5dc1059ed8ULL, 0x629a292a367cd507ULL, 0x9159015a3070dd17ULL, 0x152fecd8f70e5939ULL,
                    0x67332667ffc00b31ULL, 0x8eb44a8768581511ULL, 0xdb0c2e0d64f98fa7ULL, 0x47b5481dbefa4fa4ULL};
    for (int i = 0; i < 8; i++) s[i] = v[i];
  }
  static EB_HD void compress(W* st, const W* win) {
    W w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = win[i];
    W a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 8
    for (int i = 0; i < 80; i++) {
      W wi;
      if (i < 16) wi = w[i];
      else {
        W w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
        W s0 = rotr64(w15, 1) ^ rotr64(w15, 8) ^ (w15 >> 7);
        W s1 = rotr64(w2, 19) ^ rotr64(w2, 61) ^ (w2 >> 6);
        wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        w[i & 15] = wi;
      }
      W S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
      W ch = (e & f) ^ (~e & g);
      W t1 = h + S1 + ch + sha512_k(i) + wi;
      W S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
      W mj = (a & b) ^ (a & c) ^ (b & c);
      h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + S0 + mj;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
  }
};

template <class H>
struct HmacDrbgW {
  typedef typename H::W W;
  static constexpr int D = H::D;
  static constexpr int WB = H::WB;
  static constexpr int DBYTES = D * WB / 8;          // digest = key = V length in bytes
  static constexpr int BBYTES = 16 * WB / 8;         // block length in bytes
  W kin[8], kout[8];                                  // hash states after the ipad / opad block of the current K
  W V[D];
  bool first;

  EB_HD void set_key(const W* key) {
    W w[16];
    const W ipad = (W)0x3636363636363636ULL, opad = (W)0x5c5c5c5c5c5c5c5cULL;
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = (i < D ? key[i] : (W)0) ^ ipad;
    H::iv(kin);
    H::compress(kin, w);
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = (i < D ? key[i] : (W)0) ^ opad;
    H::iv(kout);
    H::compress(kout, w);
  }
  // finishes an HMAC: `inner` is the inner hash state after all message blocks
  EB_HD void outer(const W* inner, W* out) const {
    W w[16], st[8];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
#pragma unroll
    for (int i = 0; i < D; i++) w[i] = inner[i];
    w[D] = (W)0x80 << (WB - 8);
    w[15] = (W)((BBYTES + DBYTES) * 8);
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = kout[i];
    H::compress(st, w);
#pragma unroll
    for (int i = 0; i < D; i++) out[i] = st[i];
  }
  // HMAC(K, V) and HMAC(K, V || sep): one inner message block
  EB_HD void mac_v(W* out, bool with_sep, W sep) const {
    W w[16], st[8];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
#pragma unroll
    for (int i = 0; i < D; i++) w[i] = V[i];
    if (with_sep) w[D] = (sep << (WB - 8)) | ((W)0x80 << (WB - 16));
    else w[D] = (W)0x80 << (WB - 8);
    w[15] = (W)((BBYTES + DBYTES + (with_sep ? 1 : 0)) * 8);
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = kin[i];
    H::compress(st, w);
    outer(st, out);
  }
  // HMAC(K, V || sep || a || b), a and b one digest long each: 3 D + 1 words in two blocks
  EB_HD void mac_v_sep_seed(W sep, const W* a, const W* b, W* out) const {
    W m[32], st[8];
#pragma unroll
    for (int i = 0; i < 32; i++) m[i] = 0;
#pragma unroll
    for (int i = 0; i < D; i++) m[i] = V[i];
    m[D] = (sep << (WB - 8)) | (a[0] >> 8);
#pragma unroll
    for (int i = 1; i < D; i++) m[D + i] = (a[i - 1] << (WB - 8)) | (a[i] >> 8);
    m[2 * D] = (a[D - 1] << (WB - 8)) | (b[0] >> 8);
#pragma unroll
    for (int i = 1; i < D; i++) m[2 * D + i] = (b[i - 1] << (WB - 8)) | (b[i] >> 8);
    m[3 * D] = (b[D - 1] << (WB - 8)) | ((W)0x80 << (WB - 16));
    m[31] = (W)((BBYTES + 3 * DBYTES + 1) * 8);
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = kin[i];
    H::compress(st, m);
    H::compress(st, m + 16);
    outer(st, out);
  }

  // new HmacDRBG({hash, entropy: priv, nonce: msg})   (dist:8692-8733)
  EB_HD void init(const W* priv, const W* msg) {
    W K[D];
#pragma unroll
    for (int i = 0; i < D; i++) { K[i] = 0; V[i] = (W)0x0101010101010101ULL; }
    set_key(K);
    mac_v_sep_seed(0x00, priv, msg, K);
    set_key(K);
    mac_v(V, false, 0);
    mac_v_sep_seed(0x01, priv, msg, K);
    set_key(K);
    mac_v(V, false, 0);
    first = true;
  }
  // generate(digest length)   (dist:8771-8797); the trailing _update() is deferred to the next call
  EB_HD void generate(W* out) {
    if (!first) {
      W K[D];
      mac_v(K, true, 0x00);
      set_key(K);
      mac_v(V, false, 0);
    }
    first = false;
    mac_v(V, false, 0);
#pragma unroll
    for (int i = 0; i < D; i++) out[i] = V[i];
  }
};

}  // namespace eb

// ---------------------------------------------------------------------------------------------------------------
// The same generator for key / message lengths that are not one digest long (p192 and p224 with SHA-256, p521 with
// SHA-512; curves.js:43-71,109-134): messages are assembled as bytes and hashed block by block.  Not a hot path.
namespace eb {

struct Sha512W : Sha384W {
  static constexpr int D = 8;
  static EB_HD void iv(W* s) {
    const W v[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                    0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    for (int i = 0; i < 8; i++) s[i] = v[i];
  }
};

template <class H>
struct HashStreamW {                      // Merkle-Damgard streaming over H::compress
  typedef typename H::W W;
  static constexpr int WBY = H::WB / 8, BB = 16 * WBY;
  W st[8];
  uint8_t buf[BB];
  int fill;
  u64 total;
  EB_HD void block() {
    W w[16];
    for (int i = 0; i < 16; i++) {
      W v = 0;
      for (int k = 0; k < WBY; k++) v = (v << 8) | buf[WBY * i + k];
      w[i] = v;
    }
    H::compress(st, w);
    fill = 0;
  }
  EB_HD void start(const W* state, u64 already) { for (int i = 0; i < 8; i++) st[i] = state[i]; fill = 0; total = already; }
  EB_HD void absorb(const uint8_t* p, int n) {
    total += (u64)n;
    for (int i = 0; i < n; i++) { buf[fill++] = p[i]; if (fill == BB) block(); }
  }
  EB_HD void finish(uint8_t* out, int nbytes) {
    u64 bits = total * 8;
    buf[fill++] = 0x80;
    if (fill > BB - 2 * WBY) { while (fill < BB) buf[fill++] = 0; block(); }
    while (fill < BB - 8) buf[fill++] = 0;                    // the high half of a 128-bit length is zero
    for (int k = 0; k < 8; k++) buf[BB - 8 + k] = (uint8_t)(bits >> (56 - 8 * k));
    fill = BB;
    block();
    for (int i = 0; i < nbytes; i++) out[i] = (uint8_t)(st[i / WBY] >> (8 * (WBY - 1 - i % WBY)));
  }
};

template <class H>
struct HmacDrbgB {
  typedef typename H::W W;
  static constexpr int DB = H::D * H::WB / 8;                 // digest bytes
  static constexpr int BB = 16 * H::WB / 8;
  W kin[8], kout[8];
  uint8_t V[DB];
  bool first;

  EB_HD void set_key(const uint8_t* key) {
    uint8_t pad[BB];
    HashStreamW<H> s;
    W iv[8];
    H::iv(iv);
    for (int i = 0; i < BB; i++) pad[i] = (i < DB ? key[i] : 0) ^ 0x36;
    s.start(iv, 0); s.absorb(pad, BB);
    for (int i = 0; i < 8; i++) kin[i] = s.st[i];
    for (int i = 0; i < BB; i++) pad[i] = (i < DB ? key[i] : 0) ^ 0x5c;
    s.start(iv, 0); s.absorb(pad, BB);
    for (int i = 0; i < 8; i++) kout[i] = s.st[i];
  }
  // HMAC(K, V || [sep] || a || b || c)
  EB_HD void mac(bool with_sep, uint8_t sep, const uint8_t* a, int na, const uint8_t* b, int nb, uint8_t* out,
                 const uint8_t* c = nullptr, int nc = 0) const {
    HashStreamW<H> s;
    uint8_t inner[DB];
    s.start(kin, BB);
    s.absorb(V, DB);
    if (with_sep) s.absorb(&sep, 1);
    if (na) s.absorb(a, na);
    if (nb) s.absorb(b, nb);
    if (nc) s.absorb(c, nc);
    s.finish(inner, DB);
    s.start(kout, BB);
    s.absorb(inner, DB);
    s.finish(out, DB);
  }
  // seed = entropy || nonce || pers  (HmacDRBG._init, dist:8719-8733)
  EB_HD void init(const uint8_t* entropy, int ne, const uint8_t* nonce, int nn, const uint8_t* pers = nullptr, int np = 0) {
    uint8_t K[DB];
    for (int i = 0; i < DB; i++) { K[i] = 0x00; V[i] = 0x01; }
    set_key(K);
    mac(true, 0x00, entropy, ne, nonce, nn, K, pers, np);
    set_key(K);
    mac(false, 0, nullptr, 0, nullptr, 0, V);
    mac(true, 0x01, entropy, ne, nonce, nn, K, pers, np);
    set_key(K);
    mac(false, 0, nullptr, 0, nullptr, 0, V);
    first = true;
  }
  EB_HD void generate(uint8_t* out, int len) {
    if (!first) {
      uint8_t K[DB];
      mac(true, 0x00, nullptr, 0, nullptr, 0, K);
      set_key(K);
      mac(false, 0, nullptr, 0, nullptr, 0, V);
    }
    first = false;
    int got = 0;
    while (got < len) {
      mac(false, 0, nullptr, 0, nullptr, 0, V);
      for (int i = 0; i < DB && got < len; i++) out[got++] = V[i];
    }
  }
};

}  // namespace eb
