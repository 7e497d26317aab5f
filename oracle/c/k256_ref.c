/* ORACLE (test infrastructure / CPU baseline only -- never linked into the product).
 *
 * C restatement of the reference's secp256k1 ECDSA-verify path, algorithm for
 * algorithm (so that it is a fair "reference CPU path" stand-in):
 *   EC.verify                      lib/elliptic/ec/index.js:188-229
 *   Point.jmulAdd                  lib/elliptic/curve/short.js:443-450
 *   ShortCurve._endoWnafMulAdd     short.js:218-249   (GLV, sign fix-ups)
 *   ShortCurve._endoSplit          short.js:168-185   (exact divRound, dist:6387-6404)
 *   BaseCurve._wnafMulAdd          lib/elliptic/curve/base.js:128-253
 *   utils.getNAF / getJSF          lib/elliptic/utils.js:15-101
 *   JPoint.add/mixedAdd/dbl/eqXToP short.js:532-603, 668-737, 908-925
 *   Point.add (affine), _getBeta   short.js:365-392, 282-310
 * Field elements are canonical residues in 4 x 64-bit limbs (bn.js keeps
 * 10 x 26-bit limbs; representation is not observable).  Off-curve public keys
 * are processed exactly as the reference processes them (no validation).
 * Parity: checked against oracle/ref_py on golden vectors (tests/test_oracle_c.py).
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

/* ---------------------------------------------------------------- 256-bit helpers */
typedef struct { u64 v[4]; } u256;

static const u256 P_ = {{0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
static const u256 N_ = {{0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};

static int u_is_zero(const u256* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static int u_cmp(const u256* a, const u256* b) {
  for (int i = 3; i >= 0; i--) { if (a->v[i] < b->v[i]) return -1; if (a->v[i] > b->v[i]) return 1; }
  return 0;
}
static u64 u_add(u256* r, const u256* a, const u256* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; r->v[i] = (u64)c; c >>= 64; }
  return (u64)c;
}
static u64 u_sub(u256* r, const u256* a, const u256* b) {
  u64 bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a->v[i] - b->v[i] - bw;
    r->v[i] = (u64)t; bw = (u64)(t >> 64) & 1;
  }
  return bw;
}
static void u_shr1(u256* a, u64 top) {
  for (int i = 0; i < 3; i++) a->v[i] = (a->v[i] >> 1) | (a->v[i + 1] << 63);
  a->v[3] = (a->v[3] >> 1) | (top << 63);
}
static void u_from_be(u256* r, const uint8_t* p) {
  for (int i = 0; i < 4; i++) {
    u64 w = 0;
    for (int k = 0; k < 8; k++) w = (w << 8) | p[8 * (3 - i) + k];
    r->v[i] = w;
  }
}
static int u_bitlen(const u256* a) {
  for (int i = 3; i >= 0; i--) if (a->v[i]) return 64 * i + 64 - __builtin_clzll(a->v[i]);
  return 0;
}

/* ---------------------------------------------------------------- field mod p (Red over K256) */
typedef u256 fe;
static __thread unsigned long g_count_mul; /* per-thread op counter */

static void fe_reduce_once(fe* a) { if (u_cmp(a, &P_) >= 0) u_sub(a, a, &P_); }
static void fe_set(fe* r, const u256* a) { *r = *a; fe_reduce_once(r); }  /* toRed: umod p (inputs < 2^256 < 2p) */
static void fe_add(fe* r, const fe* a, const fe* b) {
  u64 c = u_add(r, a, b);
  if (c || u_cmp(r, &P_) >= 0) u_sub(r, r, &P_);
}
static void fe_sub(fe* r, const fe* a, const fe* b) { if (u_sub(r, a, b)) u_add(r, r, &P_); }
static void fe_neg(fe* r, const fe* a) { if (u_is_zero(a)) *r = *a; else u_sub(r, &P_, a); }
static void fe_mul(fe* r, const fe* a, const fe* b) {
  u64 t[8] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a->v[j] * b->v[i] + t[i + j]; t[i + j] = (u64)c; c >>= 64; }
    t[i + 4] = (u64)c;
  }
  /* fold: 2^256 = 0x1000003D1 (mod p) */
  const u64 K = 0x1000003D1ULL;
  u128 c = 0; u64 lo[5];
  for (int i = 0; i < 4; i++) { c += (u128)t[4 + i] * K + t[i]; lo[i] = (u64)c; c >>= 64; }
  lo[4] = (u64)c;
  c = (u128)lo[4] * K + lo[0]; r->v[0] = (u64)c; c >>= 64;
  for (int i = 1; i < 4; i++) { c += lo[i]; r->v[i] = (u64)c; c >>= 64; }
  if ((u64)c) { /* wrapped once more */
    u128 d = (u128)r->v[0] + K; r->v[0] = (u64)d; d >>= 64;
    for (int i = 1; i < 4 && (u64)d; i++) { d += r->v[i]; r->v[i] = (u64)d; d >>= 64; }
  }
  fe_reduce_once(r);
  g_count_mul++;
}
static void fe_sqr(fe* r, const fe* a) { fe_mul(r, a, a); }

/* binary modular inverse (value-equivalent to bn.js _invmp, dist:6518-6582; inv(0) = 0) */
static void mod_inv(u256* r, const u256* a, const u256* m) {
  u256 u = *a, v = *m, x1 = {{1, 0, 0, 0}}, x2 = {{0, 0, 0, 0}}, one = {{1, 0, 0, 0}};
  if (u_is_zero(a)) { *r = x2; return; }
  while (u_cmp(&u, &one) != 0 && u_cmp(&v, &one) != 0) {
    while ((u.v[0] & 1) == 0) {
      u_shr1(&u, 0);
      if (x1.v[0] & 1) { u64 c = u_add(&x1, &x1, m); u_shr1(&x1, c); } else u_shr1(&x1, 0);
    }
    while ((v.v[0] & 1) == 0) {
      u_shr1(&v, 0);
      if (x2.v[0] & 1) { u64 c = u_add(&x2, &x2, m); u_shr1(&x2, c); } else u_shr1(&x2, 0);
    }
    if (u_cmp(&u, &v) >= 0) { u_sub(&u, &u, &v); if (u_sub(&x1, &x1, &x2)) u_add(&x1, &x1, m); }
    else { u_sub(&v, &v, &u); if (u_sub(&x2, &x2, &x1)) u_add(&x2, &x2, m); }
  }
  *r = (u_cmp(&u, &one) == 0) ? x1 : x2;
}
static void fe_inv(fe* r, const fe* a) { mod_inv(r, a, &P_); }

/* ---------------------------------------------------------------- scalars mod n */
/* r = a*b mod n via 512-bit product and bit-serial reduction of the high half using
 * 2^256 = 2^256 - n (mod n) folding (n = 2^256 - c, c < 2^129). */
static const u64 NC[3] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 1ULL}; /* 2^256 - n */
static void sc_mul(u256* r, const u256* a, const u256* b) {
  u64 t[9] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a->v[j] * b->v[i] + t[i + j]; t[i + j] = (u64)c; c >>= 64; }
    t[i + 4] = (u64)c;
  }
  /* fold high limbs three times: hi * c added to lo */
  for (int round = 0; round < 3; round++) {
    u64 hi[5] = {t[4], t[5], t[6], t[7], t[8]};
    u64 acc[9] = {t[0], t[1], t[2], t[3], 0, 0, 0, 0, 0};
    for (int i = 0; i < 5; i++) {
      if (!hi[i]) continue;
      u128 c = 0;
      for (int j = 0; j < 3; j++) { c += (u128)hi[i] * NC[j] + acc[i + j]; acc[i + j] = (u64)c; c >>= 64; }
      for (int k = i + 3; k < 9 && (u64)c; k++) { c += acc[k]; acc[k] = (u64)c; c >>= 64; }
    }
    memcpy(t, acc, sizeof acc);
  }
  u256 x = {{t[0], t[1], t[2], t[3]}};
  /* t[4] is now 0 or tiny: finish by subtraction */
  while (t[4] || u_cmp(&x, &N_) >= 0) { if (u_sub(&x, &x, &N_)) t[4]--; }
  *r = x;
}

/* ---------------------------------------------------------------- points */
typedef struct { fe x, y; int inf; } apoint;           /* affine Point, short.js:251-271 */
typedef struct { fe x, y, z; int zone; } jpoint;       /* JPoint, short.js:490-509 */

static const fe FE_ONE = {{1, 0, 0, 0}};
static fe BETA;  /* curves.js:187 */

static void jp_inf(jpoint* r) { r->x = FE_ONE; r->y = FE_ONE; memset(&r->z, 0, sizeof(fe)); r->zone = 0; }
static int jp_is_inf(const jpoint* p) { return u_is_zero(&p->z); }
static void ap_to_j(jpoint* r, const apoint* p) {
  if (p->inf) { jp_inf(r); return; }
  r->x = p->x; r->y = p->y; r->z = FE_ONE; r->zone = 1;
}
static void ap_neg(apoint* r, const apoint* p) { *r = *p; if (!p->inf) fe_neg(&r->y, &p->y); }
static void jp_neg(jpoint* r, const jpoint* p) { *r = *p; fe_neg(&r->y, &p->y); }  /* keeps zOne: same z object in JS */

/* JPoint._zeroDbl, short.js:668-737 (both Z branches give identical coordinates; the
 * zOne branch is kept because it costs 1M+5S instead of 2M+5S, as in the reference). */
static void jp_dbl(jpoint* r, const jpoint* p) {
  if (jp_is_inf(p)) { *r = *p; return; }
  fe a, b, c, d, e, f, t, c8, nx, ny, nz;
  fe_sqr(&a, &p->x); fe_sqr(&b, &p->y); fe_sqr(&c, &b);
  fe_add(&t, &p->x, &b); fe_sqr(&t, &t); fe_sub(&t, &t, &a); fe_sub(&t, &t, &c);
  fe_add(&d, &t, &t);
  fe_add(&e, &a, &a); fe_add(&e, &e, &a);
  fe_sqr(&f, &e);
  fe_add(&c8, &c, &c); fe_add(&c8, &c8, &c8); fe_add(&c8, &c8, &c8);
  fe_sub(&nx, &f, &d); fe_sub(&nx, &nx, &d);
  fe_sub(&t, &d, &nx); fe_mul(&ny, &e, &t); fe_sub(&ny, &ny, &c8);
  if (p->zone) fe_add(&nz, &p->y, &p->y);
  else { fe_mul(&nz, &p->y, &p->z); fe_add(&nz, &nz, &nz); }
  r->x = nx; r->y = ny; r->z = nz; r->zone = 0;
}
static void jp_dblp(jpoint* r, const jpoint* p, int k) {
  *r = *p;
  if (k == 0 || jp_is_inf(p)) return;
  for (int i = 0; i < k; i++) { jpoint t; jp_dbl(&t, r); *r = t; }
}

/* JPoint.mixedAdd, short.js:569-603 */
static void jp_madd(jpoint* r, const jpoint* a, const apoint* p) {
  if (jp_is_inf(a)) { ap_to_j(r, p); return; }
  if (p->inf) { *r = *a; return; }
  fe z2, u2, s2, h, rr, h2, h3, v, nx, ny, nz, t;
  fe_sqr(&z2, &a->z);
  fe_mul(&u2, &p->x, &z2);
  fe_mul(&s2, &p->y, &z2); fe_mul(&s2, &s2, &a->z);
  fe_sub(&h, &a->x, &u2);
  fe_sub(&rr, &a->y, &s2);
  if (u_is_zero(&h)) {
    if (!u_is_zero(&rr)) { jp_inf(r); return; }
    jp_dbl(r, a); return;
  }
  fe_sqr(&h2, &h); fe_mul(&h3, &h2, &h); fe_mul(&v, &a->x, &h2);
  fe_sqr(&nx, &rr); fe_add(&nx, &nx, &h3); fe_sub(&nx, &nx, &v); fe_sub(&nx, &nx, &v);
  fe_sub(&t, &v, &nx); fe_mul(&ny, &rr, &t); fe_mul(&t, &a->y, &h3); fe_sub(&ny, &ny, &t);
  fe_mul(&nz, &a->z, &h);
  r->x = nx; r->y = ny; r->z = nz; r->zone = 0;
}

/* JPoint.add, short.js:532-567 */
static void jp_add(jpoint* r, const jpoint* a, const jpoint* b) {
  if (jp_is_inf(a)) { *r = *b; return; }
  if (jp_is_inf(b)) { *r = *a; return; }
  fe pz2, z2, u1, u2, s1, s2, h, rr, h2, h3, v, nx, ny, nz, t;
  fe_sqr(&pz2, &b->z); fe_sqr(&z2, &a->z);
  fe_mul(&u1, &a->x, &pz2); fe_mul(&u2, &b->x, &z2);
  fe_mul(&t, &pz2, &b->z); fe_mul(&s1, &a->y, &t);
  fe_mul(&t, &z2, &a->z); fe_mul(&s2, &b->y, &t);
  fe_sub(&h, &u1, &u2); fe_sub(&rr, &s1, &s2);
  if (u_is_zero(&h)) {
    if (!u_is_zero(&rr)) { jp_inf(r); return; }
    jp_dbl(r, a); return;
  }
  fe_sqr(&h2, &h); fe_mul(&h3, &h2, &h); fe_mul(&v, &u1, &h2);
  fe_sqr(&nx, &rr); fe_add(&nx, &nx, &h3); fe_sub(&nx, &nx, &v); fe_sub(&nx, &nx, &v);
  fe_sub(&t, &v, &nx); fe_mul(&ny, &rr, &t); fe_mul(&t, &s1, &h3); fe_sub(&ny, &ny, &t);
  fe_mul(&nz, &a->z, &b->z); fe_mul(&nz, &nz, &h);
  r->x = nx; r->y = ny; r->z = nz; r->zone = 0;
}

static int ap_eq(const apoint* a, const apoint* b) {
  if (a->inf != b->inf) return 0;
  return a->inf || (u_cmp(&a->x, &b->x) == 0 && u_cmp(&a->y, &b->y) == 0);
}
/* Point.dbl, short.js:394-412 */
static void ap_dbl(apoint* r, const apoint* p) {
  if (p->inf) { *r = *p; return; }
  fe ys1, x2, dyinv, c, nx, ny, t;
  fe_add(&ys1, &p->y, &p->y);
  if (u_is_zero(&ys1)) { memset(r, 0, sizeof *r); r->inf = 1; return; }
  fe_sqr(&x2, &p->x); fe_inv(&dyinv, &ys1);
  fe_add(&c, &x2, &x2); fe_add(&c, &c, &x2); /* + a, a = 0 */
  fe_mul(&c, &c, &dyinv);
  fe_sqr(&nx, &c); fe_add(&t, &p->x, &p->x); fe_sub(&nx, &nx, &t);
  fe_sub(&t, &p->x, &nx); fe_mul(&ny, &c, &t); fe_sub(&ny, &ny, &p->y);
  r->x = nx; r->y = ny; r->inf = 0;
}
/* Point.add, short.js:365-392 */
static void ap_add(apoint* r, const apoint* a, const apoint* b) {
  if (a->inf) { *r = *b; return; }
  if (b->inf) { *r = *a; return; }
  if (ap_eq(a, b)) { ap_dbl(r, a); return; }
  apoint na; ap_neg(&na, a);
  if (ap_eq(&na, b) || u_cmp(&a->x, &b->x) == 0) { memset(r, 0, sizeof *r); r->inf = 1; return; }
  fe c, t, nx, ny;
  fe_sub(&c, &a->y, &b->y);
  if (!u_is_zero(&c)) { fe_sub(&t, &a->x, &b->x); fe_inv(&t, &t); fe_mul(&c, &c, &t); }
  fe_sqr(&nx, &c); fe_sub(&nx, &nx, &a->x); fe_sub(&nx, &nx, &b->x);
  fe_sub(&t, &a->x, &nx); fe_mul(&ny, &c, &t); fe_sub(&ny, &ny, &a->y);
  r->x = nx; r->y = ny; r->inf = 0;
}

/* ---------------------------------------------------------------- precomputed G tables */
#define NAF_W 7
#define NAF_PTS 128
static apoint G_, G_NAF[NAF_PTS], BG_NAF[NAF_PTS];  /* (2i+1)G and their beta images (short.js:291-307) */
static int g_inited;

static void hex_fe(fe* r, const char* h) {
  uint8_t b[32];
  for (int i = 0; i < 32; i++) { unsigned v; sscanf(h + 2 * i, "%2x", &v); b[i] = (uint8_t)v; }
  u_from_be(r, b);
}
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void init_tables(void) {
  hex_fe(&G_.x, "79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798");
  hex_fe(&G_.y, "483ada7726a3c4655da4fbfc0e1108a8fd17b448a68554199c47d08ffb10d4b8");
  G_.inf = 0;
  hex_fe(&BETA, "7ae96a2b657c07106e64479eac3434e99cf0497512f58995c1396c28719501ee");
  apoint d; ap_dbl(&d, &G_);
  G_NAF[0] = G_;
  for (int i = 1; i < NAF_PTS; i++) ap_add(&G_NAF[i], &G_NAF[i - 1], &d);   /* base.js:357-370 */
  for (int i = 0; i < NAF_PTS; i++) { BG_NAF[i] = G_NAF[i]; fe_mul(&BG_NAF[i].x, &G_NAF[i].x, &BETA); }
  g_inited = 1;
}

/* ---------------------------------------------------------------- scalar recoding */
/* signed integers: sign + 256-bit magnitude */
typedef struct { u256 m; int neg; } sint;

/* utils.getNAF, utils.js:15-44 (k >= 0). Returns length. */
static int get_naf(int8_t* naf, const u256* k, int w, int bits) {
  int len = u_bitlen(k); if (bits > len) len = bits; len += 1;
  int ws = 1 << (w + 1);
  u256 t = *k;
  for (int i = 0; i < len; i++) {
    int z = 0;
    int mod = (int)(t.v[0] & (u64)(ws - 1));
    if (t.v[0] & 1) {
      z = (mod > (ws >> 1) - 1) ? (ws >> 1) - mod : mod;
      u256 zz = {{(u64)(z < 0 ? -z : z), 0, 0, 0}};
      if (z >= 0) u_sub(&t, &t, &zz); else u_add(&t, &t, &zz);
    }
    naf[i] = (int8_t)z;
    u_shr1(&t, 0);
  }
  return len;
}

/* utils.getJSF, utils.js:47-101.  k1,k2 >= 0. */
static int get_jsf(int8_t* j1, int8_t* j2, const u256* a, const u256* b) {
  u256 k1 = *a, k2 = *b;
  int d1 = 0, d2 = 0, len = 0;
  /* k.cmpn(-d) > 0  <=>  k > -d ; with k >= 0: d=0 -> k != 0 ; d=1 -> always (k > -1) */
  while ((d1 ? 1 : !u_is_zero(&k1)) || (d2 ? 1 : !u_is_zero(&k2))) {
    int m14 = (int)((k1.v[0] & 3) + d1) & 3;
    int m24 = (int)((k2.v[0] & 3) + d2) & 3;
    if (m14 == 3) m14 = -1;
    if (m24 == 3) m24 = -1;
    int u1, u2;
    if ((m14 & 1) == 0) u1 = 0;
    else {
      int m8 = (int)((k1.v[0] & 7) + d1) & 7;
      u1 = ((m8 == 3 || m8 == 5) && m24 == 2) ? -m14 : m14;
    }
    j1[len] = (int8_t)u1;
    if ((m24 & 1) == 0) u2 = 0;
    else {
      int m8 = (int)((k2.v[0] & 7) + d2) & 7;
      u2 = ((m8 == 3 || m8 == 5) && m14 == 2) ? -m24 : m24;
    }
    j2[len] = (int8_t)u2;
    len++;
    if (2 * d1 == u1 + 1) d1 = 1 - d1;
    if (2 * d2 == u2 + 1) d2 = 1 - d2;
    u_shr1(&k1, 0); u_shr1(&k2, 0);
  }
  return len;
}

/* exact round(a*k / n) for a < 2^128, k < n: BN.divRound (dist:6387-6404) on non-negative values */
static void mul_div_round_n(u256* q, const u64 a[2], const u256* k) {
  /* t = a*k  (6 limbs) */
  u64 t[7] = {0};
  for (int i = 0; i < 2; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)k->v[j] * a[i] + t[i + j]; t[i + j] = (u64)c; c >>= 64; }
    t[i + 4] = (u64)c;
  }
  /* long division of the 384-bit t by n, bit-serial (restoring) -- exact, simple */
  u64 rem[5] = {0}; u64 quo[6] = {0};
  for (int bit = 383; bit >= 0; bit--) {
    /* rem = rem*2 + bit */
    u64 carry = (t[bit >> 6] >> (bit & 63)) & 1;
    for (int i = 0; i < 5; i++) { u64 nc = rem[i] >> 63; rem[i] = (rem[i] << 1) | carry; carry = nc; }
    /* if rem >= n: rem -= n */
    int ge = rem[4] != 0;
    if (!ge) { u256 r4 = {{rem[0], rem[1], rem[2], rem[3]}}; ge = u_cmp(&r4, &N_) >= 0; }
    if (ge) {
      u256 r4 = {{rem[0], rem[1], rem[2], rem[3]}};
      u64 bw = u_sub(&r4, &r4, &N_);
      rem[0] = r4.v[0]; rem[1] = r4.v[1]; rem[2] = r4.v[2]; rem[3] = r4.v[3]; rem[4] -= bw;
      quo[bit >> 6] |= (u64)1 << (bit & 63);
    }
  }
  u256 qq = {{quo[0], quo[1], quo[2], quo[3]}};
  u256 r4 = {{rem[0], rem[1], rem[2], rem[3]}};
  if (!u_is_zero(&r4)) {
    /* half = n >> 1 ; n odd: round down iff mod < half or mod == half (dist:6396-6401) */
    u256 half = N_; u_shr1(&half, 0);
    int c = u_cmp(&r4, &half);
    if (!(c < 0 || c == 0)) { u256 one = {{1, 0, 0, 0}}; u_add(&qq, &qq, &one); }
  }
  *q = qq;
}

static void s_from(sint* r, const u256* m, int neg) { r->m = *m; r->neg = neg && !u_is_zero(m); }
static void s_addsub(sint* r, const sint* a, const sint* b, int sub) {
  int bneg = b->neg ^ sub;
  if (u_is_zero(&b->m)) { *r = *a; return; }
  if (a->neg == bneg) { u_add(&r->m, &a->m, &b->m); r->neg = a->neg; }
  else {
    int c = u_cmp(&a->m, &b->m);
    if (c >= 0) { u_sub(&r->m, &a->m, &b->m); r->neg = a->neg; }
    else { u_sub(&r->m, &b->m, &a->m); r->neg = bneg; }
  }
  if (u_is_zero(&r->m)) r->neg = 0;
}
static void mul_128x128(u256* r, const u256* a, const u64 b[3]) { /* a < 2^129, b < 2^129: low 256 bits suffice */
  u64 t[8] = {0};
  for (int i = 0; i < 3; i++) {
    u128 c = 0;
    for (int j = 0; j < 4 && i + j < 8; j++) { c += (u128)a->v[j] * b[i] + t[i + j]; t[i + j] = (u64)c; c >>= 64; }
    if (i + 4 < 8) t[i + 4] += (u64)c;
  }
  r->v[0] = t[0]; r->v[1] = t[1]; r->v[2] = t[2]; r->v[3] = t[3];
}

/* ShortCurve._endoSplit, short.js:168-185; basis curves.js:189-198 */
static void endo_split(sint* k1, sint* k2, const u256* k) {
  static const u64 A1[3] = {0xE86C90E49284EB15ULL, 0x3086D221A7D46BCDULL, 0};
  static const u64 MB1[3] = {0x6F547FA90ABFE4C3ULL, 0xE4437ED6010E8828ULL, 0};   /* -b1 */
  static const u64 A2[3] = {0x57C1108D9D44CFD8ULL, 0x14CA50F7A8E2F3F6ULL, 1};
  static const u64 B2[3] = {0xE86C90E49284EB15ULL, 0x3086D221A7D46BCDULL, 0};
  u256 c1, c2, p1, p2, q1, q2;
  mul_div_round_n(&c1, B2, k);     /* c1 = round(b2*k/n) */
  mul_div_round_n(&c2, MB1, k);    /* c2 = round(-b1*k/n) */
  mul_128x128(&p1, &c1, A1);  mul_128x128(&p2, &c2, A2);
  mul_128x128(&q1, &c1, MB1); mul_128x128(&q2, &c2, B2);   /* q1 = -(c1*b1) */
  sint K, P1, P2, t;
  s_from(&K, k, 0); s_from(&P1, &p1, 0); s_from(&P2, &p2, 0);
  s_addsub(&t, &K, &P1, 1); s_addsub(k1, &t, &P2, 1);        /* k1 = k - p1 - p2 */
  /* k2 = -(c1*b1 + c2*b2) = q1' - q2 with q1' = c1*(-b1) */
  sint Q1, Q2; s_from(&Q1, &q1, 0); s_from(&Q2, &q2, 0);
  s_addsub(k2, &Q1, &Q2, 1);
}

/* ---------------------------------------------------------------- _wnafMulAdd for [G, bG, Q, bQ] */
typedef struct { int is_j; apoint a; jpoint j; } anyp;   /* comb entries may be affine or Jacobian */

static void acc_add_any(jpoint* acc, const anyp* p, int neg) {
  jpoint r;
  if (!p->is_j) {
    apoint q = p->a; if (neg) ap_neg(&q, &p->a);
    jp_madd(&r, acc, &q);
  } else {
    jpoint q = p->j; if (neg) jp_neg(&q, &p->j);
    jp_add(&r, acc, &q);
  }
  *acc = r;
}

/* returns 1/0 for verify true/false */
static int verify_one(const uint8_t* e32, const uint8_t* r32, const uint8_t* s32, const uint8_t* x32, const uint8_t* y32) {
  pthread_once(&g_once, init_tables);
  u256 e, r, s;
  u_from_be(&e, e32); u_from_be(&r, r32); u_from_be(&s, s32);
  /* _truncateToN for a 32-byte hash: one conditional subtraction (ec/index.js:104-105) */
  if (u_cmp(&e, &N_) >= 0) u_sub(&e, &e, &N_);
  if (u_is_zero(&r) || u_cmp(&r, &N_) >= 0) return 0;
  if (u_is_zero(&s) || u_cmp(&s, &N_) >= 0) return 0;
  u256 sinv, u1, u2;
  mod_inv(&sinv, &s, &N_);
  sc_mul(&u1, &sinv, &e); sc_mul(&u2, &sinv, &r);

  apoint Q; u256 t;
  u_from_be(&t, x32); fe_set(&Q.x, &t); u_from_be(&t, y32); fe_set(&Q.y, &t); Q.inf = 0;  /* no validation */

  /* _endoWnafMulAdd, short.js:218-249 */
  sint k1g, k2g, k1q, k2q;
  endo_split(&k1g, &k2g, &u1);
  endo_split(&k1q, &k2q, &u2);
  int neg_g = k1g.neg, neg_bg = k2g.neg;   /* whole-table negation == negate at use (short.js:458-480) */
  apoint Qp = Q, Qb = Q;
  fe_mul(&Qb.x, &Q.x, &BETA);              /* _getBeta, short.js:290 */
  if (k1q.neg) ap_neg(&Qp, &Qp);
  if (k2q.neg) ap_neg(&Qb, &Qb);

  /* base.js:150-203: pair (2,3) = (Q, bQ) has wnd 1 -> JSF + comb; pair (0,1) wnd 7 -> getNAF */
  anyp comb[4];
  comb[0].is_j = 0; comb[0].a = Qp;
  comb[3].is_j = 0; comb[3].a = Qb;
  fe nyb; fe_neg(&nyb, &Qb.y);
  if (u_cmp(&Qp.y, &Qb.y) == 0) {
    comb[1].is_j = 0; ap_add(&comb[1].a, &Qp, &Qb);
    apoint nb; ap_neg(&nb, &Qb); jpoint j; ap_to_j(&j, &Qp);
    comb[2].is_j = 1; jp_madd(&comb[2].j, &j, &nb);
  } else if (u_cmp(&Qp.y, &nyb) == 0) {
    jpoint j; ap_to_j(&j, &Qp);
    comb[1].is_j = 1; jp_madd(&comb[1].j, &j, &Qb);
    apoint nb; ap_neg(&nb, &Qb);
    comb[2].is_j = 0; ap_add(&comb[2].a, &Qp, &nb);
  } else {
    jpoint j; ap_to_j(&j, &Qp);
    comb[1].is_j = 1; jp_madd(&comb[1].j, &j, &Qb);
    apoint nb; ap_neg(&nb, &Qb);
    comb[2].is_j = 1; jp_madd(&comb[2].j, &j, &nb);
  }
  static const int8_t INDEX[9] = {-3, -1, -5, -7, 0, 7, 5, 1, 3};
  int8_t j1[264], j2[264], nq[264], ng[264], nbg[264];
  int max = get_jsf(j1, j2, &k1q.m, &k2q.m);
  for (int j = 0; j < max; j++) nq[j] = INDEX[(j1[j] + 1) * 3 + (j2[j] + 1)];
  int lg = get_naf(ng, &k1g.m, NAF_W, 256);
  int lbg = get_naf(nbg, &k2g.m, NAF_W, 256);
  int lq = max;
  if (lg > max) max = lg;
  if (lbg > max) max = lbg;

  /* main loop, base.js:205-244 */
  jpoint acc; jp_inf(&acc);
  for (int i = max; i >= 0; i--) {
    int k = 0, zg = 0, zbg = 0, zq = 0;
    while (i >= 0) {
      zg = i < lg ? ng[i] : 0; zbg = i < lbg ? nbg[i] : 0; zq = i < lq ? nq[i] : 0;
      if (zg || zbg || zq) break;
      k++; i--;
    }
    if (i >= 0) k++;
    { jpoint t2; jp_dblp(&t2, &acc, k); acc = t2; }
    if (i < 0) break;
    if (zg) {
      int az = zg < 0 ? -zg : zg; apoint p = G_NAF[(az - 1) >> 1];
      if ((zg < 0) ^ neg_g) ap_neg(&p, &p);
      jpoint t2; jp_madd(&t2, &acc, &p); acc = t2;
    }
    if (zbg) {
      int az = zbg < 0 ? -zbg : zbg; apoint p = BG_NAF[(az - 1) >> 1];
      if ((zbg < 0) ^ neg_bg) ap_neg(&p, &p);
      jpoint t2; jp_madd(&t2, &acc, &p); acc = t2;
    }
    if (zq) {
      int az = zq < 0 ? -zq : zq;
      acc_add_any(&acc, &comb[(az - 1) >> 1], zq < 0);
    }
  }
  if (jp_is_inf(&acc)) return 0;
  /* eqXToP, short.js:908-925 */
  fe zs, rx, rf, tn, nf;
  fe_sqr(&zs, &acc.z);
  fe_set(&rf, &r); fe_mul(&rx, &rf, &zs);
  if (u_cmp(&acc.x, &rx) == 0) return 1;
  u256 xc = r;
  fe_set(&nf, &N_); fe_mul(&tn, &nf, &zs);
  for (;;) {
    if (u_add(&xc, &xc, &N_)) return 0;
    if (u_cmp(&xc, &P_) >= 0) return 0;
    fe_add(&rx, &rx, &tn);
    if (u_cmp(&acc.x, &rx) == 0) return 1;
  }
}

/* ---------------------------------------------------------------- batch driver */
/* Workers pull 64-item chunks from a shared counter, so one slow or shared core (the host is a
   multi-tenant box) delays the batch by one chunk instead of by its whole static share. */
typedef struct { size_t n; size_t* next; const uint8_t *e, *r, *s, *pub; uint8_t* st; } job;
#define K256_CHUNK 64
static void* worker(void* arg) {
  job* j = (job*)arg;
  for (;;) {
    size_t lo = __atomic_fetch_add(j->next, (size_t)K256_CHUNK, __ATOMIC_RELAXED);
    if (lo >= j->n) break;
    size_t hi = lo + K256_CHUNK < j->n ? lo + K256_CHUNK : j->n;
    for (size_t i = lo; i < hi; i++)
      j->st[i] = (uint8_t)verify_one(j->e + 32 * i, j->r + 32 * i, j->s + 32 * i, j->pub + 64 * i, j->pub + 64 * i + 32);
  }
  return 0;
}

/* e, r, s: n x 32 bytes BE; pub: n x 64 (x||y); status: n bytes (0/1); threads >= 1 */
int k256_ref_verify_batch(size_t n, const uint8_t* e, const uint8_t* r, const uint8_t* s, const uint8_t* pub,
                          uint8_t* status, int threads) {
  pthread_once(&g_once, init_tables);
  if (threads < 1) threads = 1;
  if ((size_t)threads > (n + K256_CHUNK - 1) / K256_CHUNK) threads = n ? (int)((n + K256_CHUNK - 1) / K256_CHUNK) : 1;
  size_t next = 0;
  job jb = {n, &next, e, r, s, pub, status};
  if (threads == 1) { worker(&jb); return 0; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  int started = 0;
  for (int t = 0; t < threads; t++) if (pthread_create(&th[t], 0, worker, &jb) == 0) th[started++] = th[t];
  if (!started) worker(&jb);
  for (int t = 0; t < started; t++) pthread_join(th[t], 0);
  free(th);
  return 0;
}

unsigned long k256_ref_fm_count(void) { return g_count_mul; }
void k256_ref_fm_reset(void) { g_count_mul = 0; }
