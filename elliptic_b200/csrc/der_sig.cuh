// der_sig.cuh -- Signature._importDER (lib/elliptic/ec/signature.js:24-134) restated for one item.
//
// Same accept / reject decisions as the reference, including its JavaScript edge behaviour: reads past the
// end of the array yield `undefined` (which fails `!== 0x30` / `!== 0x02`, passes `& 128`, and makes
// getLength return undefined, which no later length test survives), lengths are 32-bit unsigned, and
// `slice` clamps to the array.  A rejected encoding is the reference's
// Error('Signature without r or s') (ec/signature.js:15).
#pragma once
#include "limbs.cuh"

namespace eb {

struct der_cursor { const uint8_t* p; size_t n; size_t place; };

EB_HD int der_at(const der_cursor& c, size_t i) { return i < c.n ? (int)c.p[i] : -1; }   // -1 = undefined

// getLength, signature.js:24-57.  Returns false for `false` and for `undefined`.
EB_HD bool der_get_length(der_cursor& c, size_t* out) {
  if (c.place >= c.n) { c.place++; return false; }        // undefined -> every caller's test fails
  int initial = c.p[c.place++];
  if (!(initial & 0x80)) { *out = (size_t)initial; return true; }
  int octets = initial & 0xf;
  if (octets == 0 || octets > 4) return false;
  if (der_at(c, c.place) == 0) return false;
  u32 val = 0;
  size_t off = c.place;
  for (int i = 0; i < octets; i++, off++) {
    int b = der_at(c, off);
    val = (val << 8) | (u32)(b < 0 ? 0 : b);
  }
  if (val <= 0x7f) return false;
  c.place = off;
  *out = (size_t)val;
  return true;
}

// Writes r and s as LEN-byte big-endian integers.  Returns false for a rejected encoding.  A value that
// does not fit LEN bytes is >= n for every supported curve, i.e. verify() answers false: it is stored as
// zero, which the range check (ec/index.js:199-202) rejects the same way.
EB_HD bool der_import(const uint8_t* data, size_t n, size_t LEN, uint8_t* r_out, uint8_t* s_out) {
  der_cursor c; c.p = data; c.n = n; c.place = 0;
  if (der_at(c, c.place++) != 0x30) return false;
  size_t len, rlen, slen;
  if (!der_get_length(c, &len)) return false;
  if (len + c.place != n) return false;
  if (der_at(c, c.place++) != 0x02) return false;
  if (!der_get_length(c, &rlen)) return false;
  int b = der_at(c, c.place);
  if (b >= 0 && (b & 128)) return false;
  size_t r0 = c.place < n ? c.place : n;
  size_t r1 = (rlen > n - r0) ? n : r0 + rlen;             // slice clamps
  c.place += rlen;
  if (der_at(c, c.place++) != 0x02) return false;
  if (!der_get_length(c, &slen)) return false;
  if (n != slen + c.place) return false;
  b = der_at(c, c.place);
  if (b >= 0 && (b & 128)) return false;
  size_t s0 = c.place < n ? c.place : n;
  size_t s1 = (slen > n - s0) ? n : s0 + slen;
  // leading zero: allowed only in front of a byte with the top bit set (signature.js:112-127)
  if (r1 > r0 && data[r0] == 0) { if (r1 - r0 > 1 && (data[r0 + 1] & 0x80)) r0++; else return false; }
  if (s1 > s0 && data[s0] == 0) { if (s1 - s0 > 1 && (data[s0 + 1] & 0x80)) s0++; else return false; }
  for (int h = 0; h < 2; h++) {
    size_t a = h ? s0 : r0, e = h ? s1 : r1;
    uint8_t* out = h ? s_out : r_out;
    while (a < e && data[a] == 0) a++;                     // new BN(bytes) ignores leading zeros
    bool fits = (e - a) <= LEN;
    for (size_t k = 0; k < LEN; k++) {
      size_t back = LEN - 1 - k;                           // distance from the last byte
      out[k] = (fits && back < e - a) ? data[e - 1 - back] : 0;
    }
  }
  return true;
}

}  // namespace eb
