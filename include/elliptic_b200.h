/* elliptic_b200.h -- C ABI of libelliptic_b200.so, the drop-in boundary.
 *
 * The reference (indutny/elliptic, pure JavaScript) has no FFI of its own
 * (SURVEY.md 8b); these are the entry points an N-API addon for
 * `require('elliptic')` binds so that whole batches of the reference's
 * single-item calls run on B200 GPUs.  Each function names the reference
 * method whose per-item semantics it reproduces bit-exactly.
 *
 * Conventions: field elements / scalars are fixed-width big-endian byte
 * strings (32 bytes for secp256k1, the reference's toArray('be', len),
 * lib/elliptic/curve/base.js:298-306); arrays are item-major and contiguous
 * (item i of `r` is r[32*i .. 32*i+31]).  The caller owns every buffer.
 * Every function returns EB200_OK (0) or a negative error code; nothing
 * throws or aborts.  Per-item outcomes are written to `status`.
 * There is NO CPU fallback: without a usable CUDA device every compute entry
 * point returns EB200_ERR_NO_DEVICE.
 */
#ifndef ELLIPTIC_B200_H
#define ELLIPTIC_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define EB200_OK 0
#define EB200_ERR_NO_DEVICE (-1)   /* no CUDA device / driver */
#define EB200_ERR_CUDA (-2)        /* a CUDA call failed: eb200_last_error() */
#define EB200_ERR_ARG (-3)         /* bad argument (NULL buffer, unknown curve/format) */
#define EB200_ERR_NOT_INIT (-4)    /* eb200_init() has not succeeded */
#define EB200_ERR_UNSUPPORTED (-5) /* curve / format not built yet */

/* per-item status byte: what the reference's call would have done */
#define EB200_ST_FALSE 0                /* returned false */
#define EB200_ST_TRUE 1                 /* returned true */
#define EB200_ST_THROW_INVALID_POINT 2  /* threw Error('invalid point')  short.js:195, edwards.js:84 */
#define EB200_ST_THROW_NOT_VALIDATED 3  /* threw Error('public point not validated')  ec/key.js:104 */
#define EB200_ST_NEEDS_HOST 4           /* internal: the fast kernel's flag for an off-curve un-validated key
                                           (SURVEY 8a Q1).  Every verify entry point re-runs flagged items
                                           through an exact replay of the reference's own schedule on the GPU,
                                           so callers never see this value */
#define EB200_ST_THROW_ASSERT 5         /* threw Error('Assertion failed') (bn.js sqrt / hybrid parity) */
#define EB200_ST_THROW_POINT_FORMAT 6   /* threw Error('Unknown point format')  base.js:291 */
#define EB200_ST_INFINITY 7             /* (recover) returned the point at infinity */
#define EB200_ST_THROW_SECOND_KEY 8     /* (recover) threw Error('Unable to find sencond key candinate')  ec/index.js:244 */
#define EB200_ST_THROW_SIG_FORMAT 9     /* threw Error('Signature without r or s')  ec/signature.js:15 (DER rejected by _importDER) */
#define EB200_ST_RETRY 10               /* (sign with caller nonces) the reference's loop `continue`s: k outside [2, n-2], r = 0 or s = 0;
                                           the caller supplies its next k(iter), ec/index.js:153-185 */

/* curve ids (names of lib/elliptic/curves.js presets) */
#define EB200_CURVE_SECP256K1 1
#define EB200_CURVE_P256 2
#define EB200_CURVE_P384 3
#define EB200_CURVE_ED25519 4
#define EB200_CURVE_CURVE25519 5
#define EB200_CURVE_P521 6       /* 66-byte fields; verify / recover / mul / mulAdd / derive */
#define EB200_CURVE_P192 7       /* 24-byte fields; same entry points as p521 */
#define EB200_CURVE_P224 8       /* 28-byte fields; p = 1 mod 4: compressed keys go through bn.js's Tonelli-Shanks */

/* public-key encodings accepted by eb200_ecdsa_verify_batch (KeyPair._importPublic,
 * lib/elliptic/ec/key.js:84-99 -> BaseCurve.decodePoint, curve/base.js:270-292) */
#define EB200_PUB_XY 0          /* {x, y}: 2*len bytes per item, NOT validated (as the reference) */
#define EB200_PUB_SEC1_65 1     /* 04|06|07 || x || y : 1+2*len bytes per item */
#define EB200_PUB_SEC1_33 2     /* 02|03 || x : 1+len bytes per item (pointFromX, short.js:187-204) */

typedef struct eb200_timing {
  float h2d_ms;     /* host->device copies of the last host-buffer call */
  float kernel_ms;  /* all kernels of the last call (CUDA events on the launch stream) */
  float d2h_ms;     /* device->host copy of the results */
  float main_kernel_ms; /* the dominant kernel only */
  uint32_t launches; /* kernels launched by the last call */
} eb200_timing;

/* Create (or keep) one context per listed CUDA ordinal and build the fixed-base tables there
 * (eb200_init(devices[], ndev, flags)).  devices == NULL or ndev <= 0: every visible device.
 * Idempotent per device; later calls add devices.  Host-pointer entry points split a batch into contiguous
 * blocks over the initialised devices (no exchange between blocks) when it has at least 2^14 items per
 * device, each block driven by its own host thread; smaller calls take one device, rotating, so that concurrent
 * callers spread out.  Device-pointer (`_dev`) entry points run on the device that owns `d_status`.
 * Every entry point is safe to call from several threads. */
#define EB200_INIT_ALL_TABLES 1u   /* build every curve's fixed-base table now instead of on first use */
int eb200_init(const int* devices, int ndev, uint32_t flags);
int eb200_shutdown(void);
int eb200_device_count(void);          /* devices initialised so far */
const char* eb200_strerror(int code);
const char* eb200_last_error(void);   /* text of the last CUDA error on this thread's context */
int eb200_last_timing(eb200_timing* out);

/* Batch of EC.prototype.verify (lib/elliptic/ec/index.js:188-229) for `curve`.
 *   e   : n x len  message hashes already truncated as _truncateToN does for a len-byte
 *                  input (ec/index.js:81-108); values >= n are accepted like the reference
 *   r,s : n x len  signature halves (Signature{r,s}, ec/signature.js:8-22)
 *   pub : n x (pub_fmt-dependent) public keys
 *   status : n bytes out
 * Host pointers (pinned or pageable: pageable buffers are staged through an internal pinned ring with a
 * parallel memcpy, so an unpinned caller such as a Node.js Buffer gets the pinned transfer rate); copies are done internally on the library's stream. */
int eb200_ecdsa_verify_batch(int curve, size_t n, const uint8_t* e, const uint8_t* r,
                             const uint8_t* s, const uint8_t* pub, uint32_t pub_fmt,
                             uint8_t* status);

/* Same call with the signatures as the reference takes them off the wire: DER (`new Signature(der)`,
 * lib/elliptic/ec/signature.js:73-134), parsed on the GPU.
 *   sigs    : the DER encodings back to back
 *   sig_off : n + 1 byte offsets into `sigs` (item i is sigs[sig_off[i] .. sig_off[i+1]))
 * status adds THROW_SIG_FORMAT for an encoding _importDER rejects; a key that throws takes precedence
 * (keyFromPublic runs before new Signature, ec/index.js:194-195). */
int eb200_ecdsa_verify_batch_der(int curve, size_t n, const uint8_t* e, const uint8_t* sigs, const uint64_t* sig_off,
                                 const uint8_t* pub, uint32_t pub_fmt, uint8_t* status);

/* Same, with DEVICE pointers and a caller-supplied CUDA stream (cudaStream_t cast to void*;
 * NULL = the CUDA default stream).  Asynchronous: the caller synchronises the stream.
 * `workspace` must hold eb200_ecdsa_verify_workspace_bytes(curve, n) bytes of device memory. */
size_t eb200_ecdsa_verify_workspace_bytes(int curve, size_t n);
int eb200_ecdsa_verify_batch_dev(int curve, size_t n, const uint8_t* d_e, const uint8_t* d_r,
                                 const uint8_t* d_s, const uint8_t* d_pub, uint32_t pub_fmt,
                                 uint8_t* d_status, void* d_workspace, void* stream);

/* Batch of EC.prototype.sign (lib/elliptic/ec/index.js:110-186) on every short preset (len = the curve's field
 * byte length), with the curve's default hash (sha256; sha384 on p384, sha512 on p521; curves.js:43-206) and no
 * `pers` / custom `k`:
 * the RFC 6979 nonces come from HMAC-DRBG over that hash, generated on the GPU.
 *   e    : n x len  _truncateToN(msg) (ec/index.js:127) including its final `- n`, i.e. e < n, big-endian
 *   priv : n x len  private scalars as the key pair holds them (reduced mod n at import, ec/key.js:76-82)
 *   flags: EB200_SIGN_CANONICAL = the `canonical` option (s <= n/2, recovery bit flipped)
 *   out_r, out_s : n x len big-endian; out_recid : n bytes (recoveryParam)
 * status: TRUE for every item (the reference's retry loop runs inside the kernel). */
#define EB200_SIGN_CANONICAL 1u
int eb200_ecdsa_sign_batch(int curve, size_t n, const uint8_t* e, const uint8_t* priv, uint32_t flags,
                           uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid, uint8_t* status);

/* Batch of EC.prototype.recoverPubKey (lib/elliptic/ec/index.js:231-259) on secp256k1 / p256 / p384 / p521
 * (len = 32 / 32 / 48 / 66 in place of the 32 and 64 below):
 *   e     : n x 32  `new BN(msg)` reduced mod n (NOT truncated -- the reference does not truncate here)
 *   r, s  : n x 32  signature halves (no range check in the reference: r = 0 yields the point at infinity)
 *   recid : n bytes, the recovery parameter j in 0..3 (bit 0 = y parity, bit 1 = use r + n)
 *   out_xy: n x 64  recovered public key x || y big-endian (zeroed unless status is TRUE)
 * status: TRUE (point written), INFINITY, THROW_INVALID_POINT (pointFromX, short.js:195), THROW_SECOND_KEY. */
int eb200_ecdsa_recover_batch(int curve, size_t n, const uint8_t* e, const uint8_t* r, const uint8_t* s,
                              const uint8_t* recid, uint8_t* out_xy, uint8_t* status);

/* Batch of BasePoint.mul (lib/elliptic/curve/short.js:422-432) on secp256k1 / p256 / p384 (len = 32/32/48):
 *   k         : n x len big-endian scalars, any value below 2^(8 len) (the reference does not reduce them)
 *   points_xy : n x 2len x || y big-endian, or NULL for the base point (G.mul(k) -> _fixedNafMul, base.js:52-84)
 *   out_xy    : n x 2len affine result as Point.toP / getX / getY give it (zeroed unless status is TRUE)
 * status: TRUE (point written) or INFINITY.  Points are not validated, exactly as `curve.point(x, y)`
 * (short.js:251-271); an off-curve point gets the result of the reference's own add/double sequence. */
int eb200_scalar_mul_batch(int curve, size_t n, const uint8_t* k, const uint8_t* points_xy, uint8_t* out_xy,
                           uint8_t* status);

/* Batch of KeyPair.prototype.derive (lib/elliptic/ec/key.js:102-107) on the short curves: ECDH shared x.
 *   priv   : n x len private scalars, big-endian (reduced mod n as _importPrivate does, ec/key.js:76-82)
 *   pub_xy : n x 2len peer points x || y
 *   out_x  : n x len  pub.mul(priv).getX(), big-endian (zeroed unless status is TRUE)
 * status: TRUE, THROW_NOT_VALIDATED (the peer point is not on the curve), INFINITY (priv = 0 mod n: the
 * reference then dies with a TypeError inside getX()).  curve25519 has its own entry point below. */
int eb200_ecdh_derive_batch(int curve, size_t n, const uint8_t* priv, const uint8_t* pub_xy, uint8_t* out_x,
                            uint8_t* status);

/* Batch of G.mulAdd(k1, P2, k2) = k1*G + k2*P2 (lib/elliptic/curve/short.js:434-441; _endoWnafMulAdd
 * short.js:218-249 on secp256k1, _wnafMulAdd base.js:128-253 on p256 / p384).  Arguments and status as
 * eb200_scalar_mul_batch. */
int eb200_mul_add_batch(int curve, size_t n, const uint8_t* k1, const uint8_t* k2, const uint8_t* p2_xy,
                        uint8_t* out_xy, uint8_t* status);

/* Batch of EDDSA.prototype.verify (lib/elliptic/eddsa/index.js:52-63) on ed25519.
 *   R, S : the two 32-byte halves of each signature as on the wire (eddsa/signature.js:17-40)
 *   A    : 32-byte encoded public keys (eddsa/key.js:17-30)
 *   h    : SHA512(R || A || M) as a little-endian integer reduced mod n, 32 bytes LE (hashInt,
 *          eddsa/index.js:65-70) -- computed by the caller (the host wrapper hashes with hashlib /
 *          the N-API shim with hash.js); h MUST be < n.
 * status: TRUE / FALSE, or THROW_INVALID_POINT / THROW_ASSERT where the reference throws while
 * decoding R or A (edwards.js:84, bn.js Red.sqrt assertion). */
int eb200_eddsa_verify_batch(size_t n, const uint8_t* R, const uint8_t* S, const uint8_t* A,
                             const uint8_t* h, uint8_t* status);
/* Same, hashing on the GPU: msgs = all messages concatenated, message i = msgs[msg_off[i] .. msg_off[i+1])
 * (msg_off has n+1 entries).  h = SHA512(R || A || M) mod n is computed in a first kernel. */
int eb200_eddsa_verify_batch_msgs(size_t n, const uint8_t* R, const uint8_t* S, const uint8_t* A,
                                  const uint8_t* msgs, const uint64_t* msg_off, uint8_t* status);
size_t eb200_eddsa_verify_workspace_bytes(size_t n);
int eb200_eddsa_verify_batch_dev(size_t n, const uint8_t* d_R, const uint8_t* d_S, const uint8_t* d_A,
                                 const uint8_t* d_h, uint8_t* d_status, void* d_workspace, void* stream);

/* Batch of KeyPair.prototype.derive (lib/elliptic/ec/key.js:102-107) on curve25519:
 *   priv : n x 32 bytes big-endian, the key pair's private scalar as the reference holds it
 *          (reduced mod n at import, ec/key.js:76-82; no clamping)
 *   pubx : n x 32 bytes big-endian x coordinate of the peer point (mont.js:46-48 decodePoint)
 *   out  : n x 32 bytes big-endian shared x (BN -> toArray('be', 32)); zeroed when the call throws
 * status: TRUE = value returned; THROW_ASSERT = the reference throws inside validate()
 *         (twist point: Red.sqrt assertion, mont.js:21-28). */
int eb200_x25519_derive_batch(size_t n, const uint8_t* priv, const uint8_t* pubx, uint8_t* out, uint8_t* status);
int eb200_x25519_derive_batch_dev(size_t n, const uint8_t* d_priv, const uint8_t* d_pubx, uint8_t* d_out,
                                  uint8_t* d_status, void* stream);

/* EC.prototype.sign with the `k` option (options.k(iter), lib/elliptic/ec/index.js:154-157): one attempt of the
 * reference's loop with the caller's nonces.  k: n x len bytes big-endian, what k(iter) returned (it goes through
 * _truncateToN(k, true) here).  status: EB200_ST_TRUE, or EB200_ST_RETRY where the reference would call k(iter + 1). */
int eb200_ecdsa_sign_batch_k(int curve, size_t n, const uint8_t* e, const uint8_t* priv, const uint8_t* k, uint32_t flags,
                             uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid, uint8_t* status);
/* EC.prototype.sign with the `pers` option (ec/index.js:143-151): HMAC-DRBG seeded with key || msg || pers; pers is the
 * personalisation string after `persEnc` decoding, shared by the whole batch. */
int eb200_ecdsa_sign_batch_pers(int curve, size_t n, const uint8_t* e, const uint8_t* priv, const uint8_t* pers, size_t pers_len,
                                uint32_t flags, uint8_t* out_r, uint8_t* out_s, uint8_t* out_recid, uint8_t* status);
/* EC.prototype.genKeyPair({entropy, pers}) (ec/index.js:55-79): per item HMAC-DRBG(hash, entropy_i, nonce = n.toArray(),
 * pers), the first candidate <= n - 2 plus one as the private key, and its public point.
 *   entropy : n x entropy_len bytes (the reference requires >= hmacStrength / 8 = 24);  out_priv : n x len;
 *   out_pub_xy : n x 2 len or NULL */
int eb200_ec_keygen_batch(int curve, size_t n, const uint8_t* entropy, size_t entropy_len, const uint8_t* pers, size_t pers_len,
                          uint8_t* out_priv, uint8_t* out_pub_xy, uint8_t* status);

/* The `ec` / `.curve` API over the other two curve types of lib/elliptic/curves.js:
 *  - EB200_CURVE_ED25519 is accepted by eb200_ecdsa_verify_batch (+ _der, SEC1 formats: BaseCurve.decodePoint and
 *    EdwardsCurve.pointFromX, edwards.js:46-69), eb200_ecdsa_sign_batch (+ _k, _pers), eb200_ec_keygen_batch,
 *    eb200_scalar_mul_batch / eb200_mul_add_batch (Point.mul / mulAdd, edwards.js:362-375; 32-byte big-endian x || y,
 *    the neutral element is the ordinary point (0, 1)) and eb200_ecdh_derive_batch: new elliptic.ec('ed25519')
 *    (test/ecdsa-test.js:130, test/ecdh-test.js:26; eqXToP edwards.js:415-431).  An un-validated off-curve point is
 *    reported as EB200_ST_NEEDS_HOST (the reference's answer then depends on its own wNAF schedule, which is not
 *    replayed for this curve); eb200_ecdsa_recover_batch returns EB200_ERR_UNSUPPORTED.
 *  - curve25519 points are x-only: eb200_x25519_mul_batch is MontCurve Point.mul(k).getX() (mont.js:130-153) without
 *    the validation that eb200_x25519_derive_batch (KeyPair.derive) performs; mulAdd throws in the reference. */
int eb200_x25519_mul_batch(size_t n, const uint8_t* k, const uint8_t* px, uint8_t* out_x, uint8_t* status);

/* Short Weierstrass curves given at run time -- the batch form of `new elliptic.curve.short({p, a, b})`
 * (lib/elliptic/curve/short.js:10-24) and of Point.mul / mulAdd / add / dbl / validate on its points
 * (short.js:365-450, 206-216).  p: any odd prime > 3 of up to 576 bits; p, a, b: `len` bytes big-endian; points:
 * x || y, `len` bytes each (values >= p are reduced on entry like toRed); scalars: `klen` bytes big-endian, any value.
 * status: EB200_ST_TRUE = affine point written, EB200_ST_INFINITY = the point at infinity (output zeroed),
 * EB200_ST_NEEDS_HOST = an input does not satisfy the curve equation (the reference does not validate and its result
 * is then an artefact of its own schedule: not accelerated); validate: EB200_ST_TRUE / EB200_ST_FALSE.
 * The six presets keep their tuned entry points above; this generic path is not tuned (one thread per item,
 * double-and-add). */
typedef struct eb200_short_curve { uint32_t len; const uint8_t* p; const uint8_t* a; const uint8_t* b; } eb200_short_curve;
int eb200_curve_mul_batch(const eb200_short_curve* curve, size_t n, const uint8_t* k, size_t klen, const uint8_t* points_xy,
                          uint8_t* out_xy, uint8_t* status);
int eb200_curve_mul_add_batch(const eb200_short_curve* curve, size_t n, const uint8_t* k1, const uint8_t* p1_xy, const uint8_t* k2,
                              const uint8_t* p2_xy, size_t klen, uint8_t* out_xy, uint8_t* status);
int eb200_curve_add_batch(const eb200_short_curve* curve, size_t n, const uint8_t* p1_xy, const uint8_t* p2_xy, uint8_t* out_xy,
                          uint8_t* status);
int eb200_curve_dbl_batch(const eb200_short_curve* curve, size_t n, const uint8_t* p_xy, uint8_t* out_xy, uint8_t* status);
int eb200_curve_validate_batch(const eb200_short_curve* curve, size_t n, const uint8_t* p_xy, uint8_t* status);

/* Batch of EDDSA.prototype.sign (lib/elliptic/eddsa/index.js:34-44) with keys given as 32-byte secrets
 * (eddsa.keyFromSecret, eddsa/key.js:52-75: SHA-512 of the secret, clamped scalar, message prefix).
 *   secrets : n x 32 bytes;  msgs / msg_off : concatenated raw messages and n + 1 offsets
 *   out_sig : n x 64 bytes  Rencoded || S (little-endian), byte-identical to sig.toBytes()
 *   out_pub : n x 32 bytes  key.getPublic('bytes'), or NULL
 *   status  : n bytes, always EB200_ST_TRUE (the reference cannot fail on a 32-byte secret)
 * Everything (three SHA-512 per item, two fixed-base multiplications, the arithmetic mod n) runs on the GPU. */
int eb200_eddsa_sign_batch(size_t n, const uint8_t* secrets, const uint8_t* msgs, const uint64_t* msg_off,
                           uint8_t* out_sig, uint8_t* out_pub, uint8_t* status);

/* Self-test hooks used by the parity tests (device arithmetic vs the oracle).
 * op: 0 mul, 1 sqr, 2 add, 3 sub, 4 neg, 5 mul_small(b[0]), 6 normalize, 7 inv, 8 sqrt candidate.
 * a, b, out: n x 8 little-endian 32-bit limbs (host pointers). */
int eb200_selftest_fe(int curve, int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out);
/* Geometry of the fixed-base table: entry (j, i) = (2i+1) * 2^(wbits*j) * G, 16 words (x||y limbs). */
int eb200_selftest_gtab_dims(int curve, int* windows, int* entries, int* wbits);
/* Copy the fixed-base table of `curve` to the host (n_words 32-bit words available). */
int eb200_selftest_gtab(int curve, uint32_t* out, size_t n_words);

#ifdef __cplusplus
}
#endif
#endif
