'use strict';
// Drop-in for require('elliptic'): every export of lib/elliptic.js:5-13 is the reference's own object
// (single-item behaviour unchanged); the batch entry points below are added on the same prototypes and
// run on the GPU through the N-API addon -> libelliptic_b200.so (include/elliptic_b200.h).
// Inputs whose reference result depends on the reference's own add/double schedule (un-validated off-curve
// keys and points) are replayed with that exact schedule inside the library, so a status is always the
// reference's own verdict; the `=== 4` branch below is a defensive leftover and is never taken.
var elliptic = require('elliptic');
var native = require('./elliptic_b200.node');
var CURVE = { secp256k1: 1, p256: 2, p384: 3 };
var THROW = { 2: 'invalid point', 3: 'public point not validated', 5: 'Assertion failed', 6: 'Unknown point format',
  8: 'Unable to find sencond key candinate', 9: 'Signature without r or s' };
var inited = false;
function init() { if (!inited) { native.init(0); inited = true; } }

function be(bn, len) { return bn.toArray('be', len); }

// EC#verifyBatch(msgs, sigs, keys[, enc]) -> Array<boolean>; throws where a loop over verify() would.
elliptic.ec.prototype.verifyBatch = function verifyBatch(msgs, sigs, keys, enc) {
  var name = Object.keys(CURVE).filter(function(n) { return elliptic.curves[n].curve === this.curve; }, this)[0];
  if (!name) return msgs.map(function(m, i) { return this.verify(m, sigs[i], keys[i], enc); }, this);
  init();
  var len = this.curve.p.byteLength(), n = msgs.length;
  var e = new Uint8Array(n * len), r = new Uint8Array(n * len), s = new Uint8Array(n * len);
  var pub = new Uint8Array(n * 2 * len), early = {};
  var Signature = this.sign('00', '01').constructor;                 // lib/elliptic/ec/signature.js
  for (var i = 0; i < n; i++) {
    var msg = this._truncateToN(msgs[i], false);                      // ec/index.js:192
    var key = this.keyFromPublic(keys[i], enc).getPublic();           // ec/index.js:193 (may throw, as verify)
    var sig = new Signature(sigs[i], 'hex');                          // ec/index.js:194
    if (sig.r.cmpn(1) < 0 || sig.r.cmp(this.n) >= 0 || sig.s.cmpn(1) < 0 || sig.s.cmp(this.n) >= 0) { early[i] = false; continue; }
    e.set(be(msg, len), i * len); r.set(be(sig.r, len), i * len); s.set(be(sig.s, len), i * len);
    pub.set(be(key.getX(), len), 2 * i * len); pub.set(be(key.getY(), len), (2 * i + 1) * len);
  }
  var st = native.ecdsaVerifyBatch(CURVE[name], e, r, s, pub, 0);
  var out = new Array(n);
  for (i = 0; i < n; i++) {
    if (i in early) out[i] = false;
    else if (st[i] === 4) out[i] = this.verify(msgs[i], sigs[i], keys[i], enc);   // reference path, exact
    else if (st[i] > 1) throw new Error(THROW[st[i]]);
    else out[i] = st[i] === 1;
  }
  return out;
};

// EDDSA#verifyBatch(messages, sigs, pubs) -> Array<boolean>
elliptic.eddsa.prototype.verifyBatch = function verifyBatch(messages, sigs, pubs) {
  init();
  var n = messages.length, R = new Uint8Array(32 * n), S = new Uint8Array(32 * n), A = new Uint8Array(32 * n), h = new Uint8Array(32 * n);
  for (var i = 0; i < n; i++) {
    var sig = this.makeSignature(sigs[i]);                             // asserts the size (eddsa/signature.js:23-24)
    var key = this.keyFromPublic(pubs[i]);
    R.set(sig.Rencoded(), 32 * i); S.set(sig.Sencoded(), 32 * i); A.set(key.pubBytes(), 32 * i);
    h.set(this.hashInt(sig.Rencoded(), key.pubBytes(), elliptic.utils.parseBytes(messages[i])).toArray('le', 32), 32 * i);
  }
  var st = native.eddsaVerifyBatch(R, S, A, h);
  return Array.prototype.map.call(st, function(v) { if (v > 1) throw new Error(THROW[v]); return v === 1; });
};

// EC#deriveBatch(privs, pubs) on curve25519 -> Array<BN>
elliptic.ec.prototype.deriveBatch = function deriveBatch(privs, pubs) {
  if (this.curve.type !== 'mont') return privs.map(function(p, i) { return this.keyFromPrivate(p).derive(this.keyFromPublic(pubs[i]).getPublic()); }, this);
  init();
  var n = privs.length, k = new Uint8Array(32 * n), x = new Uint8Array(32 * n);
  for (var i = 0; i < n; i++) {
    k.set(be(this.keyFromPrivate(privs[i]).getPrivate(), 32), 32 * i);
    x.set(be(this.keyFromPublic(pubs[i]).getPublic().x.fromRed(), 32), 32 * i);
  }
  var res = native.x25519DeriveBatch(k, x, {});
  var BN = this.n.constructor, out = [];
  for (i = 0; i < n; i++) {
    if (res.status[i] !== 1) throw new Error(THROW[res.status[i]]);
    out.push(new BN(res.out.subarray(32 * i, 32 * i + 32)));
  }
  return out;
};

function curveId(ec) {
  return CURVE[Object.keys(CURVE).filter(function(n) { return elliptic.curves[n].curve === ec.curve; })[0]];
}
function pack(list, len, f) {
  var out = new Uint8Array(list.length * len);
  list.forEach(function(v, i) { out.set(f(v), i * len); });
  return out;
}

// EC#signBatch(msgs, keys[, {canonical}]) -> Array<Signature>  (secp256k1; RFC 6979 nonces made on the GPU)
elliptic.ec.prototype.signBatch = function signBatch(msgs, keys, options) {
  init();
  var self = this, len = 32, n = msgs.length, BN = this.n.constructor;
  var e = pack(msgs, len, function(m) { return be(self._truncateToN(new BN(m, 16)), len); });
  var d = pack(keys, len, function(k) { return be(self.keyFromPrivate(k).getPrivate(), len); });
  var res = native.ecdsaSignBatch(curveId(this), e, d, options && options.canonical ? 1 : 0, {});
  var Signature = this.sign('00', '01').constructor, out = [];
  for (var i = 0; i < n; i++)
    out.push(new Signature({ r: new BN(res.r.subarray(len * i, len * i + len)), s: new BN(res.s.subarray(len * i, len * i + len)),
      recoveryParam: res.recid[i] }));
  return out;
};

// EC#recoverPubKeyBatch(msgs, sigs, js) -> Array<Point>
elliptic.ec.prototype.recoverPubKeyBatch = function recoverPubKeyBatch(msgs, sigs, js) {
  init();
  var self = this, len = 32, n = msgs.length, BN = this.n.constructor;
  var Signature = this.sign('00', '01').constructor;
  var S = sigs.map(function(s) { return new Signature(s, 'hex'); });
  var e = pack(msgs, len, function(m) { return be(new BN(m).umod(self.n), len); });
  var r = pack(S, len, function(s) { return be(s.r.maskn(256), len); }), s = pack(S, len, function(x) { return be(x.s.umod(self.n), len); });
  var res = native.ecdsaRecoverBatch(curveId(this), e, r, s, Uint8Array.from(js), {});
  var out = [];
  for (var i = 0; i < n; i++) {
    if (res.status[i] === 7) out.push(this.curve.point(null, null));
    else if (res.status[i] !== 1) throw new Error(THROW[res.status[i]]);
    else out.push(this.curve.point(new BN(res.pub.subarray(64 * i, 64 * i + 32)), new BN(res.pub.subarray(64 * i + 32, 64 * i + 64))));
  }
  return out;
};

// curve#mulBatch(points | null, ks) and curve#mulAddBatch(k1s, p2s, k2s) on the short curves -> Array<Point>
function pointsOut(ec, res, n, len) {
  var BN = ec.n.constructor, out = [];
  for (var i = 0; i < n; i++)
    out.push(res.status[i] === 1 ? ec.curve.point(new BN(res.points.subarray(2 * len * i, 2 * len * i + len)),
      new BN(res.points.subarray(2 * len * i + len, 2 * len * (i + 1)))) : ec.curve.point(null, null));
  return out;
}
elliptic.ec.prototype.mulBatch = function mulBatch(points, ks) {
  init();
  var len = this.curve.p.byteLength(), BN = this.n.constructor, big = new BN(1).ushln(8 * len), n = this.n;
  var k = pack(ks, len, function(v) { v = new BN(v, 16); return be(v.cmp(big) >= 0 ? v.umod(n) : v, len); });
  var p = points && pack(points, 2 * len, function(pt) { return be(pt.getX(), len).concat(be(pt.getY(), len)); });
  return pointsOut(this, native.mulAddBatch(curveId(this), null, k, p, {}), ks.length, len);
};
elliptic.ec.prototype.mulAddBatch = function mulAddBatch(k1s, p2s, k2s) {
  init();
  var len = this.curve.p.byteLength(), BN = this.n.constructor;
  var f = function(v) { return be(new BN(v, 16), len); };
  var p = pack(p2s, 2 * len, function(pt) { return be(pt.getX(), len).concat(be(pt.getY(), len)); });
  return pointsOut(this, native.mulAddBatch(curveId(this), pack(k1s, len, f), pack(k2s, len, f), p, {}), k1s.length, len);
};

module.exports = elliptic;
