// Host-side BN254 Fq arithmetic (4 x u64 Montgomery limbs): normalisation of the few XYZZ results a batch of MSMs /
// point sums returns (X/ZZ, Y/ZZZ with ONE shared inversion, Montgomery's trick).  A single dependent inversion is
// ~400 serial products: microseconds on a CPU core, ~0.2 ms on one GPU thread, and the affine points are consumed by
// the host-side transcript anyway.  Product code (O(batch) scalar glue), not a fallback for any kernel.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace ja {
namespace host {

struct FqH { uint64_t l[4]; bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; } };

namespace fq {
typedef unsigned __int128 u128;
static const uint64_t P[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t INV = 0x87d20782e4866389ull;   // -q^-1 mod 2^64
static const FqH ONE = {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}};

static inline bool geq_p(const uint64_t* a) {
  for (int i = 3; i >= 0; i--) { if (a[i] > P[i]) return true; if (a[i] < P[i]) return false; }
  return true;
}
static inline void sub_p(uint64_t* a) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - P[i] - (uint64_t)b; a[i] = (uint64_t)t; b = (t >> 64) & 1; }
}
static inline FqH mul(const FqH& a, const FqH& b) {     // CIOS Montgomery product
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    const uint64_t m = t[0] * INV;
    c = (u128)m * P[0] + t[0]; c >>= 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  FqH r = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || geq_p(r.l)) sub_p(r.l);
  return r;
}
static inline FqH inv(const FqH& a) {                   // a^(q-2); inverse of 0 is 0
  const uint64_t e[4] = {P[0] - 2, P[1], P[2], P[3]};
  FqH r = ONE;
  for (int i = 253; i >= 0; i--) {
    r = mul(r, r);
    if ((e[i / 64] >> (i % 64)) & 1) r = mul(r, a);
  }
  return r;
}
}  // namespace fq

namespace fq {
static inline FqH add(const FqH& a, const FqH& b) {
  FqH r; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
  if (geq_p(r.l)) sub_p(r.l);
  return r;
}
static inline FqH sub(const FqH& a, const FqH& b) {
  FqH r; u128 br = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a.l[i] - b.l[i] - (uint64_t)br; r.l[i] = (uint64_t)t; br = (t >> 64) & 1; }
  if (br) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.l[i] + P[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
  return r;
}
static inline FqH dbl(const FqH& a) { return add(a, a); }
static const FqH ZERO = {{0, 0, 0, 0}};
}  // namespace fq

struct G1XH { FqH X, Y, ZZ, ZZZ; };                      // mirrors the device G1X (128 B)

// acc += (x, y) affine, complete (EFD xyzz madd-2008-s / dbl-2008-s-1 with a = 0): combining the per-GPU partial points
// of a sharded MSM is a handful of additions, done where the transcript lives.
static inline void xyzz_madd(G1XH& acc, const FqH& x, const FqH& y) {
  using namespace fq;
  if (acc.ZZ.is_zero()) { acc.X = x; acc.Y = y; acc.ZZ = ONE; acc.ZZZ = ONE; return; }
  const FqH U2 = mul(x, acc.ZZ), S2 = mul(y, acc.ZZZ);
  const FqH Pp = sub(U2, acc.X), R = sub(S2, acc.Y);
  if (Pp.is_zero()) {
    if (!R.is_zero()) { acc.X = ZERO; acc.Y = ZERO; acc.ZZ = ZERO; acc.ZZZ = ZERO; return; }   // P + (-P)
    const FqH U = dbl(y), V = mul(U, U), W = mul(U, V), S = mul(x, V), XX = mul(x, x);          // 2P from affine
    const FqH M = add(dbl(XX), XX);
    acc.X = sub(sub(mul(M, M), S), S);
    acc.Y = sub(mul(M, sub(S, acc.X)), mul(W, y));
    acc.ZZ = V; acc.ZZZ = W;
    return;
  }
  const FqH PP = mul(Pp, Pp), PPP = mul(Pp, PP), Q = mul(acc.X, PP);
  const FqH X3 = sub(sub(sub(mul(R, R), PPP), Q), Q);
  const FqH Y3 = sub(mul(R, sub(Q, X3)), mul(acc.Y, PPP));
  acc.X = X3; acc.Y = Y3;
  acc.ZZ = mul(acc.ZZ, PP);
  acc.ZZZ = mul(acc.ZZZ, PPP);
}

// Affine coordinates of `n` XYZZ points with one inversion.  out_xy = n x 8 limbs, is_inf[i] = 1 for ZZ == 0.
static inline void xyzz_batch_to_affine(const G1XH* pts, size_t n, uint64_t* out_xy, int32_t* is_inf) {
  std::vector<FqH> den(n), pre(n);
  FqH run = fq::ONE;
  for (size_t i = 0; i < n; i++) {
    const bool inf = pts[i].ZZ.is_zero();
    den[i] = inf ? fq::ONE : fq::mul(pts[i].ZZ, pts[i].ZZZ);       // 1/(ZZ*ZZZ): X/ZZ = X*ZZZ*i, Y/ZZZ = Y*ZZ*i
    pre[i] = run;
    run = fq::mul(run, den[i]);
  }
  FqH iv = fq::inv(run);
  for (size_t i = n; i-- > 0;) {
    const FqH di = fq::mul(iv, pre[i]);
    iv = fq::mul(iv, den[i]);
    const bool inf = pts[i].ZZ.is_zero();
    if (is_inf) is_inf[i] = inf ? 1 : 0;
    if (inf) { memset(out_xy + 8 * i, 0, 64); continue; }
    const FqH x = fq::mul(pts[i].X, fq::mul(di, pts[i].ZZZ));
    const FqH y = fq::mul(pts[i].Y, fq::mul(di, pts[i].ZZ));
    memcpy(out_xy + 8 * i, x.l, 32);
    memcpy(out_xy + 8 * i + 4, y.l, 32);
  }
}

}  // namespace host
}  // namespace ja
