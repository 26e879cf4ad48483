// BN254 G1 (y^2 = x^3 + 3 over Fq) group law for the MSM / point-sum kernels.
//
// Device replacement for the curve arithmetic the reference gets from ark-ec / jolt-optimizations
// (external, a16z/arkworks-algebra@76bb3a4) at joltworks/src/msm/mod.rs:27-181 and
// joltworks/src/poly/commitment/hyperkzg/mod.rs:520-596 (batch_g1_additions_multi).
//
// Accumulators use extended Jacobian ("XYZZ") coordinates: x = X/ZZ, y = Y/ZZZ with ZZ^3 = ZZZ^2;
// infinity is ZZ == 0.  Mixed addition of an affine base costs 8M + 2S, a full addition 12M + 2S,
// doubling 6M + 4S (EFD xyzz: madd-2008-s, add-2008-s, dbl-2008-s-1; a = 0).
// Every routine is COMPLETE for the prime-order group (handles infinity, P == Q and P == -Q),
// because the result must equal the reference's group element for every input, not just random ones.
#pragma once
#include "fp.cuh"

namespace ja {

struct alignas(16) G1Aff { Fq x, y; };                 // 64 B; infinity never stored in the SRS
struct alignas(16) G1X { Fq X, Y, ZZ, ZZZ; };          // 128 B

JA_DEV Fq fq_add(const Fq& a, const Fq& b) { return fp_add<FqParams>(a, b); }
JA_DEV Fq fq_sub(const Fq& a, const Fq& b) { return fp_sub<FqParams>(a, b); }
JA_DEV Fq fq_mul(const Fq& a, const Fq& b) { return fp_mul<FqParams>(a, b); }
JA_DEV Fq fq_sqr(const Fq& a) { return fp_sqr<FqParams>(a); }
JA_DEV Fq fq_dbl(const Fq& a) { return fp_add<FqParams>(a, a); }
JA_DEV Fq fq_neg(const Fq& a) { return fp_is_zero(a) ? a : fp_sub<FqParams>(fp_zero<FqParams>(), a); }

JA_DEV G1X g1x_inf() { G1X r; r.X = fp_zero<FqParams>(); r.Y = fp_zero<FqParams>(); r.ZZ = fp_zero<FqParams>(); r.ZZZ = fp_zero<FqParams>(); return r; }
JA_DEV bool g1x_is_inf(const G1X& p) { return fp_is_zero(p.ZZ); }
JA_DEV G1X g1x_from_aff(const G1Aff& p) { G1X r; r.X = p.x; r.Y = p.y; r.ZZ = fp_one<FqParams>(); r.ZZZ = fp_one<FqParams>(); return r; }

JA_DEV G1Aff g1aff_load(const G1Aff* p) {
  G1Aff r; r.x = fp_load(&p->x); r.y = fp_load(&p->y); return r;
}
JA_DEV G1X g1x_load(const G1X* p) {
  G1X r; r.X = fp_load(&p->X); r.Y = fp_load(&p->Y); r.ZZ = fp_load(&p->ZZ); r.ZZZ = fp_load(&p->ZZZ); return r;
}
JA_DEV void g1x_store(G1X* p, const G1X& v) {
  fp_store(&p->X, v.X); fp_store(&p->Y, v.Y); fp_store(&p->ZZ, v.ZZ); fp_store(&p->ZZZ, v.ZZZ);
}

// 2 * (affine p)
JA_DEV G1X g1x_dbl_aff(const G1Aff& p) {
  G1X r;
  Fq U = fq_dbl(p.y);
  Fq V = fq_sqr(U);
  Fq W = fq_mul(U, V);
  Fq S = fq_mul(p.x, V);
  Fq XX = fq_sqr(p.x);
  Fq M = fq_add(fq_dbl(XX), XX);
  r.X = fq_sub(fq_sub(fq_sqr(M), S), S);
  r.Y = fq_sub(fq_mul(M, fq_sub(S, r.X)), fq_mul(W, p.y));
  r.ZZ = V; r.ZZZ = W;
  return r;
}

JA_DEV G1X g1x_dbl(const G1X& p) {
  if (g1x_is_inf(p)) return p;
  G1X r;
  Fq U = fq_dbl(p.Y);
  Fq V = fq_sqr(U);
  Fq W = fq_mul(U, V);
  Fq S = fq_mul(p.X, V);
  Fq XX = fq_sqr(p.X);
  Fq M = fq_add(fq_dbl(XX), XX);
  r.X = fq_sub(fq_sub(fq_sqr(M), S), S);
  r.Y = fq_sub(fq_mul(M, fq_sub(S, r.X)), fq_mul(W, p.Y));
  r.ZZ = fq_mul(V, p.ZZ);
  r.ZZZ = fq_mul(W, p.ZZZ);
  return r;
}

// acc += (affine q); `neg` negates q first (signed-digit buckets)
JA_DEV void g1x_madd(G1X& acc, const G1Aff& q_in, bool neg) {
  G1Aff q = q_in;
  if (neg) q.y = fq_neg(q.y);
  if (g1x_is_inf(acc)) { acc = g1x_from_aff(q); return; }
  Fq U2 = fq_mul(q.x, acc.ZZ);
  Fq S2 = fq_mul(q.y, acc.ZZZ);
  Fq P = fq_sub(U2, acc.X);
  Fq R = fq_sub(S2, acc.Y);
  if (fp_is_zero(P)) {
    if (fp_is_zero(R)) acc = g1x_dbl_aff(q);
    else acc = g1x_inf();
    return;
  }
  Fq PP = fq_sqr(P);
  Fq PPP = fq_mul(P, PP);
  Fq Q = fq_mul(acc.X, PP);
  Fq X3 = fq_sub(fq_sub(fq_sub(fq_sqr(R), PPP), Q), Q);
  Fq Y3 = fq_sub(fq_mul(R, fq_sub(Q, X3)), fq_mul(acc.Y, PPP));
  acc.X = X3; acc.Y = Y3;
  acc.ZZ = fq_mul(acc.ZZ, PP);
  acc.ZZZ = fq_mul(acc.ZZZ, PPP);
}

// a += b (both XYZZ)
JA_DEV void g1x_add(G1X& a, const G1X& b) {
  if (g1x_is_inf(b)) return;
  if (g1x_is_inf(a)) { a = b; return; }
  Fq U1 = fq_mul(a.X, b.ZZ);
  Fq U2 = fq_mul(b.X, a.ZZ);
  Fq S1 = fq_mul(a.Y, b.ZZZ);
  Fq S2 = fq_mul(b.Y, a.ZZZ);
  Fq P = fq_sub(U2, U1);
  Fq R = fq_sub(S2, S1);
  if (fp_is_zero(P)) {
    if (fp_is_zero(R)) a = g1x_dbl(a);
    else a = g1x_inf();
    return;
  }
  Fq PP = fq_sqr(P);
  Fq PPP = fq_mul(P, PP);
  Fq Q = fq_mul(U1, PP);
  Fq X3 = fq_sub(fq_sub(fq_sub(fq_sqr(R), PPP), Q), Q);
  Fq Y3 = fq_sub(fq_mul(R, fq_sub(Q, X3)), fq_mul(S1, PPP));
  a.X = X3; a.Y = Y3;
  a.ZZ = fq_mul(fq_mul(a.ZZ, b.ZZ), PP);
  a.ZZZ = fq_mul(fq_mul(a.ZZZ, b.ZZZ), PPP);
}

// k * p for a small unsigned k (bucket-segment base weight), MSB-first double-and-add
JA_DEV G1X g1x_mul_small(const G1X& p, uint32_t k) {
  G1X r = g1x_inf();
  if (k == 0 || g1x_is_inf(p)) return r;
  for (int bit = 31 - __clz(k); bit >= 0; bit--) {
    r = g1x_dbl(r);
    if ((k >> bit) & 1) g1x_add(r, p);
  }
  return r;
}

// a^(q-2) (Fermat); inverse of 0 is 0
JA_DEV Fq fq_inv(const Fq& a) {
  const uint32_t e[8] = {FqParams::P0 - 2u, FqParams::P1, FqParams::P2, FqParams::P3,
                         FqParams::P4, FqParams::P5, FqParams::P6, FqParams::P7};
  Fq r = fp_one<FqParams>();
  for (int i = 253; i >= 0; i--) {
    r = fq_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) r = fq_mul(r, a);
  }
  return r;
}

// affine coordinates of p (undefined if p is infinity)
JA_DEV G1Aff g1x_to_aff(const G1X& p) {
  Fq i = fq_inv(fq_mul(p.ZZ, p.ZZZ));
  G1Aff r;
  r.x = fq_mul(p.X, fq_mul(i, p.ZZZ));   // X / ZZ
  r.y = fq_mul(p.Y, fq_mul(i, p.ZZ));    // Y / ZZZ
  return r;
}

}  // namespace ja
