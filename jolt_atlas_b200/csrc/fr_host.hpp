// Host-side BN254 Fr arithmetic (4 x u64 Montgomery limbs) for the O(1)-per-round glue the C ABI
// and the C++ host driver keep on the CPU: GruenSplitEqPolynomial::bind's scalar update
// (split_eq_poly.rs:331-372), gruen_poly_deg_2/3 (:379-471), UniPoly interpolation/evaluation
// (unipoly.rs), transcript scalar serialisation (blake2b.rs:138-146).
// This is product code (tiny scalar math between kernel launches), not a CPU fallback for any kernel.
#pragma once
#include <alloca.h>
#include <cstdint>
#include <cstring>

namespace ja {
namespace host {

typedef unsigned __int128 u128;

struct FrH {
  uint64_t l[4];
  bool operator==(const FrH& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
  bool operator!=(const FrH& o) const { return !(*this == o); }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
};

static const uint64_t FR_P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t FR_INV = 0xc2e1f593efffffffull;  // -p^-1 mod 2^64
static const FrH FR_ONE = {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}};
static const FrH FR_R2 = {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}};
static const FrH FR_ZERO = {{0, 0, 0, 0}};

static inline bool geq_p(const uint64_t* a) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > FR_P[i]) return true;
    if (a[i] < FR_P[i]) return false;
  }
  return true;
}
static inline void sub_p(uint64_t* a) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - FR_P[i] - (uint64_t)b;
    a[i] = (uint64_t)t;
    b = (t >> 64) & 1;
  }
}
static inline FrH add(const FrH& a, const FrH& b) {
  FrH r; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
  if (geq_p(r.l)) sub_p(r.l);
  return r;
}
static inline FrH sub(const FrH& a, const FrH& b) {
  FrH r; u128 br = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a.l[i] - b.l[i] - (uint64_t)br;
    r.l[i] = (uint64_t)t; br = (t >> 64) & 1;
  }
  if (br) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.l[i] + FR_P[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
  return r;
}
static inline FrH neg(const FrH& a) { return a.is_zero() ? a : sub(FR_ZERO, a); }
static inline FrH dbl(const FrH& a) { return add(a, a); }

// CIOS Montgomery product
static inline FrH mul(const FrH& a, const FrH& b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * FR_INV;
    c = (u128)m * FR_P[0] + t[0]; c >>= 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * FR_P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  FrH r = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || geq_p(r.l)) sub_p(r.l);
  return r;
}
// a * r for a 125-bit challenge r in its Montgomery form {0, 0, lo, hi} (field/challenge/mont_ark_u128.rs:25-92): the CIOS iterations of
// the two zero limbs are no-ops (t stays 0, m = 0), so the product is exactly mul(a, r) at half the cost
static inline FrH mul_chal(const FrH& a, const FrH& r) {
  if (r.l[0] | r.l[1]) return mul(a, r);
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 2; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a.l[j] * r.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * FR_INV;
    c = (u128)m * FR_P[0] + t[0]; c >>= 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * FR_P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  FrH o = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || geq_p(o.l)) sub_p(o.l);
  return o;
}
static inline FrH sqr(const FrH& a) { return mul(a, a); }
static inline FrH from_u64(uint64_t v) { FrH t = {{v, 0, 0, 0}}; return mul(t, FR_R2); }
static inline FrH from_i64(int64_t v) {
  return v < 0 ? neg(from_u64((uint64_t)(-(v + 1)) + 1)) : from_u64((uint64_t)v);
}
// Montgomery limbs -> canonical integer limbs
static inline void to_canonical(const FrH& a, uint64_t out[4]) {
  // one REDC pass (a * R^-1 mod p): the four reduction steps of the CIOS product without its multiply rows - every scalar the
  // transcript absorbs goes through here
  uint64_t t[5] = {a.l[0], a.l[1], a.l[2], a.l[3], 0};
  for (int i = 0; i < 4; i++) {
    const uint64_t m = t[0] * FR_INV;
    u128 c = (u128)m * FR_P[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * FR_P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = (uint64_t)(c >> 64);
  }
  FrH r = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || geq_p(r.l)) sub_p(r.l);
  memcpy(out, r.l, 32);
}
static inline FrH from_canonical(const uint64_t in[4]) { FrH t; memcpy(t.l, in, 32); return mul(t, FR_R2); }
static inline FrH pow(const FrH& a, const uint64_t e[4]) {
  FrH r = FR_ONE;
  for (int i = 255; i >= 0; i--) {
    r = sqr(r);
    if ((e[i / 64] >> (i % 64)) & 1) r = mul(r, a);
  }
  return r;
}
static inline FrH inv_fermat(const FrH& a) {  // a^(p-2); inverse of 0 is 0
  uint64_t e[4] = {FR_P[0] - 2, FR_P[1], FR_P[2], FR_P[3]};
  return pow(a, e);
}
// Inverse by the binary extended Euclidean algorithm on the Montgomery limbs read as a plain integer x = a*R:
// it yields x^-1 = a^-1 * R^-1 (mod p); one Montgomery product with R^3 turns that into a^-1 * R.
// ~5x faster than the Fermat ladder; the per-round field division of gruen_poly_deg_2/3 sits on the critical path
// between two kernel launches.  Inverse of 0 is 0 (as ark's `inverse().unwrap_or(zero)` call sites expect).
namespace detail {
static inline bool is_one(const uint64_t* a) { return a[0] == 1 && (a[1] | a[2] | a[3]) == 0; }
static inline bool is_even(const uint64_t* a) { return (a[0] & 1) == 0; }
static inline void shr1(uint64_t* a, uint64_t top) {
  a[0] = (a[0] >> 1) | (a[1] << 63); a[1] = (a[1] >> 1) | (a[2] << 63);
  a[2] = (a[2] >> 1) | (a[3] << 63); a[3] = (a[3] >> 1) | (top << 63);
}
static inline uint64_t add4(uint64_t* a, const uint64_t* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; a[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
static inline void sub4(uint64_t* a, const uint64_t* b) {
  u128 br = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - b[i] - (uint64_t)br; a[i] = (uint64_t)t; br = (t >> 64) & 1; }
}
static inline bool geq4(const uint64_t* a, const uint64_t* b) {
  for (int i = 3; i >= 0; i--) { if (a[i] > b[i]) return true; if (a[i] < b[i]) return false; }
  return true;
}
// halve x modulo p (x < p)
static inline void half_mod(uint64_t* x) {
  uint64_t top = 0;
  if (!is_even(x)) top = add4(x, FR_P);
  shr1(x, top);
}
}  // namespace detail
static inline FrH inv(const FrH& a) {
  if (a.is_zero()) return a;
  using namespace detail;
  uint64_t u[4], v[4], x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
  memcpy(u, a.l, 32); memcpy(v, FR_P, 32);
  while (!is_one(u) && !is_one(v)) {
    while (is_even(u)) { shr1(u, 0); half_mod(x1); }
    while (is_even(v)) { shr1(v, 0); half_mod(x2); }
    if (geq4(u, v)) { sub4(u, v); if (geq4(x1, x2)) sub4(x1, x2); else { sub4(x1, x2); add4(x1, FR_P); } }
    else            { sub4(v, u); if (geq4(x2, x1)) sub4(x2, x1); else { sub4(x2, x1); add4(x2, FR_P); } }
  }
  FrH r; memcpy(r.l, is_one(u) ? x1 : x2, 32);            // = a^-1 * R^-1 as a plain integer
  static const FrH R3 = mul(FR_R2, FR_R2);                // R^2 * R^2 * R^-1
  return mul(r, R3);
}
// Montgomery's trick: every element of v replaced by its inverse with ONE inversion and 3 (n - 1) products; zeros stay
// zero (same convention as inv).
static inline void batch_inv(FrH* v, size_t n) {
  if (n == 0) return;
  FrH* pre = static_cast<FrH*>(alloca(n * sizeof(FrH)));
  FrH run = FR_ONE;
  for (size_t i = 0; i < n; i++) { pre[i] = run; if (!v[i].is_zero()) run = mul(run, v[i]); }
  FrH iv = inv(run);
  for (size_t i = n; i-- > 0;) {
    if (v[i].is_zero()) continue;
    const FrH t = mul(iv, pre[i]);
    iv = mul(iv, v[i]);
    v[i] = t;
  }
}
// challenge limbs {0,0,lo,hi} are already a valid Montgomery representation (mont_ark_u128.rs:79-84)
static inline FrH from_limbs(const uint64_t* p) { FrH r; memcpy(r.l, p, 32); return r; }

}  // namespace host
}  // namespace ja
