// Host-side O(degree) glue of the sumcheck driver: round-polynomial assembly from the reduced sums the kernels
// return, compression, evaluation.  Tiny per-round scalar math between kernel launches (not a fallback for any kernel).
//   UniPoly::{from_evals, from_evals_and_hint, from_evals_toom, from_coeff, compress, evaluate}   unipoly.rs:39-153,219-245,307-318
//   GruenSplitEqPolynomial::{gruen_poly_deg_2, gruen_poly_deg_3}                                  split_eq_poly.rs:379-471
//   finish_mles_product_sum_from_evals                                                            mles_product_sum.rs:330-376
// Interpolation results are unique polynomials, so the closed forms / Lagrange sums below give the same coefficients
// as the reference's Gaussian elimination; only the TRIMMING rules change transcript bytes and are kept exactly:
// from_evals of 3 or 4 values keeps its length, the general path and from_coeff drop trailing zeros.
#pragma once
#include <climits>
#include <map>
#include <mutex>
#include <vector>
#include "fr_host.hpp"

namespace ja {
namespace host {

typedef std::vector<FrH> Coeffs;

static inline Coeffs trim(Coeffs c) {           // UniPoly::from_coeff (unipoly.rs:39-52)
  while (!c.empty() && c.back().is_zero()) c.pop_back();
  if (c.empty()) c.push_back(FR_ZERO);
  return c;
}

// coefficients of the unique polynomial of degree < m through (i, e[i]), i = 0..m-1 (no trimming)
static inline Coeffs interpolate_0_to_m(const FrH* e, size_t m) {
  // master(x) = prod_j (x - j), by straightforward convolution
  Coeffs master;
  master.assign(m + 1, FR_ZERO);
  master[0] = FR_ONE;
  for (size_t j = 0; j < m; j++) {
    const FrH fj = from_u64(j);
    Coeffs next(m + 1, FR_ZERO);
    for (size_t k = 0; k <= j; k++) {
      next[k + 1] = add(next[k + 1], master[k]);
      next[k] = sub(next[k], mul(master[k], fj));
    }
    master = next;
  }
  // factorials
  std::vector<FrH> fact(m, FR_ONE);
  for (size_t i = 1; i < m; i++) fact[i] = mul(fact[i - 1], from_u64(i));
  Coeffs out(m, FR_ZERO);
  for (size_t i = 0; i < m; i++) {
    if (e[i].is_zero()) continue;
    // w_i = e[i] / (i! * (m-1-i)! * (-1)^(m-1-i))
    FrH w = mul(e[i], inv(mul(fact[i], fact[m - 1 - i])));
    if ((m - 1 - i) & 1) w = neg(w);
    // L_i numerator = master / (x - i) by synthetic division
    const FrH fi = from_u64(i);
    FrH carry = FR_ZERO;
    for (size_t k = m; k-- > 0;) {
      carry = add(master[k + 1], mul(carry, fi));
      out[k] = add(out[k], mul(w, carry));
    }
  }
  return out;
}

// Interpolation is linear in the evaluations: coeffs = M_n * e.  The matrices depend only on n and are built once
// (from unit vectors through the routines above) and cached; per round the glue is then n^2 products, no inversion.
struct InterpCache {
  std::mutex mu;
  std::map<size_t, std::vector<FrH>> plain, toom;
};
static inline InterpCache& interp_cache() { static InterpCache c; return c; }
static inline Coeffs from_evals_toom_slow(const std::vector<FrH>& e);
static inline const std::vector<FrH>& interp_matrix(size_t n, bool toom) {
  InterpCache& c = interp_cache();
  std::lock_guard<std::mutex> lk(c.mu);
  auto& tab = toom ? c.toom : c.plain;
  auto it = tab.find(n);
  if (it != tab.end()) return it->second;
  std::vector<FrH> m(n * n, FR_ZERO);
  for (size_t i = 0; i < n; i++) {
    std::vector<FrH> unit(n, FR_ZERO);
    unit[i] = FR_ONE;
    const Coeffs col = toom ? from_evals_toom_slow(unit) : interpolate_0_to_m(unit.data(), n);
    for (size_t k = 0; k < n; k++) m[k * n + i] = col[k];
  }
  return tab.emplace(n, std::move(m)).first->second;
}
static inline Coeffs apply_matrix(const std::vector<FrH>& m, const std::vector<FrH>& e) {
  const size_t n = e.size();
  Coeffs out(n, FR_ZERO);
  for (size_t k = 0; k < n; k++) {
    FrH acc = FR_ZERO;
    for (size_t i = 0; i < n; i++)
      if (!m[k * n + i].is_zero() && !e[i].is_zero()) acc = add(acc, mul(m[k * n + i], e[i]));
    out[k] = acc;
  }
  return out;
}

static inline Coeffs from_evals(const std::vector<FrH>& e) {    // unipoly.rs:55-92,136-153
  const size_t n = e.size();
  static const FrH two_inv = inv(from_u64(2)), six_inv = inv(from_u64(6));
  if (n == 3) {
    const FrH c2 = mul(add(sub(sub(e[0], e[1]), e[1]), e[2]), two_inv);
    const FrH c1 = sub(sub(e[1], e[0]), c2);
    return Coeffs{e[0], c1, c2};
  }
  if (n == 4) {
    const FrH c3 = mul(add(sub(e[3], e[0]), mul(sub(e[1], e[2]), from_u64(3))), six_inv);
    const FrH c2 = sub(sub(sub(mul(add(sub(sub(e[0], e[1]), e[1]), e[2]), two_inv), c3), c3), c3);
    const FrH c1 = sub(sub(sub(e[1], e[0]), c2), c3);
    return Coeffs{e[0], c1, c2, c3};
  }
  return trim(apply_matrix(interp_matrix(n, false), e));
}
static inline Coeffs from_evals_and_hint(const FrH& hint, const std::vector<FrH>& evals) {   // unipoly.rs:96-101
  std::vector<FrH> e = evals;
  e.insert(e.begin() + 1, sub(hint, e[0]));
  return from_evals(e);
}
// Toom interpolation with a SMALL-INTEGER matrix.  p(x) = lead * prod_{j<m} (x - j) + (interpolant of e[0..m) on 0..m-1), so
//   c_k = (1 / (m-1)!) * sum_i N[k][i] e[i] + s_k * lead,   N[k][i] = (-1)^(m-1-i) C(m-1, i) [x^k] prod_{j != i} (x - j),
// s_k = [x^k] prod_j (x - j): N fits in 64 bits for m <= 17, so a row is m products of 256 x 64 bits accumulated in 320-bit
// integers (positive and negative entries apart) and three Montgomery products (two fold the 320-bit sum times 1/(m-1)!
// back into the field: lo * K + hi * K R; one is s_k * lead) instead of m + 1 of them.  The degree-17 round polynomial of
// a product of 16 MLEs drops from 289 to ~100 product-equivalents; this sits on the Fiat-Shamir critical path of every
// RA-check round.  Same unique polynomial, exact field arithmetic: same coefficients.
struct ToomInt {
  size_t m = 0;
  bool ok = false;
  std::vector<int64_t> N;          // m x m
  std::vector<FrH> s;              // s_k, Montgomery form
  std::vector<int64_t> s_int;      // s_k as integers (|s_k| <= (m-1)! < 2^63)
  FrH fact;                        // Mont((m-1)!)
  FrH k_lo, k_hi;                  // Mont(1 / (m-1)!) and the same times R
};
static inline const ToomInt& toom_int(size_t n /* values: m = n - 1 points and the leading coefficient */) {
  static std::mutex mu;
  static std::map<size_t, ToomInt> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(n);
  if (it != cache.end()) return it->second;
  ToomInt t;
  t.m = n - 1;
  const size_t m = t.m;
  if (m >= 2 && m <= 17) {
    typedef __int128 i128;
    bool ok = true;
    t.N.assign(m * m, 0);
    std::vector<i128> binom(m, 1);                       // C(m-1, i)
    for (size_t i = 1; i < m; i++) binom[i] = binom[i - 1] * (i128)(m - i) / (i128)i;
    for (size_t i = 0; i < m && ok; i++) {
      std::vector<i128> poly(1, 1);                      // prod_{j != i} (x - j)
      for (size_t j = 0; j < m; j++) {
        if (j == i) continue;
        std::vector<i128> nx(poly.size() + 1, 0);
        for (size_t k = 0; k < poly.size(); k++) { nx[k + 1] += poly[k]; nx[k] -= poly[k] * (i128)j; }
        poly.swap(nx);
      }
      for (size_t k = 0; k < m; k++) {
        i128 v = poly[k] * binom[i];
        if ((m - 1 - i) & 1) v = -v;
        if (v > (i128)INT64_MAX / 2 || v < -((i128)INT64_MAX / 2)) { ok = false; break; }
        t.N[k * m + i] = (int64_t)v;
      }
    }
    if (ok) {
      std::vector<i128> full(1, 1);                      // prod_j (x - j)
      for (size_t j = 0; j < m; j++) {
        std::vector<i128> nx(full.size() + 1, 0);
        for (size_t k = 0; k < full.size(); k++) { nx[k + 1] += full[k]; nx[k] -= full[k] * (i128)j; }
        full.swap(nx);
      }
      t.s.resize(m); t.s_int.resize(m);
      for (size_t k = 0; k < m && ok; k++) {
        const i128 v = full[k];
        if (v > (i128)INT64_MAX / 2 || v < -((i128)INT64_MAX / 2)) { ok = false; break; }
        t.s[k] = from_i64((int64_t)v); t.s_int[k] = (int64_t)v;
      }
      FrH fact = FR_ONE;
      for (size_t i = 2; i < m; i++) fact = mul(fact, from_u64(i));
      t.fact = fact;
      t.k_lo = inv(fact);
      t.k_hi = mul(t.k_lo, FR_R2);                       // K * R (Montgomery form of K R)
    }
    t.ok = ok;
  }
  return cache.emplace(n, std::move(t)).first->second;
}
// acc (5 limbs) += a (4 limbs) * w
static inline void mac_256x64(uint64_t acc[5], const uint64_t a[4], uint64_t w) {
  u128 c = 0;
  for (int j = 0; j < 4; j++) { c += (u128)a[j] * w + acc[j]; acc[j] = (uint64_t)c; c >>= 64; }
  acc[4] += (uint64_t)c;
}
// (320-bit integer X) * K mod p for the Montgomery constants k_lo = Mont(K), k_hi = Mont(K R): X = lo + hi 2^256 and the
// CIOS product takes an unreduced first operand < 2^256 (result < 2 p before its final subtraction)
// a0 * b * R^-1 for a single-limb first operand: the CIOS product without the multiply rows of the three zero limbs
static inline FrH mul_limb(uint64_t a0, const FrH& b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = (u128)a0 * b.l[i] + t[0];
    t[0] = (uint64_t)c; c >>= 64;
    for (int j = 1; j < 4; j++) { c += t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    const uint64_t m = t[0] * FR_INV;
    c = (u128)m * FR_P[0] + t[0]; c >>= 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * FR_P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  FrH r = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || geq_p(r.l)) sub_p(r.l);
  return r;
}
static inline FrH fold320(const uint64_t x[5], const FrH& k_lo, const FrH& k_hi) {
  const FrH lo = {{x[0], x[1], x[2], x[3]}};
  return x[4] ? add(mul(lo, k_lo), mul_limb(x[4], k_hi)) : mul(lo, k_lo);
}
// values at 0..n-2 and the leading coefficient (value "at infinity") -> n coefficients, no trimming (unipoly.rs:104-134)
static inline Coeffs from_evals_toom(const std::vector<FrH>& e) {
  const size_t n = e.size();
  const ToomInt& t = toom_int(n);
  if (!t.ok) return apply_matrix(interp_matrix(n, true), e);
  const size_t m = t.m;
  const FrH& lead = e[m];
  // c_k = K (sum_i N[k][i] e_i + s_k (m-1)! lead): the leading-coefficient term joins the integer accumulation (one product for
  // (m-1)! lead instead of one per row)
  const FrH lead_f = mul(lead, t.fact);
  Coeffs out(n);
  for (size_t k = 0; k < m; k++) {
    uint64_t pos[5] = {0, 0, 0, 0, 0}, ngt[5] = {0, 0, 0, 0, 0};
    const int64_t* row = t.N.data() + k * m;
    for (size_t i = 0; i < m; i++) {
      const int64_t w = row[i];
      if (w > 0) mac_256x64(pos, e[i].l, (uint64_t)w);
      else if (w < 0) mac_256x64(ngt, e[i].l, (uint64_t)(-w));
    }
    const int64_t sk = t.s_int[k];
    if (sk > 0) mac_256x64(pos, lead_f.l, (uint64_t)sk);
    else if (sk < 0) mac_256x64(ngt, lead_f.l, (uint64_t)(-sk));
    // |pos - ngt| as a 320-bit integer, one fold, sign applied in the field
    bool ge = true;
    for (int j = 4; j >= 0; j--) if (pos[j] != ngt[j]) { ge = pos[j] > ngt[j]; break; }
    const uint64_t* hi_v = ge ? pos : ngt;
    const uint64_t* lo_v = ge ? ngt : pos;
    uint64_t df[5];
    unsigned char br = 0;
    for (int j = 0; j < 5; j++) {
      const u128 d = (u128)hi_v[j] - lo_v[j] - br;
      df[j] = (uint64_t)d; br = (unsigned char)((d >> 64) & 1);
    }
    const FrH f = fold320(df, t.k_lo, t.k_hi);
    out[k] = ge ? f : neg(f);
  }
  out[m] = lead;
  return out;
}
static inline Coeffs from_evals_toom_slow(const std::vector<FrH>& e) {
  const size_t n = e.size();
  const FrH lead = e[n - 1];
  std::vector<FrH> low(n - 1);
  for (size_t i = 0; i + 1 < n; i++) {
    FrH pw = FR_ONE; const FrH x = from_u64(i);
    for (size_t k = 0; k + 1 < n; k++) pw = mul(pw, x);      // i^(n-1)
    low[i] = sub(e[i], mul(lead, pw));
  }
  Coeffs c = interpolate_0_to_m(low.data(), n - 1);
  c.push_back(lead);
  return c;
}
static inline FrH evaluate(const Coeffs& c, const FrH& r) {    // unipoly.rs:219-245
  FrH acc = c.back();                                            // Horner: one product per coefficient, same field value
  for (size_t i = c.size() - 1; i-- > 0;) acc = add(mul_chal(acc, r), c[i]);     // r is a 125-bit challenge on the round path (plain product otherwise)
  return acc;
}
static inline Coeffs compress(const Coeffs& c) {                 // unipoly.rs:307-318: everything but the linear term
  if (c.size() < 2) return c;
  Coeffs o; o.push_back(c[0]);
  o.insert(o.end(), c.begin() + 2, c.end());
  return o;
}

// split_eq_poly.rs:432-471
// `eq1_inv` = (current_scalar * current_w)^-1: it does not depend on the kernel's sums, so the driver computes it while
// the round-evaluation kernel is in flight.
static inline FrH gruen_eq1(const FrH& current_scalar, const FrH& current_w) { return mul(current_scalar, current_w); }
static inline Coeffs gruen_poly_deg_2(const FrH& current_scalar, const FrH& current_w, const FrH& q0, const FrH& prev,
                                      const FrH& eq1_inv) {
  const FrH eq1 = mul(current_scalar, current_w);
  const FrH eq0 = sub(current_scalar, eq1);
  const FrH eqm = sub(eq1, eq0), eq2 = add(eq1, eqm);
  const FrH c0 = mul(eq0, q0), c1 = sub(prev, c0);
  const FrH l1 = mul(c1, eq1_inv);
  const FrH l2 = sub(add(l1, l1), q0);
  return from_evals({c0, c1, mul(eq2, l2)});
}
// The same round polynomial from q(1) directly.  The engine carries the eq-free polynomial's own running claim
// n = claim / current_scalar (n' = q(r): s(r) = current_scalar' * q(r) exactly), so q(1) = (n - (1 - w) q(0)) / w needs
// 1/w only - known for every round when the instance is built (one batch inversion) - instead of one inversion of
// current_scalar * w per round on the Fiat-Shamir critical path.  Field arithmetic is exact: same coefficients.
static inline Coeffs gruen_poly_deg_2_q1(const FrH& current_scalar, const FrH& current_w, const FrH& q0, const FrH& prev, const FrH& l1) {
  const FrH eq1 = mul(current_scalar, current_w);
  const FrH eq0 = sub(current_scalar, eq1);
  const FrH eqm = sub(eq1, eq0), eq2 = add(eq1, eqm);
  const FrH c0 = mul(eq0, q0), c1 = sub(prev, c0);
  const FrH l2 = sub(add(l1, l1), q0);
  return from_evals({c0, c1, mul(eq2, l2)});
}
static inline Coeffs gruen_poly_deg_3_q1(const FrH& current_scalar, const FrH& current_w, const FrH& q_constant,
                                         const FrH& q_quadratic, const FrH& s01, const FrH& q1) {
  const FrH eq1 = mul(current_scalar, current_w);
  const FrH eq0 = sub(current_scalar, eq1);
  const FrH eqm = sub(eq1, eq0), eq2 = add(eq1, eqm), eq3 = add(eq2, eqm);
  const FrH c0 = mul(eq0, q_constant), c1 = sub(s01, c0);
  const FrH e2 = add(q_quadratic, q_quadratic);
  const FrH q2 = add(sub(add(q1, q1), q_constant), e2);
  const FrH q3 = add(add(sub(add(q2, q1), q_constant), e2), e2);
  return from_evals({c0, c1, mul(eq2, q2), mul(eq3, q3)});
}
// split_eq_poly.rs:379-426
static inline Coeffs gruen_poly_deg_3(const FrH& current_scalar, const FrH& current_w, const FrH& q_constant,
                                      const FrH& q_quadratic, const FrH& s01, const FrH& eq1_inv) {
  const FrH eq1 = mul(current_scalar, current_w);
  const FrH eq0 = sub(current_scalar, eq1);
  const FrH eqm = sub(eq1, eq0), eq2 = add(eq1, eqm), eq3 = add(eq2, eqm);
  const FrH c0 = mul(eq0, q_constant), c1 = sub(s01, c0);
  const FrH q1 = mul(c1, eq1_inv);
  const FrH e2 = add(q_quadratic, q_quadratic);
  const FrH q2 = add(sub(add(q1, q1), q_constant), e2);
  const FrH q3 = add(add(sub(add(q2, q1), q_constant), e2), e2);
  return from_evals({c0, c1, mul(eq2, q2), mul(eq3, q3)});
}
// mles_product_sum.rs:330-376: sums = values of the eq-free product polynomial on {1..d-1, inf} (already times
// current_scalar); recover its value at 0 from the claim, interpolate, multiply by the linear eq factor.
// `eq0_inv` = (1 - r)^-1, precomputable like gruen's eq1_inv.
static inline Coeffs finish_mles_product_sum_from_evals(const std::vector<FrH>& sum_evals, const FrH& claim, const FrH& r,
                                                        const FrH& eq0_inv) {
  const FrH eq1 = r;
  FrH at0 = sub(claim, mul(eq1, sum_evals[0]));
  if (sum_evals.size() != 1) at0 = mul(at0, eq0_inv);
  std::vector<FrH> toom; toom.push_back(at0);
  toom.insert(toom.end(), sum_evals.begin(), sum_evals.end());
  const Coeffs tmp = from_evals_toom(toom);
  // times the linear eq factor (1 - r) + (2 r - 1) X: t (1 - r) = t - t r and t (2 r - 1) = 2 t r - t, one product per coefficient
  Coeffs coeffs(tmp.size() + 1, FR_ZERO);
  for (size_t i = 0; i < tmp.size(); i++) {
    const FrH tr = mul(tmp[i], r);
    coeffs[i] = add(coeffs[i], sub(tmp[i], tr));
    coeffs[i + 1] = add(coeffs[i + 1], sub(add(tr, tr), tmp[i]));
  }
  return trim(coeffs);
}

}  // namespace host
}  // namespace ja
