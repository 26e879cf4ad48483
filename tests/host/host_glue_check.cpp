// CPU check of the host-side round glue (sumcheck_host.hpp): the small-integer Toom interpolation, the one-product linear
// factor, Horner evaluation, batch inversion and the inversion-free Gruen round polynomials against their plain forms.
// Built and run by tests/test_host_glue.py (g++, no GPU).
#include <cstdio>
#include <random>
#include <vector>
#include "fr_host.hpp"
#include "sumcheck_host.hpp"
using namespace ja; using namespace ja::host;
int main() {
  std::mt19937_64 rng(7);
  auto rnd = [&]() { FrH x = {{rng(), rng(), rng(), rng() >> 3}}; return mul(x, FR_R2); };
  int bad = 0;
  for (size_t n = 2; n <= 19; n++) {
    for (int it = 0; it < 50; it++) {
      std::vector<FrH> e(n);
      for (auto& x : e) x = rnd();
      if (it == 1) for (auto& x : e) x = FR_ZERO;
      if (it == 2) { for (auto& x : e) x = FR_ZERO; e[n - 1] = rnd(); }
      if (it == 3) { for (auto& x : e) x = sub(FR_ZERO, FR_ONE); }
      Coeffs a = from_evals_toom(e);
      Coeffs b = apply_matrix(interp_matrix(n, true), e);
      if (a.size() != b.size()) { bad++; printf("size mismatch n=%zu\n", n); continue; }
      for (size_t k = 0; k < a.size(); k++) if (a[k] != b[k]) { bad++; printf("mismatch n=%zu k=%zu it=%d\n", n, k, it); break; }
    }
  }
  // finish_mles_product_sum vs the two-product form
  for (int it = 0; it < 100; it++) {
    std::vector<FrH> se(16); for (auto& x : se) x = rnd();
    FrH claim = rnd(), r = rnd(), q = rnd();
    Coeffs a = finish_mles_product_sum_from_evals(se, claim, r, q);
    // reference form
    FrH at0 = mul(sub(claim, mul(r, se[0])), q);
    std::vector<FrH> toom; toom.push_back(at0); toom.insert(toom.end(), se.begin(), se.end());
    Coeffs tmp = apply_matrix(interp_matrix(toom.size(), true), toom);
    const FrH cc = sub(FR_ONE, r), xc = sub(add(r, r), FR_ONE);
    Coeffs c(tmp.size() + 1, FR_ZERO);
    for (size_t i = 0; i < tmp.size(); i++) { c[i] = add(c[i], mul(tmp[i], cc)); c[i + 1] = add(c[i + 1], mul(tmp[i], xc)); }
    c = trim(c);
    if (a.size() != c.size()) { bad++; continue; }
    for (size_t k = 0; k < a.size(); k++) if (a[k] != c[k]) { bad++; break; }
  }
  // batch inversion vs single inversions (zeros stay zero)
  for (int it = 0; it < 50; it++) {
    std::vector<FrH> v(1 + it % 19), w;
    for (auto& x : v) x = rnd();
    if (it % 5 == 0) v[it % v.size()] = FR_ZERO;
    w = v;
    batch_inv(w.data(), w.size());
    for (size_t i = 0; i < v.size(); i++) if (w[i] != inv(v[i])) { bad++; printf("batch_inv mismatch it=%d i=%zu\n", it, i); break; }
  }
  // Horner evaluation vs the power form
  for (int it = 0; it < 50; it++) {
    Coeffs c(1 + it % 19); for (auto& x : c) x = rnd();
    const FrH r = rnd();
    FrH acc = c[0], pw = r;
    for (size_t i = 1; i < c.size(); i++) { acc = add(acc, mul(pw, c[i])); pw = mul(pw, r); }
    if (evaluate(c, r) != acc) { bad++; printf("evaluate mismatch it=%d\n", it); }
  }
  // gruen_poly_deg_{2,3}_q1 with q(1) from the normalised claim vs the per-round inversion form
  for (int it = 0; it < 100; it++) {
    const FrH cs = rnd(), cw = rnd(), q0 = rnd(), qq = rnd(), nclaim = rnd();
    const FrH prev = mul(nclaim, cs);
    const FrH eq1_inv = inv(gruen_eq1(cs, cw));
    const FrH q1 = mul(sub(nclaim, mul(sub(FR_ONE, cw), q0)), inv(cw));
    const Coeffs a2 = gruen_poly_deg_2(cs, cw, q0, prev, eq1_inv), b2 = gruen_poly_deg_2_q1(cs, cw, q0, prev, q1);
    const Coeffs a3 = gruen_poly_deg_3(cs, cw, q0, qq, prev, eq1_inv), b3 = gruen_poly_deg_3_q1(cs, cw, q0, qq, prev, q1);
    if (a2 != b2 || a3 != b3) { bad++; printf("gruen mismatch it=%d\n", it); }
    // the running claim: s(r) = cs' q(r)
    const FrH r = rnd();
    const FrH f = add(sub(sub(FR_ONE, cw), r), dbl(mul(cw, r)));
    const FrH n3 = add(q0, mul(r, add(sub(sub(q1, q0), qq), mul(r, qq))));
    if (evaluate(a3, r) != mul(mul(cs, f), n3)) { bad++; printf("running claim (deg 3) mismatch it=%d\n", it); }
    const FrH n2 = add(q0, mul(r, sub(q1, q0)));
    if (evaluate(a2, r) != mul(mul(cs, f), n2)) { bad++; printf("running claim (deg 2) mismatch it=%d\n", it); }
  }
  // round-2 primitives: the challenge-aware product, the single-limb product, the one-pass REDC and Horner on a 125-bit challenge
  for (int it = 0; it < 2000; it++) {
    const FrH a = it == 0 ? FR_ZERO : (it == 1 ? sub(FR_ZERO, FR_ONE) : rnd());
    FrH ch = {{0, 0, rng(), rng() >> 3}};                       // Montgomery limbs {0, 0, lo, hi} of a challenge
    if (it == 2) ch = FR_ZERO;
    if (it == 3) ch = {{0, 0, ~0ull, ~0ull >> 3}};
    if (mul_chal(a, ch) != mul(a, ch)) { bad++; printf("mul_chal mismatch it=%d\n", it); }
    const FrH full = rnd();
    if (mul_chal(a, full) != mul(a, full)) { bad++; printf("mul_chal (full operand) mismatch it=%d\n", it); }
    const uint64_t x = it == 4 ? ~0ull : (it == 5 ? 0 : rng());
    const FrH xl = {{x, 0, 0, 0}};
    if (mul_limb(x, a) != mul(xl, a)) { bad++; printf("mul_limb mismatch it=%d\n", it); }
    uint64_t c1[4]; to_canonical(a, c1);
    const FrH one = {{1, 0, 0, 0}};
    const FrH c2 = mul(a, one);
    if (c1[0] != c2.l[0] || c1[1] != c2.l[1] || c1[2] != c2.l[2] || c1[3] != c2.l[3]) { bad++; printf("to_canonical mismatch it=%d\n", it); }
    if (from_canonical(c1) != a) { bad++; printf("canonical round trip mismatch it=%d\n", it); }
    std::vector<FrH> c(1 + it % 18); for (auto& y : c) y = rnd();
    FrH acc = c[0], pw = ch;
    for (size_t i = 1; i < c.size(); i++) { acc = add(acc, mul(pw, c[i])); pw = mul(pw, ch); }
    if (evaluate(c, ch) != acc) { bad++; printf("evaluate on a challenge mismatch it=%d\n", it); }
  }
  printf("bad=%d\n", bad);
  return bad != 0;
}
