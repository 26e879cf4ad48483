// TEST INFRASTRUCTURE ONLY — CPU oracle (C++ restatement) of the multilinear-polynomial layer.
// rayon par_iter -> OpenMP; same loop structure and the same delayed outer product as the reference.
//   joltworks/src/poly/eq_poly.rs:149-167 (evals_serial), :174-217 (cached / cached_rev), :225-252 (evals_parallel)
//   joltworks/src/poly/dense_mlpoly.rs:126-141 (bind HighToLow), :219-239 (bind LowToHigh), :265-305 (split_eq_evaluate)
//   joltworks/src/poly/split_eq_poly.rs:86-145, :331-372, :379-471, :473-493, :526-597
//   joltworks/src/poly/multilinear_polynomial.rs:873-905 (sumcheck_evals)
// Parity unpinned at the byte level (no reference KATs); cross-checked with oracle/pyref.
#pragma once
#include <omp.h>
#include <functional>
#include <vector>
#include "field.hpp"

namespace orc {

enum { LOW_TO_HIGH = 0, HIGH_TO_LOW = 1 };
typedef std::vector<Fr> FrVec;

// eq_poly.rs:225-252 evals_parallel (big-endian: r[0] = MSB)
inline FrVec eq_evals(const Fr* r, size_t m, Fr scale = Fr::one()) {
  FrVec ev(size_t(1) << m);
  ev[0] = scale;
  size_t size = 1;
  for (size_t jj = m; jj-- > 0;) {
    const Fr rj = r[jj];
#pragma omp parallel for if (size >= 4096)
    for (size_t i = 0; i < size; i++) {
      Fr y = ev[i] * rj;
      ev[i + size] = y;
      ev[i] = ev[i] - y;
    }
    size *= 2;
  }
  // the loop above produces index bit k <-> r[m-1-k] ... i.e. MSB <-> r[0], as evals_parallel does
  return ev;
}
// eq_poly.rs:174-194
inline std::vector<FrVec> eq_evals_cached(const Fr* r, size_t m) {
  std::vector<FrVec> out(m + 1);
  out[0] = FrVec{Fr::one()};
  for (size_t j = 0; j < m; j++) {
    out[j + 1].resize(size_t(2) << j);
    for (size_t i = 0; i < (size_t(1) << j); i++) {
      Fr s = out[j][i];
      out[j + 1][2 * i + 1] = s * r[j];
      out[j + 1][2 * i] = s - out[j + 1][2 * i + 1];
    }
  }
  return out;
}
// eq_poly.rs:198-217
inline std::vector<FrVec> eq_evals_cached_rev(const Fr* r, size_t m) {
  std::vector<FrVec> out(m + 1);
  out[0] = FrVec{Fr::one()};
  for (size_t j = 0; j < m; j++) {
    out[j + 1].resize(size_t(2) << j);
    const Fr rj = r[m - 1 - j];
    for (size_t i = 0; i < (size_t(1) << j); i++) {
      Fr s = out[j][i];
      out[j + 1][i + (size_t(1) << j)] = s * rj;
      out[j + 1][i] = s - out[j + 1][i + (size_t(1) << j)];
    }
  }
  return out;
}

// dense_mlpoly.rs:126-141 / :219-239
inline void bind_poly(FrVec& z, const Fr& r, int order) {
  const size_t n = z.size() / 2;
  if (order == HIGH_TO_LOW) {
#pragma omp parallel for if (n >= 4096)
    for (size_t i = 0; i < n; i++) {
      if (z[i] != z[i + n]) z[i] += r * (z[i + n] - z[i]);
    }
    z.resize(n);
  } else {
    FrVec out(n);
#pragma omp parallel for if (n >= 512)
    for (size_t i = 0; i < n; i++) {
      Fr m = z[2 * i + 1] - z[2 * i];
      out[i] = m.is_zero() ? z[2 * i] : z[2 * i] + r * m;
    }
    z.swap(out);
  }
}

// dense_mlpoly.rs:265-305: sum_{x1} eq1[x1] * sum_{x2} eq2[x2] * Z[x1*|eq2| + x2]
inline Fr evaluate(const FrVec& z, const Fr* r, size_t m) {
  const size_t mh = m / 2;
  FrVec e1 = eq_evals(r, mh), e2 = eq_evals(r + mh, m - mh);
  const size_t n1 = e1.size(), n2 = e2.size();
  std::vector<Fr> parts(n1);
#pragma omp parallel for if (z.size() >= 4096)
  for (size_t x1 = 0; x1 < n1; x1++) {
    Fr acc = Fr::zero();
    for (size_t x2 = 0; x2 < n2; x2++) acc += e2[x2] * z[x1 * n2 + x2];
    parts[x1] = e1[x1] * acc;
  }
  Fr tot = Fr::zero();
  for (auto& p : parts) tot += p;
  return tot;
}

struct UniPoly;  // unipoly.hpp

// split_eq_poly.rs:67-598
struct GruenSplitEq {
  int order = LOW_TO_HIGH;
  size_t current_index = 0;
  Fr current_scalar = Fr::one();
  FrVec w;
  std::vector<FrVec> E_in_vec, E_out_vec;

  GruenSplitEq() { E_in_vec = {FrVec{Fr::one()}}; E_out_vec = {FrVec{Fr::one()}}; }
  GruenSplitEq(const Fr* w_, size_t n, int order_, Fr scale = Fr::one()) : order(order_), current_scalar(scale), w(w_, w_ + n) {
    if (n == 0) { E_in_vec = {FrVec{Fr::one()}}; E_out_vec = {FrVec{Fr::one()}}; return; }
    const size_t m = n / 2;
    if (order == LOW_TO_HIGH) {
      E_out_vec = eq_evals_cached(w.data(), m);
      E_in_vec = eq_evals_cached(w.data() + m, n - 1 - m);
      current_index = n;
    } else {
      size_t n_in = m > n - 1 ? n - 1 : m;
      E_in_vec = eq_evals_cached_rev(w.data() + 1, n_in);
      E_out_vec = eq_evals_cached_rev(w.data() + 1 + n_in, n - 1 - n_in);
      current_index = 0;
    }
  }
  const FrVec& E_in() const { return E_in_vec.back(); }
  const FrVec& E_out() const { return E_out_vec.back(); }
  Fr current_w() const { return order == LOW_TO_HIGH ? w[current_index - 1] : w[current_index]; }
  void bind(const Fr& r) {
    const Fr wv = current_w();
    const Fr prod = wv * r;
    current_scalar *= Fr::one() - wv - r + prod + prod;
    const size_t n = w.size();
    if (order == LOW_TO_HIGH) {
      current_index -= 1;
      if (n / 2 < current_index && E_in_vec.size() > 1) E_in_vec.pop_back();
      else if (0 < current_index && E_out_vec.size() > 1) E_out_vec.pop_back();
    } else {
      current_index += 1;
      if (current_index <= n / 2 && E_in_vec.size() > 1) E_in_vec.pop_back();
      else if (current_index <= n && E_out_vec.size() > 1) E_out_vec.pop_back();
    }
  }
  FrVec merge() const {
    if (order == LOW_TO_HIGH) return eq_evals(w.data(), current_index, current_scalar);
    return eq_evals(w.data() + current_index, w.size() - current_index, current_scalar);
  }
  // par_fold_out_in_unreduced::<9, NUM_OUT> (:569-597): rayon over x_out -> OpenMP; inner x_in sequential.
  template <int NUM_OUT, class PerG>
  void fold(const PerG& per_g, Fr (&out)[NUM_OUT]) const {
    const FrVec& eo = E_out(); const FrVec& ei = E_in();
    const size_t out_len = eo.size(), in_len = ei.size();
    int bits_in = 0; while ((size_t(1) << bits_in) < in_len) bits_in++;
    const int nt = omp_get_max_threads();
    std::vector<Fr> part((size_t)nt * NUM_OUT, Fr::zero());
#pragma omp parallel if (out_len * in_len >= 256)
    {
      const int tid = omp_get_thread_num();
      Fr acc[NUM_OUT];
      for (int k = 0; k < NUM_OUT; k++) acc[k] = Fr::zero();
#pragma omp for schedule(static)
      for (size_t xo = 0; xo < out_len; xo++) {
        Fr inner[NUM_OUT];
        for (int k = 0; k < NUM_OUT; k++) inner[k] = Fr::zero();
        for (size_t xi = 0; xi < in_len; xi++) {
          Fr v[NUM_OUT];
          per_g((xo << bits_in) | xi, v);
          for (int k = 0; k < NUM_OUT; k++) inner[k] += ei[xi] * v[k];
        }
        for (int k = 0; k < NUM_OUT; k++) acc[k] += eo[xo] * inner[k];
      }
      for (int k = 0; k < NUM_OUT; k++) part[(size_t)tid * NUM_OUT + k] = acc[k];
    }
    for (int k = 0; k < NUM_OUT; k++) {
      out[k] = Fr::zero();
      for (int t = 0; t < nt; t++) out[k] += part[(size_t)t * NUM_OUT + k];
    }
  }
  // dynamic-width variant for the product-of-d rounds
  void fold_dyn(size_t num_out, const std::function<void(size_t, Fr*)>& per_g, Fr* out) const {
    const FrVec& eo = E_out(); const FrVec& ei = E_in();
    const size_t out_len = eo.size(), in_len = ei.size();
    int bits_in = 0; while ((size_t(1) << bits_in) < in_len) bits_in++;
    const int nt = omp_get_max_threads();
    std::vector<Fr> part((size_t)nt * num_out, Fr::zero());
#pragma omp parallel if (out_len * in_len >= 64)
    {
      const int tid = omp_get_thread_num();
      std::vector<Fr> acc(num_out, Fr::zero()), inner(num_out), v(num_out);
#pragma omp for schedule(static)
      for (size_t xo = 0; xo < out_len; xo++) {
        for (auto& x : inner) x = Fr::zero();
        for (size_t xi = 0; xi < in_len; xi++) {
          per_g((xo << bits_in) | xi, v.data());
          for (size_t k = 0; k < num_out; k++) inner[k] += ei[xi] * v[k];
        }
        for (size_t k = 0; k < num_out; k++) acc[k] += eo[xo] * inner[k];
      }
      for (size_t k = 0; k < num_out; k++) part[(size_t)tid * num_out + k] = acc[k];
    }
    for (size_t k = 0; k < num_out; k++) {
      out[k] = Fr::zero();
      for (int t = 0; t < nt; t++) out[k] += part[(size_t)t * num_out + k];
    }
  }
};

}  // namespace orc
