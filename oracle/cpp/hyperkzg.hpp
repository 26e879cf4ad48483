// TEST INFRASTRUCTURE ONLY — CPU oracle (C++ restatement) of the HyperKZG commit/open path.
//   joltworks/src/poly/commitment/hyperkzg/mod.rs:400-447 (open), :231-280 (kzg_open_batch),
//   :192-229 (kzg_batch_open_no_rem, compute_witness_polynomial), :520-554 (commit_one_hot)
//   joltworks/src/poly/commitment/hyperkzg/kzg.rs:227-243 (commit_variable_batch), :285-298 (commit_as_univariate)
//   joltworks/src/poly/unipoly.rs:247-305 (eval_as_univariate), dense_mlpoly.rs:444-499 (linear_combination)
// SRS: own test SRS g1_powers[i] = tau^i * G (the reference's ChaCha20/UniformRand sampling is external).
// Parity unpinned at the byte level; only pin upstream is the 368-byte proof length for l = 2.
#pragma once
#include "curve.hpp"
#include "poly.hpp"
#include "transcript.hpp"

namespace orc {

inline std::vector<G1Affine> srs_powers(const Fr& tau, size_t n) {
  std::vector<G1Affine> out(n);
  std::vector<Fr> pw(n);
  Fr t = Fr::one();
  for (size_t i = 0; i < n; i++) { pw[i] = t; t *= tau; }
  const G1Affine g = g1_generator();
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < n; i++) { uint64_t k[4]; pw[i].to_canonical(k); out[i] = to_affine(scalar_mul(g, k)); }
  return out;
}

inline Fr eval_as_univariate(const FrVec& f, const Fr& r) {   // unipoly.rs:247-257 (serial Horner-by-powers; exact sum)
  const size_t n = f.size();
  const int nt = omp_get_max_threads();
  std::vector<Fr> part(nt, Fr::zero());
  const size_t chunk = (n + nt - 1) / nt;
#pragma omp parallel
  {
    const int t = omp_get_thread_num();
    size_t lo = (size_t)t * chunk, hi = lo + chunk; if (hi > n) hi = n;
    if (lo < hi) {
      // r^lo
      Fr pw = Fr::one(), base = r; size_t e = lo;
      while (e) { if (e & 1) pw *= base; base = base.sqr(); e >>= 1; }
      Fr acc = Fr::zero();
      for (size_t i = lo; i < hi; i++) { acc += pw * f[i]; pw *= r; }
      part[t] = acc;
    }
  }
  Fr tot = Fr::zero();
  for (auto& p : part) tot += p;
  return tot;
}

inline FrVec witness_polynomial(const FrVec& f, const Fr& u) {   // mod.rs:213-229
  const size_t d = f.size();
  FrVec h(d, Fr::zero());
  for (size_t i = d - 1; i >= 1; i--) h[i - 1] = f[i] + h[i] * u;
  return h;
}

struct HyperKZGProof {
  std::vector<G1Affine> com, w;
  std::vector<std::vector<Fr>> v;   // 3 x ell
};

inline void append_points(Transcript& t, const std::vector<G1Affine>& pts) {   // blake2b.rs:189-195
  t.append_message("begin_append_vector");
  for (auto& p : pts) t.append_point(p.inf, p.x, p.y);
  t.append_message("end_append_vector");
}

inline HyperKZGProof hyperkzg_open(const std::vector<G1Affine>& srs, const FrVec& poly, const std::vector<Fr>& point, Transcript& t) {
  const size_t ell = point.size();
  std::vector<FrVec> polys; polys.push_back(poly);
  for (size_t i = 0; i + 1 < ell; i++) {
    const FrVec& prev = polys[i];
    FrVec pi(prev.size() / 2);
    const Fr x = point[ell - i - 1];
#pragma omp parallel for if (pi.size() >= 1024)
    for (size_t j = 0; j < pi.size(); j++) pi[j] = x * (prev[2 * j + 1] - prev[2 * j]) + prev[2 * j];
    polys.push_back(std::move(pi));
  }
  HyperKZGProof pf;
  for (size_t i = 1; i < ell; i++) pf.com.push_back(to_affine(msm_fr(srs.data(), polys[i].data(), polys[i].size())));
  append_points(t, pf.com);
  const Fr r = t.challenge_scalar();
  const Fr u[3] = {r, -r, r * r};
  pf.v.assign(3, std::vector<Fr>(ell));
  for (int i = 0; i < 3; i++) for (size_t j = 0; j < ell; j++) pf.v[i][j] = eval_as_univariate(polys[j], u[i]);
  std::vector<Fr> flat; for (auto& row : pf.v) flat.insert(flat.end(), row.begin(), row.end());
  t.append_scalars(flat);
  std::vector<Fr> q = t.challenge_scalar_powers(ell);
  FrVec B(poly.size(), Fr::zero());
  for (size_t k = 0; k < ell; k++) {
    const FrVec& f = polys[k];
#pragma omp parallel for if (f.size() >= 1024)
    for (size_t j = 0; j < f.size(); j++) B[j] += q[k] * f[j];
  }
  for (int i = 0; i < 3; i++) {
    FrVec h = witness_polynomial(B, u[i]);
    pf.w.push_back(to_affine(msm_fr(srs.data(), h.data(), h.size())));
  }
  append_points(t, pf.w);
  (void)t.challenge_scalar();
  return pf;
}

}  // namespace orc
