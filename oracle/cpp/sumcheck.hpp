// TEST INFRASTRUCTURE ONLY — CPU oracle (C++ restatement) of the sumcheck drivers and exemplar instances.
//   joltworks/src/subprotocols/sumcheck.rs:565-599 (Sumcheck::prove), :30-184 (BatchedSumcheck::prove)
//   jolt-atlas-core/src/onnx_proof/ops/{add.rs:283-304, sub.rs:267, mul.rs:160-185, square.rs:163, cube.rs:159-166}
//   jolt-atlas-core/src/onnx_proof/ops/einsum/dot.rs:290-375 (EqSchedule::None and the 3-MLE cubic)
//   joltworks/src/subprotocols/mles_product_sum.rs:15-129, :330-376
//   joltworks/src/subprotocols/hamming_weight.rs:118-139
// Parity unpinned at the byte level (no reference KATs); cross-checked against oracle/pyref.
#pragma once
#include <memory>
#include "poly.hpp"
#include "transcript.hpp"

namespace orc {

struct Instance {
  virtual ~Instance() {}
  virtual size_t num_rounds() const = 0;
  virtual size_t degree() const = 0;
  virtual Fr input_claim() const = 0;
  virtual UniPoly compute_message(size_t round, const Fr& previous_claim) = 0;
  virtual void ingest_challenge(const Fr& r, size_t round) = 0;
  virtual std::vector<Fr> final_claims() const = 0;
};

// mles_product_sum.rs:330-376
inline UniPoly finish_mles_product_sum_from_evals(const std::vector<Fr>& sum_evals, const Fr& claim, const GruenSplitEq& eq) {
  const Fr r = eq.current_w();
  const Fr eq0 = Fr::one() - r, eq1 = r;
  Fr at0 = sum_evals.size() == 1 ? claim - eq1 * sum_evals[0] : (claim - eq1 * sum_evals[0]) * eq0.inv();
  std::vector<Fr> toom; toom.push_back(at0);
  toom.insert(toom.end(), sum_evals.begin(), sum_evals.end());
  std::vector<Fr> tmp = UniPoly::from_evals_toom(toom).coeffs;
  const Fr cc = Fr::one() - r, xc = r + r - Fr::one();
  std::vector<Fr> coeffs(tmp.size() + 1, Fr::zero());
  for (size_t i = 0; i < tmp.size(); i++) { coeffs[i] += tmp[i] * cc; coeffs[i + 1] += tmp[i] * xc; }
  return UniPoly::from_coeff(coeffs);
}

// gruen_poly_deg_3 / deg_2 (split_eq_poly.rs:379-471)
inline UniPoly gruen_poly_deg_3(const GruenSplitEq& eq, const Fr& q_constant, const Fr& q_quadratic, const Fr& s01) {
  Fr eq1 = eq.current_scalar * eq.current_w();
  Fr eq0 = eq.current_scalar - eq1;
  Fr eqm = eq1 - eq0, eq2 = eq1 + eqm, eq3 = eq2 + eqm;
  Fr c0 = eq0 * q_constant, c1 = s01 - c0;
  Fr q1 = c1 * eq1.inv();
  Fr e2 = q_quadratic + q_quadratic;
  Fr q2 = q1 + q1 - q_constant + e2;
  Fr q3 = q2 + q1 - q_constant + e2 + e2;
  return UniPoly::from_evals({c0, c1, eq2 * q2, eq3 * q3});
}
inline UniPoly gruen_poly_deg_2(const GruenSplitEq& eq, const Fr& q0, const Fr& prev) {
  Fr eq1 = eq.current_scalar * eq.current_w();
  Fr eq0 = eq.current_scalar - eq1;
  Fr eqm = eq1 - eq0, eq2 = eq1 + eqm;
  Fr c0 = eq0 * q0, c1 = prev - c0;
  Fr l1 = c1 * eq1.inv();
  Fr l2 = l1 + l1 - q0;
  return UniPoly::from_evals({c0, c1, eq2 * l2});
}

enum SKind { S_ADD = 0, S_SUB = 1, S_MUL = 2, S_SQUARE = 3, S_PROD = 4, S_POW = 5, S_IDENT = 6,
             S_IFF = 8,      // jolt-atlas-core/src/onnx_proof/ops/iff.rs:189-216
             S_DIV = 9,      // ops/div.rs:329-347
             S_RSQRT = 10,   // ops/rsqrt.rs:390-418  (aux = gamma, S^3)
             S_LIN3 = 11 };  // neural_teleport/division.rs:231-246  (aux = tau)

// Family S: split-eq weighted, LowToHigh
struct SplitEqInstance : Instance {
  int kind; unsigned pow_d;
  GruenSplitEq eq;
  std::vector<FrVec> polys;
  Fr claim;
  std::vector<Fr> aux;      // scalars of the body (RSQRT: gamma, S^3; LIN3: tau)
  SplitEqInstance(int kind_, const Fr* w, size_t m, std::vector<FrVec> p, Fr claim_, unsigned pow_d_ = 0, std::vector<Fr> aux_ = {})
      : kind(kind_), pow_d(pow_d_), eq(w, m, LOW_TO_HIGH), polys(std::move(p)), claim(claim_), aux(std::move(aux_)) {}
  size_t num_rounds() const override { return eq.w.size(); }
  size_t degree() const override {
    switch (kind) { case S_ADD: case S_SUB: case S_IDENT: case S_LIN3: return 2; case S_MUL: case S_SQUARE: case S_IFF: case S_DIV: case S_RSQRT: return 3;
                    case S_POW: return pow_d + 1; default: return polys.size() + 1; }
  }
  Fr input_claim() const override { return claim; }
  UniPoly compute_message(size_t, const Fr& prev) override {
    if (kind == S_ADD || kind == S_SUB) {
      Fr q[1];
      const FrVec &l = polys[0], &r = polys[1];
      if (kind == S_ADD) eq.fold<1>([&](size_t g, Fr* v) { v[0] = l[2 * g] + r[2 * g]; }, q);
      else eq.fold<1>([&](size_t g, Fr* v) { v[0] = l[2 * g] - r[2 * g]; }, q);
      return gruen_poly_deg_2(eq, q[0], prev);
    }
    if (kind == S_IDENT) {   // ps_shout / identity-RC cycle rounds, dense opening reduction: [p0]
      Fr q[1];
      const FrVec& z = polys[0];
      eq.fold<1>([&](size_t g, Fr* v) { v[0] = z[2 * g]; }, q);
      return gruen_poly_deg_2(eq, q[0], prev);
    }
    if (kind == S_MUL) {
      Fr q[2];
      const FrVec &l = polys[0], &r = polys[1];
      eq.fold<2>([&](size_t g, Fr* v) {
        Fr l0 = l[2 * g], r0 = r[2 * g];
        v[0] = l0 * r0; v[1] = (l[2 * g + 1] - l0) * (r[2 * g + 1] - r0); }, q);
      return gruen_poly_deg_3(eq, q[0], q[1], prev);
    }
    if (kind == S_SQUARE) {
      Fr q[2];
      const FrVec& o = polys[0];
      eq.fold<2>([&](size_t g, Fr* v) { Fr d = o[2 * g + 1] - o[2 * g]; v[0] = o[2 * g].sqr(); v[1] = d.sqr(); }, q);
      return gruen_poly_deg_3(eq, q[0], q[1], prev);
    }
    if (kind == S_IFF) {      // iff.rs:197-215, literally
      Fr q[2];
      const FrVec &mk = polys[0], &a = polys[1], &b = polys[2];
      eq.fold<2>([&](size_t g, Fr* v) {
        Fr mask0 = mk[2 * g], mask1 = mk[2 * g + 1], mask_inf = mask1 - mask0;
        Fr a0 = a[2 * g], a_inf = a[2 * g + 1] - a0;
        Fr b0 = b[2 * g], b_inf = b[2 * g + 1] - b0;
        v[0] = mask0 * a0 + (Fr::one() - mask0) * b0;
        Fr f = (Fr::one() - mask1) - (Fr::one() - mask0);
        v[1] = mask_inf * a_inf + f * b_inf; }, q);
      return gruen_poly_deg_3(eq, q[0], q[1], prev);
    }
    if (kind == S_DIV) {      // div.rs:335-346
      Fr q[2];
      const FrVec &l = polys[0], &r = polys[1], &qq = polys[2], &R = polys[3];
      eq.fold<2>([&](size_t g, Fr* v) {
        v[0] = (r[2 * g] * qq[2 * g]) + R[2 * g] - l[2 * g];
        v[1] = (r[2 * g + 1] - r[2 * g]) * (qq[2 * g + 1] - qq[2 * g]); }, q);
      return gruen_poly_deg_3(eq, q[0], q[1], prev);
    }
    if (kind == S_RSQRT) {    // rsqrt.rs:399-417
      Fr q[2];
      const FrVec &x = polys[0], &quot = polys[1], &out = polys[2], &dr = polys[3], &sr = polys[4];
      const Fr gamma = aux[0], s_cubed = aux[1];
      eq.fold<2>([&](size_t g, Fr* v) {
        Fr div0 = x[2 * g] * quot[2 * g] + dr[2 * g] - s_cubed;
        Fr sqrt0 = out[2 * g] * out[2 * g] + sr[2 * g] - quot[2 * g];
        Fr div_quad = (x[2 * g + 1] - x[2 * g]) * (quot[2 * g + 1] - quot[2 * g]);
        Fr dout = out[2 * g + 1] - out[2 * g];
        v[0] = div0 + gamma * sqrt0; v[1] = div_quad + gamma * (dout * dout); }, q);
      return gruen_poly_deg_3(eq, q[0], q[1], prev);
    }
    if (kind == S_LIN3) {     // neural_teleport/division.rs:238-245
      Fr q[1];
      const FrVec &in = polys[0], &qu = polys[1], &rem = polys[2];
      const Fr tau = aux[0];
      eq.fold<1>([&](size_t g, Fr* v) { v[0] = (tau * qu[2 * g]) + rem[2 * g] - in[2 * g]; }, q);
      return gruen_poly_deg_2(eq, q[0], prev);
    }
    const size_t d = kind == S_POW ? pow_d : polys.size();
    std::vector<Fr> sums(d);
    eq.fold_dyn(d, [&](size_t g, Fr* v) {
      // prod_i (p_i0 + X*dp_i) on the grid {1, ..., d-1, inf}
      for (size_t k = 0; k < d; k++) v[k] = Fr::one();
      for (size_t i = 0; i < d; i++) {
        const FrVec& z = kind == S_POW ? polys[0] : polys[i];
        Fr p0 = z[2 * g], dp = z[2 * g + 1] - p0;
        Fr cur = p0;
        for (size_t k = 0; k + 1 < d; k++) { cur += dp; v[k] *= cur; }   // X = k+1
        v[d - 1] *= dp;                                                   // X = inf
      }
    }, sums.data());
    for (auto& s : sums) s *= eq.current_scalar;
    return finish_mles_product_sum_from_evals(sums, prev, eq);
  }
  void ingest_challenge(const Fr& r, size_t) override {
    eq.bind(r);
    for (auto& p : polys) bind_poly(p, r, LOW_TO_HIGH);
  }
  std::vector<Fr> final_claims() const override { std::vector<Fr> f; for (auto& p : polys) f.push_back(p[0]); return f; }
};

// Family D: plain products at X in {0,2,3}, HighToLow
struct DotInstance : Instance {
  std::vector<FrVec> polys; Fr claim;
  DotInstance(std::vector<FrVec> p, Fr c) : polys(std::move(p)), claim(c) {}
  size_t num_rounds() const override { size_t k = 0; while ((size_t(1) << k) < polys[0].size()) k++; return k; }
  size_t degree() const override { return polys.size(); }
  Fr input_claim() const override { return claim; }
  UniPoly compute_message(size_t, const Fr& prev) override {
    const size_t half = polys[0].size() / 2, deg = polys.size();
    const int nt = omp_get_max_threads();
    std::vector<Fr> part((size_t)nt * deg, Fr::zero());
#pragma omp parallel if (half >= 256)
    {
      std::vector<Fr> acc(deg, Fr::zero()), prod(deg);
#pragma omp for schedule(static)
      for (size_t i = 0; i < half; i++) {
        for (size_t q = 0; q < deg; q++) {
          Fr a = polys[q][i], b = polys[q][i + half];
          Fr m = b - a, e = b;
          for (size_t k = 0; k < deg; k++) {
            Fr val = k == 0 ? a : (e += m, e);
            prod[k] = q == 0 ? val : prod[k] * val;
          }
        }
        for (size_t k = 0; k < deg; k++) acc[k] += prod[k];
      }
      for (size_t k = 0; k < deg; k++) part[(size_t)omp_get_thread_num() * deg + k] = acc[k];
    }
    std::vector<Fr> ev(deg, Fr::zero());
    for (size_t k = 0; k < deg; k++) for (int t = 0; t < nt; t++) ev[k] += part[(size_t)t * deg + k];
    return UniPoly::from_evals_and_hint(prev, ev);
  }
  void ingest_challenge(const Fr& r, size_t) override { for (auto& p : polys) bind_poly(p, r, HIGH_TO_LOW); }
  std::vector<Fr> final_claims() const override { std::vector<Fr> f; for (auto& p : polys) f.push_back(p[0]); return f; }
};

// hamming_weight.rs:118-139
struct HammingInstance : Instance {
  std::vector<FrVec> polys; std::vector<Fr> gammas; Fr claim;
  HammingInstance(std::vector<FrVec> p, std::vector<Fr> g, Fr c) : polys(std::move(p)), gammas(std::move(g)), claim(c) {}
  size_t num_rounds() const override { size_t k = 0; while ((size_t(1) << k) < polys[0].size()) k++; return k; }
  size_t degree() const override { return 1; }
  Fr input_claim() const override { return claim; }
  UniPoly compute_message(size_t, const Fr& prev) override {
    Fr acc = Fr::zero();
    for (size_t i = 0; i < polys.size(); i++) {
      Fr s = Fr::zero();
      for (size_t j = 0; j < polys[i].size() / 2; j++) s += polys[i][2 * j];
      acc += gammas[i] * s;
    }
    return UniPoly::from_evals_and_hint(prev, {acc});
  }
  void ingest_challenge(const Fr& r, size_t) override { for (auto& p : polys) bind_poly(p, r, LOW_TO_HIGH); }
  std::vector<Fr> final_claims() const override { std::vector<Fr> f; for (auto& p : polys) f.push_back(p[0]); return f; }
};

struct SumcheckProof {
  std::vector<std::vector<Fr>> compressed_polys;   // coeffs except linear term, per round
  std::vector<std::array<uint64_t, 4>> challenges; // {0,0,lo,hi}
  Fr final_claim;
};

// Sumcheck::prove (sumcheck.rs:565-599)
inline SumcheckProof sumcheck_prove(Instance& inst, Transcript& t) {
  SumcheckProof pf;
  const size_t n = inst.num_rounds();
  Fr prev = inst.input_claim();
  t.append_scalar(prev);
  for (size_t round = 0; round < n; round++) {
    UniPoly uni = inst.compute_message(round, prev);
    std::vector<Fr> cp = uni.compress();
    append_compressed(t, cp);
    std::array<uint64_t, 4> c; t.challenge_optimized(c.data());
    const Fr r = Fr::from_raw(c.data());
    prev = uni.evaluate(r);
    inst.ingest_challenge(r, round);
    pf.compressed_polys.push_back(cp); pf.challenges.push_back(c);
  }
  pf.final_claim = prev;
  return pf;
}

// BatchedSumcheck::prove (sumcheck.rs:30-184)
inline SumcheckProof batched_sumcheck_prove(std::vector<Instance*>& insts, Transcript& t, std::vector<Fr>* coeffs_out = nullptr) {
  SumcheckProof pf;
  size_t max_rounds = 0;
  for (auto* i : insts) if (i->num_rounds() > max_rounds) max_rounds = i->num_rounds();
  for (auto* i : insts) t.append_scalar(i->input_claim());
  std::vector<Fr> coeffs = t.challenge_vector(insts.size());
  std::vector<Fr> claims;
  for (auto* i : insts) claims.push_back(i->input_claim().mul_pow_2((unsigned)(max_rounds - i->num_rounds())));
  // The reference walks the instances sequentially and parallelises INSIDE each one (rayon).  With hundreds of small
  // instances (the opening reduction) that leaves the cores idle, so this CPU baseline additionally spreads the
  // instances of a round over the threads (nested regions then run serially) - it only makes the baseline faster.
  const bool par_inst = insts.size() >= 8;
  for (size_t round = 0; round < max_rounds; round++) {
    const size_t remaining = max_rounds - round;
    std::vector<UniPoly> unis(insts.size());
#pragma omp parallel for schedule(dynamic, 1) if (par_inst)
    for (size_t k = 0; k < insts.size(); k++) {
      const size_t nr = insts[k]->num_rounds();
      if (remaining > nr) unis[k] = UniPoly::from_coeff({insts[k]->input_claim().mul_pow_2((unsigned)(remaining - nr - 1))});
      else unis[k] = insts[k]->compute_message(round - (max_rounds - nr), claims[k]);
    }
    UniPoly batched = UniPoly::from_coeff({});
    for (size_t k = 0; k < insts.size(); k++) batched.add_assign(unis[k].scaled(coeffs[k]));
    std::vector<Fr> cp = batched.compress();
    append_compressed(t, cp);
    std::array<uint64_t, 4> c; t.challenge_optimized(c.data());
    const Fr r = Fr::from_raw(c.data());
    for (size_t k = 0; k < insts.size(); k++) claims[k] = unis[k].evaluate(r);
#pragma omp parallel for schedule(dynamic, 1) if (par_inst)
    for (size_t k = 0; k < insts.size(); k++) {
      const size_t nr = insts[k]->num_rounds();
      if (remaining <= nr) insts[k]->ingest_challenge(r, round - (max_rounds - nr));
    }
    pf.compressed_polys.push_back(cp); pf.challenges.push_back(c);
  }
  pf.final_claim = Fr::zero();
  for (size_t k = 0; k < insts.size(); k++) pf.final_claim += claims[k] * coeffs[k];
  if (coeffs_out) *coeffs_out = coeffs;
  return pf;
}

}  // namespace orc

// ---- RA one-hot checks: shared set-up + booleanity (TEST INFRASTRUCTURE ONLY, as the rest of this file) ---------------
namespace orc {

// compute_ra_evals (joltworks/src/subprotocols/shout.rs:549-598): G[i][k] = sum_{j : idx_i[j] == k} eq(r_cycle, j).
// idx[i][j] == 0xFFFFFFFF (None) contributes nothing.
inline std::vector<FrVec> compute_ra_evals(const std::vector<std::vector<uint32_t>>& idx, size_t K, const Fr* r_cycle, size_t log_t) {
  const FrVec eq = eq_evals(r_cycle, log_t);
  std::vector<FrVec> G(idx.size(), FrVec(K, Fr::zero()));
#pragma omp parallel for
  for (size_t i = 0; i < idx.size(); i++)
    for (size_t j = 0; j < idx[i].size(); j++)
      if (idx[i][j] != 0xffffffffu) G[i][idx[i][j]] += eq[j];
  return G;
}

// RaPolynomial materialisation (poly/ra_poly.rs:31-81): out[j] = table[idx[j]] (None -> 0)
inline FrVec ra_materialise(const std::vector<uint32_t>& idx, const FrVec& table) {
  FrVec out(idx.size());
  for (size_t j = 0; j < idx.size(); j++) out[j] = idx[j] == 0xffffffffu ? Fr::zero() : table[idx[j]];
  return out;
}

// BooleanitySumcheckProver (joltworks/src/subprotocols/booleanity.rs:153-372)
struct BooleanityInstance : Instance {
  size_t d, log_k, log_t;
  std::vector<Fr> gammas;
  GruenSplitEq B, D;
  std::vector<FrVec> G;
  std::vector<std::vector<uint32_t>> H_idx;
  std::vector<FrVec> H;
  FrVec F;                       // ExpandingTable, LowToHigh (utils/expanding_table.rs:62-75)
  Fr eq_r_r = Fr::zero();
  BooleanityInstance(std::vector<FrVec> G_, std::vector<std::vector<uint32_t>> idx, std::vector<Fr> gammas_,
                     const Fr* r_address, size_t log_k_, const Fr* r_cycle, size_t log_t_)
      : d(G_.size()), log_k(log_k_), log_t(log_t_), gammas(std::move(gammas_)), B(r_address, log_k_, LOW_TO_HIGH),
        D(r_cycle, log_t_, LOW_TO_HIGH), G(std::move(G_)), H_idx(std::move(idx)) { F = FrVec{Fr::one()}; }
  size_t num_rounds() const override { return log_k + log_t; }
  size_t degree() const override { return 3; }
  Fr input_claim() const override { return Fr::zero(); }
  UniPoly compute_message(size_t round, const Fr& prev) override {
    Fr q[2];
    if (round < log_k) {                                        // :193-252
      const size_t m = round + 1;
      B.fold<2>([&](size_t k_prime, Fr* v) {
        v[0] = Fr::zero(); v[1] = Fr::zero();
        for (size_t i = 0; i < d; i++) {
          Fr s0 = Fr::zero(), s1 = Fr::zero();
          for (size_t k = 0; k < (size_t(1) << m); k++) {
            const Fr Gk = G[i][(k_prime << m) + k];
            const size_t k_m = k >> (m - 1);
            const Fr Fk = F[k % (size_t(1) << (m - 1))];
            const Fr GF = Gk * Fk;
            const Fr e_inf = GF * Fk;
            if (k_m == 0) s0 += e_inf - GF;
            s1 += e_inf;
          }
          v[0] += gammas[i] * s0; v[1] += gammas[i] * s1;
        }
      }, q);
      return gruen_poly_deg_3(B, q[0], q[1], prev);
    }
    D.fold<2>([&](size_t j, Fr* v) {                            // :254-301
      v[0] = Fr::zero(); v[1] = Fr::zero();
      for (size_t i = 0; i < d; i++) {
        const Fr h0 = H[i][2 * j], b = H[i][2 * j + 1] - h0;
        v[0] += (gammas[i] * h0) * (h0 - Fr::one());
        v[1] += (gammas[i] * b) * b;
      }
    }, q);
    const Fr adjusted = prev * eq_r_r.inv();
    return gruen_poly_deg_3(D, q[0], q[1], adjusted).scaled(eq_r_r);
  }
  void ingest_challenge(const Fr& r, size_t round) override {   // :321-348
    if (round < log_k) {
      B.bind(r);
      const size_t len = F.size();
      F.resize(2 * len);
      for (size_t i = 0; i < len; i++) { F[len + i] = F[i] * r; F[i] -= F[len + i]; }
      if (round == log_k - 1) {
        eq_r_r = B.current_scalar;
        for (auto& ix : H_idx) H.push_back(ra_materialise(ix, F));
        G.clear();
      }
    } else {
      D.bind(r);
      for (auto& h : H) bind_poly(h, r, LOW_TO_HIGH);
    }
  }
  std::vector<Fr> final_claims() const override { std::vector<Fr> f; for (auto& h : H) f.push_back(h[0]); return f; }
};

}  // namespace orc

// ---- batched opening reduction instances (TEST INFRASTRUCTURE ONLY) ---------------------------------------------------
namespace orc {

// gruen q(0) for HighToLow binding (opening_reduction.rs:355-403, :630-673): sum over the FIRST half j < len/2 with
// j = (x_in << out_bits) | x_out, weights E_in[x_in] * E_out[x_out]
inline Fr open_q0(const GruenSplitEq& D, const FrVec& z) {
  const FrVec& eo = D.E_out(); const FrVec& ei = D.E_in();
  int out_bits = 0; while ((size_t(1) << out_bits) < eo.size()) out_bits++;
  Fr tot = Fr::zero();
  for (size_t xi = 0; xi < ei.size(); xi++) {
    Fr inner = Fr::zero();
    for (size_t xo = 0; xo < eo.size(); xo++) inner += eo[xo] * z[(xi << out_bits) | xo];
    tot += ei[xi] * inner;
  }
  return tot;
}

// DensePolynomialProverOpening (opening_reduction.rs:337-424): sum_j eq(r, j) P[j], HighToLow, degree 2
struct DenseOpeningInstance : Instance {
  GruenSplitEq D; FrVec poly; Fr claim;
  DenseOpeningInstance(const Fr* r, size_t m, FrVec p, Fr c) : D(r, m, HIGH_TO_LOW), poly(std::move(p)), claim(c) {}
  size_t num_rounds() const override { return D.w.size(); }
  size_t degree() const override { return 2; }
  Fr input_claim() const override { return claim; }
  UniPoly compute_message(size_t, const Fr& prev) override { return gruen_poly_deg_2(D, open_q0(D, poly), prev); }
  void ingest_challenge(const Fr& r, size_t) override { D.bind(r); bind_poly(poly, r, HIGH_TO_LOW); }
  std::vector<Fr> final_claims() const override { return {poly[0]}; }
};

// OneHotPolynomialProverOpening (opening_reduction.rs:503-723): log K address rounds (B = eq(r_address, .) bound
// HighToLow, expanding table F HighToLow, G as in compute_ra_evals over D.merge()), then log T cycle rounds over
// H[j] = F[idx[j]] scaled by eq(r_address, r'_address).
struct OneHotOpeningInstance : Instance {
  size_t log_k, log_t; Fr claim;
  FrVec B, F, G, H;
  std::vector<uint32_t> idx;
  GruenSplitEq D;
  OneHotOpeningInstance(std::vector<uint32_t> idx_, const Fr* r_address, size_t log_k_, const Fr* r_cycle, size_t log_t_, Fr c)
      : log_k(log_k_), log_t(log_t_), claim(c), idx(std::move(idx_)), D(r_cycle, log_t_, HIGH_TO_LOW) {
    B = eq_evals(r_address, log_k);
    F = FrVec{Fr::one()};
    const FrVec dm = D.merge();                                   // :541 D_coeffs_for_G
    G.assign(size_t(1) << log_k, Fr::zero());
    for (size_t j = 0; j < idx.size(); j++) if (idx[j] != 0xffffffffu) G[idx[j]] += dm[j];
  }
  size_t num_rounds() const override { return log_k + log_t; }
  size_t degree() const override { return 2; }
  Fr input_claim() const override { return claim; }
  UniPoly compute_message(size_t round, const Fr& prev) override {
    if (round < log_k) {                                          // :579-629
      const size_t nu = log_k - round, half = B.size() / 2;
      Fr e0 = Fr::zero(), e2 = Fr::zero();
      for (size_t kp = 0; kp < half; kp++) {
        const Fr b0 = B[kp], b1 = B[kp + half];
        const Fr b2 = b1 + (b1 - b0);                             // sumcheck_evals_array::<2>: evals at 0 and 2
        Fr i0 = Fr::zero(), i2 = Fr::zero();
        for (size_t k = kp; k < G.size(); k += half) {
          const size_t k_m = (k >> (nu - 1)) & 1;
          const Fr GF = G[k] * F[k >> nu];
          if (k_m == 0) { i0 += GF; i2 -= GF; } else { i2 += GF + GF; }
        }
        e0 += b0 * i0; e2 += b2 * i2;
      }
      return UniPoly::from_evals_and_hint(prev, {e0, e2});
    }
    const Fr ea = B[0];                                           // B.final_claim()
    return gruen_poly_deg_2(D, open_q0(D, H), prev * ea.inv()).scaled(ea);   // :667-672
  }
  void ingest_challenge(const Fr& r, size_t round) override {   // :677-718
    if (round < log_k) {
      bind_poly(B, r, HIGH_TO_LOW);
      FrVec nf(F.size() * 2);                                     // ExpandingTable::update HighToLow (expanding_table.rs:76-86)
      for (size_t i = 0; i < F.size(); i++) { const Fr e1 = r * F[i]; nf[2 * i] = F[i] - e1; nf[2 * i + 1] = e1; }
      F.swap(nf);
      if (round == log_k - 1) { H = ra_materialise(idx, F); G.clear(); }
    } else {
      D.bind(r);
      bind_poly(H, r, HIGH_TO_LOW);
    }
  }
  std::vector<Fr> final_claims() const override { return {H[0]}; }
};

}  // namespace orc
