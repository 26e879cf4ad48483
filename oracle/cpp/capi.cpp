// TEST INFRASTRUCTURE ONLY — C entry points of the CPU oracle for ctypes (tests/, __graft_entry__.smoke(),
// bench.py's cpu_baseline / --impl reference leg).  The product library never links or loads this.
#include <omp.h>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "sumcheck.hpp"
#include "hyperkzg.hpp"
#include "psshout.hpp"

using namespace orc;

static inline FrVec load_fr(const uint64_t* p, size_t n) { FrVec v(n); memcpy((void*)v.data(), p, n * 32); return v; }
static inline void store_fr(uint64_t* p, const Fr& x) { memcpy(p, x.l, 32); }
static inline G1Affine load_pt(const uint64_t* p, bool inf = false) { G1Affine a; memcpy(a.x.l, p, 32); memcpy(a.y.l, p + 4, 32); a.inf = inf; return a; }
static inline void store_pt(uint64_t* p, int32_t* inf, const G1Affine& a) { memcpy(p, a.x.l, 32); memcpy(p + 4, a.y.l, 32); if (inf) *inf = a.inf ? 1 : 0; }

extern "C" {

int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_threads(int n) { omp_set_num_threads(n); }

void orc_blake2b256(const uint8_t* data, size_t n, uint8_t out[32]) { Blake2b256 h; h.update(data, n); h.finalize(out); }

// field ops on arrays (for cross-checking the arithmetic itself): op 0 add, 1 sub, 2 mul, 3 mul by challenge limbs
void orc_fr_binop(int op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
  for (size_t i = 0; i < n; i++) {
    Fr x = Fr::from_raw(a + 4 * i), y = Fr::from_raw(b + 4 * i), r;
    r = op == 0 ? x + y : op == 1 ? x - y : x * y;
    store_fr(out + 4 * i, r);
  }
}
void orc_fr_from_i64(const int64_t* v, size_t n, uint64_t* out) { for (size_t i = 0; i < n; i++) store_fr(out + 4 * i, Fr::from_i64(v[i])); }

void orc_bind(uint64_t* z, size_t n, const uint64_t r[4], int order) {
  FrVec v = load_fr(z, n);
  bind_poly(v, Fr::from_raw(r), order);
  memcpy(z, v.data(), v.size() * 32);
}
void orc_eq_evals(const uint64_t* r, size_t m, const uint64_t* scale, uint64_t* out) {
  FrVec rr = load_fr(r, m);
  FrVec ev = eq_evals(rr.data(), m, scale ? Fr::from_raw(scale) : Fr::one());
  memcpy(out, ev.data(), ev.size() * 32);
}
void orc_evaluate(const uint64_t* z, size_t n, const uint64_t* point, size_t m, uint64_t out[4]) {
  FrVec v = load_fr(z, n), p = load_fr(point, m);
  store_fr(out, evaluate(v, p.data(), m));
}

// Einsum operand fold (jolt-atlas-core/src/onnx_proof/ops/einsum/mk_kn_mn.rs:47-79): i32 matrix x eq-vector.
//   transpose == 0: out[j] = sum_i from_i32(A[i*cols+j]) * eq[i];  transpose == 1: out[i] = sum_j from_i32(A[i*cols+j]) * eq[j]
void orc_tensor_fold_i32(const int32_t* A, size_t rows, size_t cols, const uint64_t* eq, int transpose, uint64_t* out) {
  const size_t n_eq = transpose ? cols : rows, n_out = transpose ? rows : cols;
  FrVec e = load_fr(eq, n_eq);
  FrVec o(n_out, Fr::zero());
  if (transpose) {
#pragma omp parallel for if (rows * cols >= 4096)
    for (size_t i = 0; i < rows; i++) {
      Fr acc = Fr::zero();
      for (size_t j = 0; j < cols; j++) { int32_t v = A[i * cols + j]; if (v) acc += Fr::from_i64(v) * e[j]; }
      o[i] = acc;
    }
  } else {
#pragma omp parallel for if (rows * cols >= 4096)
    for (size_t j = 0; j < cols; j++) {
      Fr acc = Fr::zero();
      for (size_t i = 0; i < rows; i++) { int32_t v = A[i * cols + j]; if (v) acc += Fr::from_i64(v) * e[i]; }
      o[j] = acc;
    }
  }
  memcpy(out, o.data(), n_out * 32);
}

// Sumcheck::prove over one instance.
//   family 0 (split-eq, LowToHigh): kind = SKind, w = m Fr;  family 1 (dot, HighToLow): w ignored.
// Outputs: coeffs[rounds][max_coeffs][4] (compressed: all but the linear term), ncoeffs[rounds], challenges[rounds][4],
// final_claims[npoly][4], state[32] (transcript state after the last challenge).
static int sumcheck_prove_t(int family, int kind, unsigned pow_d, const uint64_t* polys, size_t npoly, size_t n,
                            const uint64_t* w, size_t m, const uint64_t claim[4], const uint64_t* gammas, Transcript& t,
                            size_t max_coeffs, uint64_t* coeffs, uint32_t* ncoeffs, uint64_t* challenges,
                            uint64_t* final_claims);
int orc_sumcheck_prove(int family, int kind, unsigned pow_d, const uint64_t* polys, size_t npoly, size_t n,
                       const uint64_t* w, size_t m, const uint64_t claim[4], const char* label,
                       size_t max_coeffs, uint64_t* coeffs, uint32_t* ncoeffs, uint64_t* challenges,
                       uint64_t* final_claims, uint8_t state[32]) {
  Transcript t(label);
  int rc = sumcheck_prove_t(family, kind, pow_d, polys, npoly, n, w, m, claim, nullptr, t, max_coeffs, coeffs, ncoeffs, challenges, final_claims);
  memcpy(state, t.state, 32);
  return rc;
}
// same, resuming a running transcript (state + round counter read and written back); family 2 = Hamming (gammas)
int orc_sumcheck_prove_st(int family, int kind, unsigned pow_d, const uint64_t* polys, size_t npoly, size_t n,
                          const uint64_t* w, size_t m, const uint64_t claim[4], const uint64_t* gammas,
                          uint8_t state[32], uint32_t* n_rounds,
                          size_t max_coeffs, uint64_t* coeffs, uint32_t* ncoeffs, uint64_t* challenges, uint64_t* final_claims) {
  Transcript t(state, *n_rounds);
  int rc = sumcheck_prove_t(family, kind, pow_d, polys, npoly, n, w, m, claim, gammas, t, max_coeffs, coeffs, ncoeffs, challenges, final_claims);
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
  return rc;
}
static int sumcheck_prove_t(int family, int kind, unsigned pow_d, const uint64_t* polys, size_t npoly, size_t n,
                            const uint64_t* w, size_t m, const uint64_t claim[4], const uint64_t* gammas, Transcript& t,
                            size_t max_coeffs, uint64_t* coeffs, uint32_t* ncoeffs, uint64_t* challenges,
                            uint64_t* final_claims) {
  std::vector<FrVec> ps;
  for (size_t i = 0; i < npoly; i++) ps.push_back(load_fr(polys + 4 * n * i, n));
  std::unique_ptr<Instance> inst;
  if (family == 0) {
    FrVec ww = load_fr(w, m);
    std::vector<Fr> aux;      // RSQRT: gamma, S^3; LIN3: tau (passed in `gammas`)
    if (gammas && (kind == S_RSQRT || kind == S_LIN3)) aux = load_fr(gammas, kind == S_RSQRT ? 2 : 1);
    inst.reset(new SplitEqInstance(kind, ww.data(), m, std::move(ps), Fr::from_raw(claim), pow_d, aux));
  }
  else if (family == 1) inst.reset(new DotInstance(std::move(ps), Fr::from_raw(claim)));
  else {
    std::vector<Fr> g(npoly, Fr::one());
    if (gammas) for (size_t i = 0; i < npoly; i++) g[i] = Fr::from_raw(gammas + 4 * i);
    inst.reset(new HammingInstance(std::move(ps), g, Fr::from_raw(claim)));
  }
  SumcheckProof pf = sumcheck_prove(*inst, t);
  for (size_t r = 0; r < pf.compressed_polys.size(); r++) {
    if (pf.compressed_polys[r].size() > max_coeffs) return -1;
    ncoeffs[r] = (uint32_t)pf.compressed_polys[r].size();
    for (size_t k = 0; k < pf.compressed_polys[r].size(); k++) store_fr(coeffs + 4 * (r * max_coeffs + k), pf.compressed_polys[r][k]);
    memcpy(challenges + 4 * r, pf.challenges[r].data(), 32);
  }
  std::vector<Fr> fc = inst->final_claims();
  for (size_t i = 0; i < fc.size(); i++) store_fr(final_claims + 4 * i, fc[i]);
  return (int)pf.compressed_polys.size();
}

// ---- BatchedSumcheck::prove over a list of instance descriptors (sumcheck.rs:30-184) --------------------------------
// kind: the JA_EVAL_* ids of include/jolt_atlas_b200.h (0 ADD, 1 SUB, 2 MUL, 3 SQUARE, 4 PROD, 5 POW, 6 IDENT, 16 DOT2, 17 DOT3,
// 18 SUM1 = Hamming weight with gammas) and 32 = Booleanity (polys = G tables d x K, idx = d x T addresses,
// eq_w = r_cycle, aux_fr = gammas (d) then r_address (log_k), aux_u32 = log_k).
struct orc_inst {
  int32_t kind; uint32_t aux_u32;
  uint64_t n_polys, poly_len;
  const uint64_t* polys; const uint32_t* idx;
  const uint64_t* eq_w; uint64_t eq_m;
  const uint64_t* aux_fr; uint64_t n_aux;
  uint64_t claim[4];
  uint64_t* final_claims;
};
static Instance* make_instance(const orc_inst& d) {
  std::vector<FrVec> ps;
  if (d.polys) for (size_t i = 0; i < d.n_polys; i++) ps.push_back(load_fr(d.polys + 4 * d.poly_len * i, d.poly_len));
  const Fr claim = Fr::from_raw(d.claim);
  if (d.kind <= 11) {
    FrVec w = load_fr(d.eq_w, d.eq_m);
    std::vector<Fr> aux;
    if (d.aux_fr && d.kind >= 8) aux = load_fr(d.aux_fr, d.n_aux);
    return new SplitEqInstance(d.kind, w.data(), d.eq_m, std::move(ps), claim, d.aux_u32, aux);
  }
  if (d.kind == 16 || d.kind == 17) return new DotInstance(std::move(ps), claim);
  if (d.kind == 18) {
    std::vector<Fr> g(d.n_polys, Fr::one());
    if (d.aux_fr) for (size_t i = 0; i < d.n_polys; i++) g[i] = Fr::from_raw(d.aux_fr + 4 * i);
    return new HammingInstance(std::move(ps), g, claim);
  }
  if (d.kind == 20) {   // dense opening (HighToLow): eq_w = opening point
    FrVec w = load_fr(d.eq_w, d.eq_m);
    return new DenseOpeningInstance(w.data(), d.eq_m, std::move(ps[0]), claim);
  }
  if (d.kind == 34) {   // one-hot opening: idx = T addresses, eq_w = r_cycle, aux_fr = r_address (log_k), aux_u32 = log_k
    const size_t log_k = d.aux_u32, T = size_t(1) << d.eq_m;
    std::vector<uint32_t> idx(d.idx, d.idx + T);
    FrVec ra = load_fr(d.aux_fr, log_k), rc = load_fr(d.eq_w, d.eq_m);
    return new OneHotOpeningInstance(std::move(idx), ra.data(), log_k, rc.data(), d.eq_m, claim);
  }
  if (d.kind == 32) {
    const size_t log_k = d.aux_u32, dd = d.n_polys, T = size_t(1) << d.eq_m;
    std::vector<std::vector<uint32_t>> idx(dd);
    for (size_t i = 0; i < dd; i++) idx[i].assign(d.idx + i * T, d.idx + (i + 1) * T);
    std::vector<Fr> gam(dd);
    for (size_t i = 0; i < dd; i++) gam[i] = Fr::from_raw(d.aux_fr + 4 * i);
    FrVec ra = load_fr(d.aux_fr + 4 * dd, log_k), rc = load_fr(d.eq_w, d.eq_m);
    return new BooleanityInstance(std::move(ps), std::move(idx), std::move(gam), ra.data(), log_k, rc.data(), d.eq_m);
  }
  return nullptr;
}
int orc_batched_sumcheck_prove(const orc_inst* descs, size_t n, uint8_t state[32], uint32_t* n_rounds, size_t max_coeffs,
                               uint64_t* coeffs, uint32_t* ncoeffs, uint64_t* challenges) {
  std::vector<std::unique_ptr<Instance>> own(n);
  std::vector<Instance*> insts(n);
#pragma omp parallel for schedule(dynamic, 1) if (n >= 8)
  for (size_t i = 0; i < n; i++) own[i].reset(make_instance(descs[i]));
  for (size_t i = 0; i < n; i++) { if (!own[i]) return -2; insts[i] = own[i].get(); }
  Transcript t(state, *n_rounds);
  SumcheckProof pf = batched_sumcheck_prove(insts, t);
  for (size_t r = 0; r < pf.compressed_polys.size(); r++) {
    if (pf.compressed_polys[r].size() > max_coeffs) return -1;
    ncoeffs[r] = (uint32_t)pf.compressed_polys[r].size();
    for (size_t k = 0; k < pf.compressed_polys[r].size(); k++) store_fr(coeffs + 4 * (r * max_coeffs + k), pf.compressed_polys[r][k]);
    memcpy(challenges + 4 * r, pf.challenges[r].data(), 32);
  }
  for (size_t i = 0; i < n; i++) {
    std::vector<Fr> fc = insts[i]->final_claims();
    if (descs[i].final_claims) for (size_t k = 0; k < fc.size(); k++) store_fr(descs[i].final_claims + 4 * k, fc[k]);
  }
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
  return (int)pf.compressed_polys.size();
}
// ---- prefix-suffix Shout, T-sized passes (psshout.hpp) ------------------------------------------------------------------------
void* orc_psshout_new(const uint64_t* idx, size_t T, const uint64_t* r_cycle, size_t log_t, unsigned log_k, unsigned phases) {
  PsShout* p = new PsShout();
  p->idx.assign(idx, idx + T);
  FrVec r = load_fr(r_cycle, log_t);
  p->u = eq_evals(r.data(), log_t);                 // u_evals = EqPolynomial::evals(r_node_output), mod.rs:236
  p->log_k = log_k; p->phases = phases; p->log_m = log_k / phases;
  return p;
}
void orc_psshout_init_phase(void* h, unsigned phase, const uint64_t* v_prev, const uint32_t* kinds, size_t n_suf, unsigned bound, uint64_t* out_Q) {
  PsShout* p = static_cast<PsShout*>(h);
  const size_t m = size_t(1) << p->log_m;
  FrVec v = v_prev ? load_fr(v_prev, m) : FrVec();
  std::vector<Fr> Q = p->init_phase(phase, v_prev ? v.data() : nullptr, kinds, n_suf, bound);
  for (size_t i = 0; i < Q.size(); i++) store_fr(out_Q + 4 * i, Q[i]);
}
void orc_psshout_materialize_ra(void* h, const uint64_t* v, uint64_t* out) {
  PsShout* p = static_cast<PsShout*>(h);
  FrVec vv = load_fr(v, (size_t)p->phases << p->log_m);
  std::vector<Fr> ra = p->materialize_ra(vv.data());
  for (size_t i = 0; i < ra.size(); i++) store_fr(out + 4 * i, ra[i]);
}
void orc_psshout_free(void* h) { delete static_cast<PsShout*>(h); }
// the LOG_K address rounds of Sumcheck::prove over the read-raf instance (subprotocols/sumcheck.rs:565-599 with
// ps_shout/mod.rs:464-560): compressed round polynomials [c0, c2], challenges, expanding tables, val, raf_val, running claim
void orc_psshout_prove_address(void* h, unsigned bound, const uint64_t* gamma, const uint64_t* claim_in, uint8_t state[32], uint32_t* n_rounds,
                               uint64_t* out_coeffs, uint32_t* out_ncoeffs, uint64_t* out_challenges, uint64_t* out_v, uint64_t* out_val,
                               uint64_t* out_raf_val, uint64_t* out_claim) {
  PsShout* p = static_cast<PsShout*>(h);
  PsReadRaf rr;
  rr.ps = p; rr.XLEN = p->log_k; rr.BOUND = bound; rr.gamma = Fr::from_raw(gamma);
  rr.v.resize(p->phases);
  Transcript t(state, *n_rounds);
  rr.init_phase(0);
  Fr claim = claim_in ? Fr::from_raw(claim_in) : rr.derived_input_claim();
  for (unsigned round = 0; round < p->log_k; round++) {
    Fr e[2];
    rr.message(round, e);
    UniPoly uni = UniPoly::from_evals_and_hint(claim, {e[0], e[1]});
    std::vector<Fr> cp = uni.compress();
    append_compressed(t, cp);
    uint64_t ch[4];
    t.challenge_optimized(ch);
    const Fr rj = Fr::from_raw(ch);
    claim = uni.evaluate(rj);
    rr.ingest(rj, round);
    out_ncoeffs[round] = (uint32_t)cp.size();
    for (size_t k = 0; k < cp.size() && k < 2; k++) store_fr(out_coeffs + 4 * (2 * round + k), cp[k]);
    memcpy(out_challenges + 4 * round, ch, 32);
  }
  const size_t m = size_t(1) << p->log_m;
  for (unsigned ph = 0; ph < p->phases; ph++)
    for (size_t i = 0; i < m; i++) store_fr(out_v + 4 * (ph * m + i), rr.v[ph][i]);
  store_fr(out_val, rr.val); store_fr(out_raf_val, rr.raf_val); store_fr(out_claim, claim);
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}
void orc_psshout_prove_identity_rc(void* h, const uint64_t* claim_in, uint8_t state[32], uint32_t* n_rounds, uint64_t* out_coeffs,
                                   uint32_t* out_ncoeffs, uint64_t* out_challenges, uint64_t* out_v, uint64_t* out_raf_val, uint64_t* out_claim) {
  PsShout* p = static_cast<PsShout*>(h);
  PsIdentityRC rc;
  rc.ps = p; rc.v.resize(p->phases);
  Transcript t(state, *n_rounds);
  rc.init_phase(0);
  Fr claim = claim_in ? Fr::from_raw(claim_in) : rc.derived_input_claim();
  for (unsigned round = 0; round < p->log_k; round++) {
    Fr e[2];
    rc.message(e);
    UniPoly uni = UniPoly::from_evals_and_hint(claim, {e[0], e[1]});
    std::vector<Fr> cp = uni.compress();
    append_compressed(t, cp);
    uint64_t ch[4];
    t.challenge_optimized(ch);
    const Fr rj = Fr::from_raw(ch);
    claim = uni.evaluate(rj);
    rc.ingest(rj, round);
    out_ncoeffs[round] = (uint32_t)cp.size();
    for (size_t k = 0; k < cp.size() && k < 2; k++) store_fr(out_coeffs + 4 * (2 * round + k), cp[k]);
    memcpy(out_challenges + 4 * round, ch, 32);
  }
  const size_t m = size_t(1) << p->log_m;
  for (unsigned ph = 0; ph < p->phases; ph++)
    for (size_t i = 0; i < m; i++) store_fr(out_v + 4 * (ph * m + i), rc.v[ph][i]);
  store_fr(out_raf_val, rc.raf_val); store_fr(out_claim, claim);
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}
void orc_clamp_evaluate_mle(const uint64_t* r, unsigned xlen, unsigned bound, uint64_t* out) {
  FrVec rr = load_fr(r, xlen);
  store_fr(out, clamp_evaluate_mle(rr.data(), xlen, bound));
}
void orc_signed_identity_evaluate(const uint64_t* r, unsigned n, uint64_t* out) {
  FrVec rr = load_fr(r, n);
  store_fr(out, signed_identity_evaluate(rr.data(), n));
}
uint64_t orc_suffix_mle(int kind, uint64_t bits, unsigned len, unsigned xlen, unsigned bound) { return suffix_mle(kind, bits, len, xlen, bound); }

// compute_ra_evals (shout.rs:549-598): idx = d x T addresses, out = d x K Fr
void orc_compute_ra_evals(const uint32_t* idx, size_t d, size_t T, size_t K, const uint64_t* r_cycle, size_t log_t, uint64_t* out) {
  std::vector<std::vector<uint32_t>> ix(d);
  for (size_t i = 0; i < d; i++) ix[i].assign(idx + i * T, idx + (i + 1) * T);
  FrVec rc = load_fr(r_cycle, log_t);
  std::vector<FrVec> G = compute_ra_evals(ix, K, rc.data(), log_t);
  for (size_t i = 0; i < d; i++) memcpy(out + 4 * K * i, G[i].data(), K * 32);
}

// build_materialized_rlc (poly/rlc_polynomial.rs:13-78), in place on `joint` (n Fr):
//   one-hot batch: joint[idx[i][t] * T + t] += coeffs[i];   dense: joint[i] += coeff * poly[i]
void orc_rlc_add_onehot(uint64_t* joint, const uint32_t* idx, size_t d, size_t T, const uint64_t* coeffs) {
  Fr* J = reinterpret_cast<Fr*>(joint);
  for (size_t i = 0; i < d; i++) {
    const Fr c = Fr::from_raw(coeffs + 4 * i);
    for (size_t t = 0; t < T; t++) { const uint32_t k = idx[i * T + t]; if (k != 0xffffffffu) J[(size_t)k * T + t] += c; }
  }
}
void orc_rlc_add_dense(uint64_t* joint, const uint64_t* poly, size_t len, const uint64_t coeff[4]) {
  Fr* J = reinterpret_cast<Fr*>(joint);
  const Fr c = Fr::from_raw(coeff);
#pragma omp parallel for if (len >= 4096)
  for (size_t i = 0; i < len; i++) J[i] += c * Fr::from_raw(poly + 4 * i);
}

// compute_h (subprotocols/evaluation_reduction.rs:223-249), restated as the reference does it: fold the coefficient
// table variable by variable with polynomial-valued entries (Add / Sub keep lengths, Mul trims: unipoly.rs:401-476).
int orc_eval_reduction_h(const uint64_t* mle, size_t len, const uint64_t* points, size_t n, size_t m, uint64_t* out, size_t cap) {
  typedef std::vector<Fr> Poly;
  auto trim = [](Poly p) { while (!p.empty() && p.back().is_zero()) p.pop_back(); if (p.empty()) p.push_back(Fr::zero()); return p; };
  auto addp = [](const Poly& a, const Poly& b) { Poly c(std::max(a.size(), b.size()), Fr::zero()); for (size_t i = 0; i < a.size(); i++) c[i] += a[i]; for (size_t i = 0; i < b.size(); i++) c[i] += b[i]; return c; };
  auto subp = [](const Poly& a, const Poly& b) { Poly c(std::max(a.size(), b.size()), Fr::zero()); for (size_t i = 0; i < a.size(); i++) c[i] += a[i]; for (size_t i = 0; i < b.size(); i++) c[i] -= b[i]; return c; };
  auto mulp = [&](const Poly& a, const Poly& b) { Poly c(a.size() + b.size() - 1, Fr::zero()); for (size_t i = 0; i < a.size(); i++) for (size_t j = 0; j < b.size(); j++) c[i + j] += a[i] * b[j]; return trim(c); };
  std::vector<Poly> tab(len);
  for (size_t i = 0; i < len; i++) tab[i] = trim(Poly{Fr::from_raw(mle + 4 * i)});
  for (size_t i = 0; i < m; i++) {
    const size_t half = size_t(1) << (m - i - 1);
    std::vector<Fr> e(n);
    for (size_t j = 0; j < n; j++) e[j] = Fr::from_raw(points + 4 * (j * m + i));
    const Poly var = UniPoly::from_evals(e).coeffs;
    for (size_t j = 0; j < half; j++) tab[j] = addp(tab[j], mulp(var, subp(tab[j + half], tab[j])));
  }
  if (tab[0].size() > cap) return -1;
  for (size_t i = 0; i < tab[0].size(); i++) store_fr(out + 4 * i, tab[0][i]);
  return (int)tab[0].size();
}

void orc_transcript_append_scalars(uint8_t state[32], uint32_t* n_rounds, const uint64_t* fr, size_t n) {
  Transcript t(state, *n_rounds);
  t.append_scalars(load_fr(fr, n));
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}
// n x Transcript::append_scalar: cache_openings of one instance (poly/opening_proof.rs:281, :338, :398)
void orc_transcript_append_scalar_each(uint8_t state[32], uint32_t* n_rounds, const uint64_t* fr, size_t n) {
  Transcript t(state, *n_rounds);
  for (size_t i = 0; i < n; i++) t.append_scalar(Fr::from_raw(fr + 4 * i));
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}
void orc_transcript_challenge_scalar_powers(uint8_t state[32], uint32_t* n_rounds, size_t n, uint64_t* out) {
  Transcript t(state, *n_rounds);
  std::vector<Fr> q = t.challenge_scalar_powers(n);
  for (size_t i = 0; i < n; i++) store_fr(out + 4 * i, q[i]);
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}

void orc_transcript_challenge_optimized(uint8_t state[32], uint32_t* n_rounds, size_t n, uint64_t* out) {
  Transcript t(state, *n_rounds);
  for (size_t i = 0; i < n; i++) t.challenge_optimized(out + 4 * i);       // blake2b.rs:233-238
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}
// ExpandingTable::update, HighToLow (utils/expanding_table.rs:76-86), n updates from [1]
void orc_expanding_table_h2l(const uint64_t* challenges, size_t n, uint64_t* out) {
  std::vector<Fr> v{Fr::one()};
  for (size_t j = 0; j < n; j++) {
    const Fr r = Fr::from_raw(challenges + 4 * j);
    std::vector<Fr> nv(v.size() * 2);
    for (size_t i = 0; i < v.size(); i++) { const Fr e1 = r * v[i]; nv[2 * i] = v[i] - e1; nv[2 * i + 1] = e1; }
    v.swap(nv);
  }
  for (size_t i = 0; i < v.size(); i++) store_fr(out + 4 * i, v[i]);
}

// ---- curve / MSM ----
void orc_srs_powers(const uint64_t tau_mont[4], size_t n, uint64_t* out_xy) {
  std::vector<G1Affine> s = srs_powers(Fr::from_raw(tau_mont), n);
  for (size_t i = 0; i < n; i++) store_pt(out_xy + 8 * i, nullptr, s[i]);
}
static std::vector<G1Affine> load_bases(const uint64_t* xy, size_t n) {
  std::vector<G1Affine> b(n);
  for (size_t i = 0; i < n; i++) b[i] = load_pt(xy + 8 * i);
  return b;
}
void orc_msm_fr(const uint64_t* bases_xy, const uint64_t* scalars, size_t n, uint64_t out_xy[8], int32_t* inf) {
  std::vector<G1Affine> b = load_bases(bases_xy, n);
  FrVec s = load_fr(scalars, n);
  store_pt(out_xy, inf, to_affine(msm_fr(b.data(), s.data(), n)));
}
void orc_msm_i64(const uint64_t* bases_xy, const int64_t* scalars, size_t n, uint64_t out_xy[8], int32_t* inf) {
  std::vector<G1Affine> b = load_bases(bases_xy, n);
  store_pt(out_xy, inf, to_affine(msm_i64(b.data(), scalars, n)));
}
void orc_sum_indexed(const uint64_t* bases_xy, size_t n_bases, const uint64_t* idx, size_t n, uint64_t out_xy[8], int32_t* inf) {
  std::vector<G1Affine> b = load_bases(bases_xy, n_bases);
  store_pt(out_xy, inf, to_affine(sum_indexed(b.data(), idx, n)));
}
int orc_on_curve(const uint64_t xy[8]) { return on_curve(load_pt(xy)) ? 1 : 0; }
void orc_scalar_mul(const uint64_t xy[8], const uint64_t k_mont[4], uint64_t out_xy[8], int32_t* inf) {
  uint64_t k[4]; Fr::from_raw(k_mont).to_canonical(k);
  store_pt(out_xy, inf, to_affine(scalar_mul(load_pt(xy), k)));
}

// HyperKZG::open.  point = ell challenge limbs.  Outputs: com[(ell-1)][8] + com_inf, w[3][8] + w_inf, v[3][ell][4], state[32].
static void hyperkzg_open_t(const uint64_t* srs_xy, size_t n, const uint64_t* poly, const uint64_t* point, size_t ell,
                            Transcript& t, uint64_t* com_xy, int32_t* com_inf, uint64_t* w_xy, int32_t* w_inf, uint64_t* v);
void orc_hyperkzg_open(const uint64_t* srs_xy, size_t n, const uint64_t* poly, const uint64_t* point, size_t ell,
                       const char* label, uint64_t* com_xy, int32_t* com_inf, uint64_t* w_xy, int32_t* w_inf,
                       uint64_t* v, uint8_t state[32]) {
  Transcript t(label);
  hyperkzg_open_t(srs_xy, n, poly, point, ell, t, com_xy, com_inf, w_xy, w_inf, v);
  memcpy(state, t.state, 32);
}
void orc_hyperkzg_open_st(const uint64_t* srs_xy, size_t n, const uint64_t* poly, const uint64_t* point, size_t ell,
                          uint8_t state[32], uint32_t* n_rounds, uint64_t* com_xy, int32_t* com_inf, uint64_t* w_xy,
                          int32_t* w_inf, uint64_t* v) {
  Transcript t(state, *n_rounds);
  hyperkzg_open_t(srs_xy, n, poly, point, ell, t, com_xy, com_inf, w_xy, w_inf, v);
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}
static void hyperkzg_open_t(const uint64_t* srs_xy, size_t n, const uint64_t* poly, const uint64_t* point, size_t ell,
                            Transcript& t, uint64_t* com_xy, int32_t* com_inf, uint64_t* w_xy, int32_t* w_inf, uint64_t* v) {
  std::vector<G1Affine> srs = load_bases(srs_xy, n);
  FrVec p = load_fr(poly, n);
  std::vector<Fr> pt = load_fr(point, ell);
  HyperKZGProof pf = hyperkzg_open(srs, p, pt, t);
  for (size_t i = 0; i < pf.com.size(); i++) store_pt(com_xy + 8 * i, com_inf + i, pf.com[i]);
  for (size_t i = 0; i < 3; i++) store_pt(w_xy + 8 * i, w_inf + i, pf.w[i]);
  for (size_t i = 0; i < 3; i++) for (size_t j = 0; j < ell; j++) store_fr(v + 4 * (i * ell + j), pf.v[i][j]);
}

// ---- CPU baseline timing legs (bench.py): same synthetic workloads as ja_bench_kernel, on all host threads ----
static FrVec pseudo(size_t n, uint32_t seed) {
  FrVec v(n);
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) {
    uint32_t x = (uint32_t)i * 2654435761u + seed; uint32_t l[8];
    for (int k = 0; k < 8; k++) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; l[k] = x; }
    l[7] &= 0x1fffffffu;
    memcpy(v[i].l, l, 32);
  }
  return v;
}
// which: 0 bind LowToHigh, 1 bind HighToLow, 2 MUL round eval, 3 DOT2 round eval, 4 ADD round eval.  Returns ms per pass.
double orc_bench_kernel(int which, int log_n, int iters) {
  const size_t n = size_t(1) << log_n;
  FrVec a = pseudo(n, 17), b = pseudo(n, 18);
  const uint64_t rr[4] = {0, 0, 0x0123456789abcdefull, 0x0fedcba987654321ull};
  const Fr r = Fr::from_raw(rr);
  std::vector<Fr> w(log_n);
  for (int i = 0; i < log_n; i++) { uint64_t c[4] = {0, 0, 0x9e3779b97f4a7c15ull * (i + 1), 0x0123456789abcdefull + i}; w[i] = Fr::from_raw(c); }
  GruenSplitEq eq(w.data(), (size_t)log_n, LOW_TO_HIGH);
  double best = 1e30;
  volatile uint64_t sink = 0;
  for (int it = 0; it < iters; it++) {
    FrVec z = a;
    auto t0 = std::chrono::steady_clock::now();
    if (which <= 1) { bind_poly(z, r, which); sink += z[0].l[0]; }
    else if (which == 2) { Fr q[2]; eq.fold<2>([&](size_t g, Fr* v) { Fr l0 = a[2 * g], r0 = b[2 * g]; v[0] = l0 * r0; v[1] = (a[2 * g + 1] - l0) * (b[2 * g + 1] - r0); }, q); sink += q[0].l[0]; }
    else if (which == 4) { Fr q[1]; eq.fold<1>([&](size_t g, Fr* v) { v[0] = a[2 * g] + b[2 * g]; }, q); sink += q[0].l[0]; }
    else { DotInstance d({a, b}, Fr::zero()); auto t1 = std::chrono::steady_clock::now(); UniPoly u = d.compute_message(0, Fr::zero()); sink += u.coeffs[0].l[0]; t0 = t1; }
    auto t1 = std::chrono::steady_clock::now();
    double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (ms < best) best = ms;
  }
  return best;
}

}  // extern "C"
