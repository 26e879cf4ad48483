// TEST INFRASTRUCTURE ONLY - CPU restatement of the T-sized passes of the prefix-suffix Shout read-raf prover
// (joltworks/src/subprotocols/ps_shout/mod.rs) and of the suffix MLEs they evaluate, written as the reference writes them
// (bit loops, not closed forms), so that the device kernels' closed forms are checked against an independent formulation.
// Parity unpinned against the Rust prover (no toolchain, no KATs); the suffix functions are cross-checked by their defining
// property in tests/test_oracle_psshout.py (combine(prefix indicator, suffix) == materialize_entry of the clamp table).
#pragma once
#include <cstdint>
#include <vector>
#include "field.hpp"

namespace orc {

// LookupBits::split (joltworks/src/utils/lookup_bits.rs:33-40)
inline void lb_split(uint64_t bits, unsigned len, unsigned suffix_len, uint64_t* prefix, uint64_t* suffix) {
  *suffix = suffix_len >= 64 ? bits : bits % (uint64_t(1) << suffix_len);
  *prefix = suffix_len >= 64 ? 0 : bits >> suffix_len;
  (void)len;
}

enum { SUF_ONE = 0, SUF_HIGHER_ALL_ZERO = 1, SUF_HZERO_MUL_LWORD = 2, SUF_HONE_MUL_LWORD = 3, SUF_IDENTITY = 4, SUF_SHIFT = 5 };

// suffixes/higher_all_zero.rs:9-29, hzero_mul_lword.rs:9-36, hone_mul_lword.rs:10-37, one.rs, identity as a suffix polynomial
inline uint64_t suffix_mle(int kind, uint64_t bits_u64, unsigned len, unsigned XLEN, unsigned BOUND) {
  if (kind == SUF_ONE) return 1;
  if (kind == SUF_IDENTITY) return bits_u64;
  if (kind == SUF_SHIFT) return uint64_t(1) << len;      // ShiftSuffixPolynomial (poly/identity_poly.rs:160-166)
  const unsigned bound_index = XLEN - BOUND - 1;
  const unsigned suffix_start_index = XLEN - len;
  uint64_t lower_word = 0;
  for (unsigned pos = 0; pos < len; pos++) {
    const unsigned global_index = suffix_start_index + pos;
    const uint64_t bit = (bits_u64 >> (len - 1 - pos)) & 1;
    if (kind == SUF_HIGHER_ALL_ZERO) { if (global_index <= bound_index && bit == 1) return 0; continue; }
    if (kind == SUF_HZERO_MUL_LWORD && global_index <= bound_index && bit == 1) return 0;
    if (kind == SUF_HONE_MUL_LWORD && global_index <= bound_index && bit == 0) return 0;
    if (global_index > bound_index) { const unsigned exponent = XLEN - global_index - 1; lower_word += bit << exponent; }
  }
  return kind == SUF_HIGHER_ALL_ZERO ? 1 : lower_word;
}

struct PsShout {
  std::vector<uint64_t> idx;
  std::vector<Fr> u;            // u_evals
  unsigned log_k, phases, log_m;
  // init_phase (mod.rs:269-303) + init_suffix_polys (:305-335) / init_Q (prefix_suffix.rs:294-351): returns Q[s][y]
  std::vector<Fr> init_phase(unsigned phase, const Fr* v_prev, const uint32_t* kinds, size_t n_suf, unsigned bound) {
    const size_t m = size_t(1) << log_m, m_mask = m - 1;
    if (phase != 0) {
#pragma omp parallel for schedule(static)
      for (size_t j = 0; j < idx.size(); j++) {
        uint64_t prefix, suffix;
        lb_split(idx[j], log_k, (phases - phase) * log_m, &prefix, &suffix);
        u[j] *= v_prev[prefix & m_mask];
      }
    }
    const unsigned suffix_len = (phases - 1 - phase) * log_m;
    std::vector<Fr> Q(n_suf * m, Fr::zero());
#pragma omp parallel for schedule(static)
    for (size_t s = 0; s < n_suf; s++) {
      for (size_t j = 0; j < idx.size(); j++) {
        uint64_t prefix_bits, suffix_bits;
        lb_split(idx[j], log_k, suffix_len, &prefix_bits, &suffix_bits);
        const size_t y = prefix_bits & m_mask;
        const uint64_t t = suffix_mle((int)kinds[s], suffix_bits, suffix_len, log_k, bound);
        if (t != 0) Q[s * m + y] += u[j] * Fr::from_u64(t);
      }
    }
    return Q;
  }
  // init_log_t_rounds (mod.rs:420-446)
  std::vector<Fr> materialize_ra(const Fr* v) const {
    const size_t m = size_t(1) << log_m, m_mask = m - 1;
    std::vector<Fr> ra(idx.size());
#pragma omp parallel for schedule(static)
    for (size_t j = 0; j < idx.size(); j++) {
      Fr acc = Fr::one();
      for (unsigned phase = 0; phase < phases; phase++) {
        uint64_t prefix, suffix;
        lb_split(idx[j], log_k, (phases - 1 - phase) * log_m, &prefix, &suffix);
        acc *= v[phase * m + (prefix & m_mask)];
      }
      ra[j] = acc;
    }
    return ra;
  }
};


// ---- the LOG_K address rounds of the read-raf sumcheck (ps_shout/mod.rs:337-418, :491-560), written as the reference writes
// them: per round, per b, the four prefix MLEs of the clamp table at c = 0 and c = 2 from the streaming checkpoints
// (lookup_tables/prefixes/{higher_all_zero,higher_all_one,lower_word,msb}.rs), ClampSpec::combine (lookup_tables/clamp.rs:94-118),
// the raf part through the signed identity prefix-suffix decomposition (ps_shout/unary.rs:45-87, poly/prefix_suffix.rs:437-482,
// poly/signed_identity_poly.rs:155-217), H2L binding of the suffix polynomials, the expanding tables, the checkpoint updates.
enum { PFX_HAZ = 0, PFX_HAO = 1, PFX_LW = 2, PFX_MSB = 3 };

struct OptFr { bool has = false; Fr v; };

// SparseDensePrefix::prefix_mle; r_x == nullptr <=> None; b has b_len bits
inline Fr prefix_mle(int kind, const OptFr cp[4], const Fr* r_x, uint32_t c, uint64_t b, unsigned b_len, unsigned j, unsigned XLEN, unsigned BOUND) {
  const Fr cf = Fr::from_u64(c);
  if (kind == PFX_MSB) return j == 0 ? cf : (j == 1 ? *r_x : cp[PFX_MSB].v);                     // msb.rs:10-27
  const unsigned bound_index = XLEN - BOUND - 1;
  if (kind == PFX_HAZ || kind == PFX_HAO) {                                                    // higher_all_zero.rs:13-56 / higher_all_one.rs:13-56
    if (BOUND >= XLEN) return Fr::one();
    const bool zero = kind == PFX_HAZ;
    Fr result = cp[kind].has ? cp[kind].v : Fr::one();
    if (r_x) {
      if (j > 0 && j - 1 <= bound_index) result *= zero ? Fr::one() - *r_x : *r_x;
      if (j <= bound_index) result *= zero ? Fr::one() - cf : cf;
    } else if (j <= bound_index) {
      result *= zero ? Fr::one() - cf : cf;
    }
    for (unsigned pos = 0; pos < b_len; pos++) {
      const unsigned global_index = j + 1 + pos;
      if (global_index <= bound_index) {
        const uint64_t bit = (b >> (b_len - 1 - pos)) & 1;
        result *= zero ? Fr::one() - Fr::from_u64(bit) : Fr::from_u64(bit);
      }
    }
    return result;
  }
  // lower_word.rs:13-62
  if (j + b_len >= XLEN || BOUND >= XLEN) return Fr::zero();
  Fr result = cp[PFX_LW].has ? cp[PFX_LW].v : Fr::zero();
  if (r_x) {
    if (j > 0 && j - 1 > bound_index) result += Fr::from_u64(uint64_t(1) << (XLEN - (j - 1) - 1)) * *r_x;
    if (j > bound_index) result += Fr::from_u64(uint64_t(1) << (XLEN - j - 1)) * cf;
  } else if (j > bound_index) {
    result += Fr::from_u64(uint64_t(1) << (XLEN - j - 1)) * cf;
  }
  for (unsigned pos = 0; pos < b_len; pos++) {
    const unsigned global_index = j + 1 + pos;
    if (global_index > bound_index) {
      const uint64_t bit = (b >> (b_len - 1 - pos)) & 1;
      result += Fr::from_u64(uint64_t(1) << (XLEN - global_index - 1)) * Fr::from_u64(bit);
    }
  }
  return result;
}

// SparseDensePrefix::update_prefix_checkpoint
inline OptFr update_prefix_checkpoint(int kind, const OptFr cp[4], const Fr& r_x, const Fr& r_y, unsigned j, unsigned suffix_len, unsigned XLEN, unsigned BOUND) {
  OptFr o;
  if (kind == PFX_MSB) {                                                                       // msb.rs:29-45
    if (j == 0) return o;
    if (j == 1) { o.has = true; o.v = r_x; return o; }
    return cp[PFX_MSB];
  }
  const unsigned bound_index = XLEN - BOUND - 1;
  if (kind == PFX_HAZ || kind == PFX_HAO) {
    o.has = true;
    if (BOUND >= XLEN) { o.v = Fr::one(); return o; }
    const bool zero = kind == PFX_HAZ;
    Fr result = cp[kind].has ? cp[kind].v : Fr::one();
    if (j > 0 && j - 1 <= bound_index) result *= zero ? Fr::one() - r_x : r_x;
    if (j <= bound_index) result *= zero ? Fr::one() - r_y : r_y;
    o.v = result;
    return o;
  }
  if (j + suffix_len >= XLEN || BOUND >= XLEN) return o;                                       // lower_word.rs:64-94
  Fr result = cp[PFX_LW].has ? cp[PFX_LW].v : Fr::zero();
  if (j > 0 && j - 1 > bound_index) result += Fr::from_u64(uint64_t(1) << (XLEN - (j - 1) - 1)) * r_x;
  if (j > bound_index) result += Fr::from_u64(uint64_t(1) << (XLEN - j - 1)) * r_y;
  o.has = true; o.v = result;
  return o;
}

// ClampSpec::combine, SYMMETRIC = true (clamp.rs:94-118); prefixes [HAZ, HAO, LW, MSB], suffixes [HAZ, HZ*LW, HO*LW, One]
inline Fr clamp_combine(const Fr p[4], const Fr s[4], unsigned BOUND) {
  const Fr const_upper = Fr::from_u64((uint64_t(1) << BOUND) - 1);
  const Fr lower_coeff = const_upper + const_upper + Fr::one();
  return s[3] * const_upper - p[PFX_MSB] * s[3] * lower_coeff
       + p[PFX_HAZ] * (s[1] + p[PFX_LW] * s[3] - s[0] * const_upper)
       + p[PFX_HAO] * (s[2] + p[PFX_LW] * s[3]);
}

// ClampBoundedTable::evaluate_mle, SYMMETRIC = true (clamp.rs:140-192): r = XLEN coordinates, MSB first
inline Fr clamp_evaluate_mle(const Fr* r, unsigned XLEN, unsigned BOUND) {
  const unsigned ubound_index = XLEN - BOUND - 1;
  Fr haz = Fr::one(), hao = Fr::one(), lw = Fr::zero();
  for (unsigned i = 0; i <= ubound_index; i++) { haz *= Fr::one() - r[i]; hao *= r[i]; }
  for (unsigned i = ubound_index + 1; i < XLEN; i++) lw += r[i] * Fr::from_u64(uint64_t(1) << (XLEN - i - 1));
  const Fr const_upper = Fr::from_u64((uint64_t(1) << BOUND) - 1);
  const Fr lower_coeff = const_upper + const_upper + Fr::one();
  return const_upper - r[0] * lower_coeff + haz * (lw - const_upper) + hao * lw;
}
inline Fr two_pow(unsigned k) { Fr x = Fr::one(); for (unsigned i = 0; i < k; i++) x = x + x; return x; }
// SignedIdentityPoly::evaluate (signed_identity_poly.rs:43-59)
inline Fr signed_identity_evaluate(const Fr* r, unsigned n) {
  Fr y = Fr::zero();
  for (unsigned i = 0; i < n; i++) y += two_pow(i) * r[n - 1 - i];
  return y - r[0] * two_pow(n);
}

struct PsReadRaf {
  PsShout* ps = nullptr;
  unsigned XLEN = 64, BOUND = 31;
  Fr gamma;
  OptFr cp[4];                       // PrefixCheckpoints of the table's prefixes
  OptFr cp_id;                       // PrefixRegistry checkpoint of Prefix::SignedIdentity
  std::vector<Fr> Q[4], RQ[2], P_id; // suffix polynomials of the table, of the raf decomposition; the identity prefix polynomial
  std::vector<Fr> r;
  std::vector<std::vector<Fr>> v;    // expanding tables
  Fr val, raf_val;

  void init_phase(unsigned phase) {  // mod.rs:269-303
    const uint32_t kinds[6] = {SUF_HIGHER_ALL_ZERO, SUF_HZERO_MUL_LWORD, SUF_HONE_MUL_LWORD, SUF_ONE, SUF_ONE, SUF_IDENTITY};
    const size_t m = size_t(1) << ps->log_m;
    std::vector<Fr> all = ps->init_phase(phase, phase ? v[phase - 1].data() : nullptr, kinds, 6, BOUND);
    for (int s = 0; s < 4; s++) Q[s].assign(all.begin() + s * m, all.begin() + (s + 1) * m);
    for (int s = 0; s < 2; s++) RQ[s].assign(all.begin() + (4 + s) * m, all.begin() + (5 + s) * m);
    // SignedIdentityPoly::prefix_polynomial (signed_identity_poly.rs:183-217)
    const unsigned chunk_len = ps->log_m, suffix_len = XLEN - chunk_len * (phase + 1);
    const Fr bound_value = cp_id.has ? cp_id.v : Fr::zero();
    P_id.assign(m, Fr::zero());
    for (size_t i = 0; i < m; i++) {
      if (phase == 0) {
        const uint64_t sign_bit = (i >> (chunk_len - 1)) & 1;
        P_id[i] = bound_value + Fr::from_u64(uint64_t(i) << suffix_len);                        // (i << suffix_len) mod 2^xlen: the u64 shift
        if (sign_bit) P_id[i] -= two_pow(XLEN);
      } else {
        P_id[i] = bound_value + Fr::from_u64(uint64_t(i) << suffix_len);
      }
    }
    v[phase].assign(1, Fr::one());
  }
  // compute_prefix_suffix_prover_message (mod.rs:337-352): [eval at 0, eval at 2]
  void message(unsigned j, Fr out[2]) const {
    const Fr* r_x = (j % 2 == 1) ? &r.back() : nullptr;
    const size_t half = Q[0].size() / 2;
    unsigned b_len = 0; while ((size_t(1) << b_len) < half) b_len++;
    Fr e0 = Fr::zero(), e2l = Fr::zero(), e2h = Fr::zero();
    for (size_t i = 0; i < half; i++) {                                                        // prover_msg_read_checking (:354-418)
      Fr p0[4], p2[4], lo[4], hi[4];
      for (int k = 0; k < 4; k++) {
        p0[k] = prefix_mle(k, cp, r_x, 0, i, b_len, j, XLEN, BOUND);
        p2[k] = prefix_mle(k, cp, r_x, 2, i, b_len, j, XLEN, BOUND);
        lo[k] = Q[k][i]; hi[k] = Q[k][i + half];
      }
      e0 += clamp_combine(p0, lo, BOUND);
      e2l += clamp_combine(p2, lo, BOUND);
      e2h += clamp_combine(p2, hi, BOUND);
    }
    Fr raf0 = Fr::zero(), raf2 = Fr::zero();                                                   // UnaryRafPS::prover_msg (unary.rs:54-76)
    for (size_t b = 0; b < half; b++) {                                                        // PrefixSuffixDecomposition::sumcheck_evals (:437-482)
      const Fr pl = P_id[b], ph = P_id[b + half];
      const Fr pe0 = pl, pe2 = ph + ph - pl;                                                    // dense sumcheck_evals(index, 2, HighToLow) = [P(0), P(2)]
      const Fr a0 = pe0 * RQ[0][b] + RQ[1][b];
      const Fr a2l = pe2 * RQ[0][b] + RQ[1][b];
      const Fr a2r = pe2 * RQ[0][b + half] + RQ[1][b + half];
      raf0 += a0; raf2 += a2r + a2r - a2l;
    }
    out[0] = e0 + gamma * raf0;
    out[1] = e2h + e2h - e2l + gamma * raf2;
  }
  // s(0) + s(1) of round 0 from the prover's own tables: the claimed sum rv(r_cycle) + gamma * operand(r_cycle)
  Fr derived_input_claim() const {
    const size_t half = Q[0].size() / 2;
    unsigned b_len = 0; while ((size_t(1) << b_len) < half) b_len++;
    Fr acc = Fr::zero();
    for (uint32_t c = 0; c < 2; c++)
      for (size_t i = 0; i < half; i++) {
        Fr pc[4], sv[4];
        for (int k = 0; k < 4; k++) { pc[k] = prefix_mle(k, cp, nullptr, c, i, b_len, 0, XLEN, BOUND); sv[k] = Q[k][i + c * half]; }
        acc += clamp_combine(pc, sv, BOUND) + gamma * (P_id[i + c * half] * RQ[0][i + c * half] + RQ[1][i + c * half]);
      }
    return acc;
  }
  static void bind_h2l(std::vector<Fr>& z, const Fr& rj) {
    const size_t n = z.size() / 2;
    for (size_t i = 0; i < n; i++) z[i] = z[i] + rj * (z[i + n] - z[i]);
    z.resize(n);
  }
  // ingest_challenge for round < LOG_K (mod.rs:491-560)
  void ingest(const Fr& rj, unsigned round) {
    const unsigned log_m = ps->log_m, phase = round / log_m;
    r.push_back(rj);
    for (auto& q : Q) bind_h2l(q, rj);
    for (auto& q : RQ) bind_h2l(q, rj);
    bind_h2l(P_id, rj);
    {                                                                                          // ExpandingTable::update, HighToLow
      std::vector<Fr>& t = v[phase];
      std::vector<Fr> nv(t.size() * 2);
      for (size_t i = 0; i < t.size(); i++) { const Fr e1 = rj * t[i]; nv[2 * i] = t[i] - e1; nv[2 * i + 1] = e1; }
      t.swap(nv);
    }
    if (r.size() % 2 == 0) {
      const unsigned suffix_len = ps->log_k - (round / log_m + 1) * log_m;
      OptFr prev[4] = {cp[0], cp[1], cp[2], cp[3]};
      for (int k = 0; k < 4; k++) cp[k] = update_prefix_checkpoint(k, prev, r[r.size() - 2], r[r.size() - 1], round, suffix_len, XLEN, BOUND);
    }
    if ((round + 1) % log_m == 0) {
      cp_id.has = true; cp_id.v = P_id[0];                                                     // PrefixRegistry::update_checkpoints
      if (phase != ps->phases - 1) init_phase(phase + 1);
    }
    if (round + 1 == ps->log_k) {
      Fr p[4], s[4];
      for (int k = 0; k < 4; k++) p[k] = cp[k].v;
      const int kinds[4] = {SUF_HIGHER_ALL_ZERO, SUF_HZERO_MUL_LWORD, SUF_HONE_MUL_LWORD, SUF_ONE};
      for (int k = 0; k < 4; k++) s[k] = Fr::from_u64(suffix_mle(kinds[k], 0, 0, XLEN, BOUND));
      val = clamp_combine(p, s, BOUND);
      raf_val = gamma * cp_id.v;                                                               // UnaryRafPS::raf_val (unary.rs:82-86)
    }
  }
};


// IdentityRCProver (subprotocols/identity_range_check.rs:140-325), address rounds: PrefixSuffixDecomposition<F, 2, false> over
// IdentityPolynomial (poly/identity_poly.rs:113-166), per b as the reference sums it (poly/prefix_suffix.rs:437-482)
struct PsIdentityRC {
  PsShout* ps = nullptr;
  OptFr cp;                          // PrefixRegistry checkpoint of Prefix::Identity
  std::vector<Fr> Q[2], P;
  std::vector<std::vector<Fr>> v;
  Fr raf_val;
  void init_phase(unsigned phase) {  // identity_range_check.rs:188-207
    const uint32_t kinds[2] = {SUF_SHIFT, SUF_IDENTITY};
    const size_t m = size_t(1) << ps->log_m;
    std::vector<Fr> all = ps->init_phase(phase, phase ? v[phase - 1].data() : nullptr, kinds, 2, 0);
    for (int s = 0; s < 2; s++) Q[s].assign(all.begin() + s * m, all.begin() + (s + 1) * m);
    const Fr bound_value = cp.has ? cp.v : Fr::zero();
    P.assign(m, Fr::zero());
    for (size_t i = 0; i < m; i++) P[i] = bound_value * two_pow(ps->log_m) + Fr::from_u64(i);   // identity_poly.rs:143-147
    v[phase].assign(1, Fr::one());
  }
  void message(Fr out[2]) const {    // prover_msg (:215-229)
    const size_t half = Q[0].size() / 2;
    Fr e0 = Fr::zero(), e2 = Fr::zero();
    for (size_t b = 0; b < half; b++) {
      const Fr pe0 = P[b], pe2 = P[b + half] + P[b + half] - P[b];
      const Fr a0 = pe0 * Q[0][b] + Q[1][b];
      const Fr a2l = pe2 * Q[0][b] + Q[1][b];
      const Fr a2r = pe2 * Q[0][b + half] + Q[1][b + half];
      e0 += a0; e2 += a2r + a2r - a2l;
    }
    out[0] = e0; out[1] = e2;
  }
  Fr derived_input_claim() const {
    Fr acc = Fr::zero();
    for (size_t i = 0; i < Q[0].size(); i++) acc += P[i] * Q[0][i] + Q[1][i];
    return acc;
  }
  void ingest(const Fr& rj, unsigned round) {   // :286-311
    const unsigned log_m = ps->log_m, phase = round / log_m;
    PsReadRaf::bind_h2l(Q[0], rj); PsReadRaf::bind_h2l(Q[1], rj); PsReadRaf::bind_h2l(P, rj);
    std::vector<Fr>& t = v[phase];
    std::vector<Fr> nv(t.size() * 2);
    for (size_t i = 0; i < t.size(); i++) { const Fr e1 = rj * t[i]; nv[2 * i] = t[i] - e1; nv[2 * i + 1] = e1; }
    t.swap(nv);
    if ((round + 1) % log_m == 0) {
      cp.has = true; cp.v = P[0];
      if (phase != ps->phases - 1) init_phase(phase + 1);
    }
    if (round + 1 == ps->log_k) raf_val = cp.v;
  }
};

}  // namespace orc
