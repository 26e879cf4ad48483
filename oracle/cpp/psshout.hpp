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

enum { SUF_ONE = 0, SUF_HIGHER_ALL_ZERO = 1, SUF_HZERO_MUL_LWORD = 2, SUF_HONE_MUL_LWORD = 3, SUF_IDENTITY = 4 };

// suffixes/higher_all_zero.rs:9-29, hzero_mul_lword.rs:9-36, hone_mul_lword.rs:10-37, one.rs, identity as a suffix polynomial
inline uint64_t suffix_mle(int kind, uint64_t bits_u64, unsigned len, unsigned XLEN, unsigned BOUND) {
  if (kind == SUF_ONE) return 1;
  if (kind == SUF_IDENTITY) return bits_u64;
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

}  // namespace orc
