// Multi-GPU support: both halves of the path shard with one tiny exchange per step and NO collective inside a
// kernel - a partial result is <= 17 field elements per sumcheck round or one group element per MSM and GPU.
//   MSM        index-range split: GPU g multiplies coefficients [lo_g, hi_g) with g1_powers[lo_g, hi_g) (the SRS is resident
//              on every GPU: 1 GiB at GPT-2 scale); the partial POINTS are all-gathered and added on the host.
//   sumcheck   contiguous hypercube slices for LowToHigh binding: binds are local, a round's reduced sums are partial
//              field sums over the GPU's pairs [g_offset, g_offset + len/2) against the REPLICATED split-eq tables.
// The exchange itself is the caller's (torch.distributed all_gather over NCCL / gloo in jolt_atlas_b200/parallel.py, an MPI
// or NCCL call in a Rust host); this file holds the device-side slice entry points and the host-side combine functions,
// which need no GPU (the world_size-2 gloo tests exercise them on CPU).
#include "common.hpp"
#include "fq_host.hpp"
#include "sumcheck_host.hpp"
#include "transcript_host.hpp"

extern "C" {

// ---- host-only combine functions -------------------------------------------------------------------------------------
int32_t ja_g1_sum_affine(const uint64_t* xy, const int32_t* is_inf, size_t n, uint64_t out_xy[8], int32_t* out_inf) {
  JA_REQUIRE((xy || n == 0) && out_xy && out_inf, "ja_g1_sum_affine: null argument");
  host::G1XH acc;
  memset(&acc, 0, sizeof(acc));
  for (size_t i = 0; i < n; i++) {
    if (is_inf && is_inf[i]) continue;
    host::FqH x, y;
    memcpy(x.l, xy + 8 * i, 32); memcpy(y.l, xy + 8 * i + 4, 32);
    host::xyzz_madd(acc, x, y);
  }
  host::xyzz_batch_to_affine(&acc, 1, out_xy, out_inf);
  return JA_OK;
}

int32_t ja_fr_sum(const uint64_t* vals, size_t n_parts, size_t n_vals, uint64_t* out) {
  JA_REQUIRE((vals || n_parts == 0) && out, "ja_fr_sum: null argument");
  for (size_t k = 0; k < n_vals; k++) {
    FrH acc = host::FR_ZERO;
    for (size_t p = 0; p < n_parts; p++) acc = host::add(acc, host::from_limbs(vals + 4 * (p * n_vals + k)));
    memcpy(out + 4 * k, acc.l, 32);
  }
  return JA_OK;
}

// ---- the library's Blake2b transcript for callers that own the transcript state (blake2b.rs) --------------------------------
void ja_transcript_new(const char* label, uint8_t state[32], uint32_t* n_rounds) {
  host::Blake2bTranscript t(label);
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}
void ja_transcript_append_points(uint8_t state[32], uint32_t* n_rounds, const uint64_t* xy, const int32_t* is_inf, size_t n) {
  host::Blake2bTranscript t(state, *n_rounds);
  t.append_points(xy, is_inf, n);                                   // blake2b.rs:189-195
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}
void ja_transcript_append_scalars(uint8_t state[32], uint32_t* n_rounds, const uint64_t* fr, size_t n) {
  host::Blake2bTranscript t(state, *n_rounds);
  t.append_scalars(reinterpret_cast<const FrH*>(fr), n);            // blake2b.rs:158-164
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}
void ja_transcript_challenge_scalar(uint8_t state[32], uint32_t* n_rounds, uint64_t out[4]) {
  host::Blake2bTranscript t(state, *n_rounds);
  const FrH r = t.challenge_scalar();                               // blake2b.rs:204-215
  memcpy(out, r.l, 32);
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}
void ja_transcript_challenge_scalar_powers(uint8_t state[32], uint32_t* n_rounds, size_t n, uint64_t* out) {
  host::Blake2bTranscript t(state, *n_rounds);
  const FrH q = t.challenge_scalar();                               // blake2b.rs:224-231
  FrH p = host::FR_ONE;
  for (size_t i = 0; i < n; i++) { memcpy(out + 4 * i, p.l, 32); p = host::mul(p, q); }
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}

void ja_transcript_challenge_optimized(uint8_t state[32], uint32_t* n_rounds, size_t n, uint64_t* out) {
  host::Blake2bTranscript t(state, *n_rounds);
  for (size_t i = 0; i < n; i++) t.challenge_scalar_optimized(out + 4 * i);      // blake2b.rs:233-238: 125-bit challenges {0, 0, lo, hi}
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}
// ExpandingTable (joltworks/src/utils/expanding_table.rs:62-89) after `n` updates with the given challenges, starting from [1]:
// HighToLow: values[2i] = v[i] - r v[i], values[2i+1] = r v[i]  (the first challenge ends up in the top index bit);
// LowToHigh: values[i] = v[i] - r v[i], values[i + len] = r v[i].  out = 2^n Fr.  Host-only O(2^n) glue (n = 8 per ps_shout phase).
int32_t ja_expanding_table(const uint64_t* challenges, size_t n, int32_t order, uint64_t* out) {
  JA_REQUIRE((challenges || n == 0) && out && n <= 20, "ja_expanding_table: bad argument");
  std::vector<FrH> v{host::FR_ONE};
  for (size_t j = 0; j < n; j++) {
    const FrH r = host::from_limbs(challenges + 4 * j);
    std::vector<FrH> nv(v.size() * 2);
    for (size_t i = 0; i < v.size(); i++) {
      const FrH e1 = host::mul(r, v[i]), e0 = host::sub(v[i], e1);
      if (order == JA_HIGH_TO_LOW) { nv[2 * i] = e0; nv[2 * i + 1] = e1; } else { nv[i] = e0; nv[i + v.size()] = e1; }
    }
    v.swap(nv);
  }
  memcpy(out, v.data(), v.size() * 32);
  return JA_OK;
}

// ---- evaluation reduction (joltworks/src/subprotocols/evaluation_reduction.rs:91-148, :223-249) --------------------------
// h = mle o l, where l is the degree n-1 curve through the n opening points (l(i) = point i, one UniPoly::from_evals per
// variable, :213-221).  The reference folds the table with POLYNOMIAL-valued entries, serially; h is the unique polynomial
// of degree <= m (n-1) with h(t) = mle(l(t)), so here it is m (n-1) + 1 ordinary MLE evaluations on the device
// (ja_poly_evaluate: eq table + dot) followed by one interpolation on the host (cached matrix).  Coefficients are trimmed
// like UniPoly::from_coeff (the reference's Mul trims every intermediate product, unipoly.rs:463-476).
int32_t ja_eval_reduction_h(ja_ctx* c, const ja_poly* mle, const uint64_t* points, size_t n, size_t m, uint64_t* out_coeffs,
                            size_t* out_ncoeffs) {
  JA_REQUIRE(c && mle && points && out_coeffs && out_ncoeffs && n >= 1, "ja_eval_reduction_h: null argument");
  JA_REQUIRE((size_t(1) << m) == mle->len, "ja_eval_reduction_h: points must have log2(len) coordinates");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  const size_t D = m * (n - 1);
  // l_k(t): per-variable interpolation through (j, points[j][k])
  std::vector<host::Coeffs> var(m);
  for (size_t k = 0; k < m; k++) {
    std::vector<FrH> e(n);
    for (size_t j = 0; j < n; j++) e[j] = host::from_limbs(points + 4 * (j * m + k));
    var[k] = n == 1 ? host::Coeffs{e[0]} : host::apply_matrix(host::interp_matrix(n, false), e);
  }
  std::vector<FrH> evals(D + 1);
  std::vector<uint64_t> pt(4 * (m ? m : 1));
  for (size_t t = 0; t <= D; t++) {
    const FrH ft = host::from_u64(t);
    for (size_t k = 0; k < m; k++) { const FrH v = host::evaluate(var[k], ft); memcpy(pt.data() + 4 * k, v.l, 32); }
    uint64_t out[4];
    int32_t st = ja_poly_evaluate(c, mle, pt.data(), m, out);
    if (st) return st;
    evals[t] = host::from_limbs(out);
  }
  const host::Coeffs h = D == 0 ? host::Coeffs{evals[0]} : host::trim(host::apply_matrix(host::interp_matrix(D + 1, false), evals));
  for (size_t i = 0; i < h.size(); i++) memcpy(out_coeffs + 4 * i, h[i].l, 32);
  *out_ncoeffs = h.size();
  return JA_OK;
}

// ---- device-side slice entry points --------------------------------------------------------------------------------------
int32_t ja_set_cache_openings(ja_ctx* c, int32_t on) {
  JA_REQUIRE(c, "ja_set_cache_openings: null context");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  c->cache_openings = on != 0;
  return JA_OK;
}
// n x Transcript::append_scalar (no vector framing): what cache_openings does with the claims of one instance
void ja_transcript_append_scalar_each(uint8_t state[32], uint32_t* n_rounds, const uint64_t* fr, size_t n) {
  ja::host::Blake2bTranscript t(state, *n_rounds);
  for (size_t i = 0; i < n; i++) t.append_scalar(ja::host::from_limbs(fr + 4 * i));
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
}

int32_t ja_set_msm_shard(ja_ctx* c, uint32_t index, uint32_t count) {
  JA_REQUIRE(c && count >= 1 && index < count, "ja_set_msm_shard: bad shard");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  c->msm_shard_index = index; c->msm_shard_count = count;
  return JA_OK;
}

int32_t ja_msm_fr_range(ja_ctx* c, const ja_srs* srs, const ja_poly* scalars, size_t lo, size_t hi, uint64_t out_xy[8],
                        int32_t* is_inf) {
  JA_REQUIRE(c && srs && scalars && out_xy, "ja_msm_fr_range: null argument");
  JA_REQUIRE(lo <= hi && hi <= scalars->len, "ja_msm_fr_range: range outside the polynomial");
  if (hi > srs->n) return fail(JA_ERR_KEY_LENGTH, "KeyLengthError: SRS shorter than the polynomial");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  std::vector<MsmJob> jobs{MsmJob{scalars->data() + lo, hi - lo, 0 /* MSM_FR */, 254, lo}};
  const uint32_t si = c->msm_shard_index, sc = c->msm_shard_count;
  c->msm_shard_index = 0; c->msm_shard_count = 1;                   // the range is explicit here
  const int32_t st = ja_msm_run(c, srs, jobs, out_xy, is_inf);
  c->msm_shard_index = si; c->msm_shard_count = sc;
  return st;
}

}  // extern "C"
