// Sumcheck::prove (joltworks/src/subprotocols/sumcheck.rs:565-599) with the per-round work on the device:
//   compute_message  -> ja_round_eval (one kernel, reduced sums back to the host) + O(degree) interpolation
//   transcript       -> the library's Blake2b transcript (blake2b.rs; UniPoly_begin / coeffs except linear / UniPoly_end,
//                       unipoly.rs:550-558), challenge_scalar_optimized
//   ingest_challenge -> ja_bind_many (one kernel for all participating MLEs) + GruenSplitEqPolynomial::bind
// Built only from the public C ABI above it; a Rust caller that owns the transcript drives the same three calls itself
// (INTEGRATION.md).  Instance kinds = the JA_EVAL_* bodies of include/jolt_atlas_b200.h.
#include "common.hpp"
#include "sumcheck_host.hpp"
#include "transcript_host.hpp"

#include <chrono>
#include <cstdlib>
using host::Coeffs;

// JA_SC_TRACE=1: per-phase host wall-clock of the round loop on stderr (tuning aid)
struct ScTrace {
  bool on = getenv("JA_SC_TRACE") != nullptr;
  double t[6] = {0, 0, 0, 0, 0, 0};
  std::chrono::steady_clock::time_point last;
  void start() { if (on) last = std::chrono::steady_clock::now(); }
  void lap(int k) { if (!on) return; auto n = std::chrono::steady_clock::now(); t[k] += std::chrono::duration<double, std::micro>(n - last).count(); last = n; }
};
static ScTrace g_trace;

namespace {

struct Instance {
  int32_t kind;
  std::vector<ja_poly*> polys;
  ja_spliteq* eq = nullptr;        // family S / PROD / POW
  std::vector<FrH> gammas;         // SUM1
  uint32_t pow_d = 0;
  int order = JA_LOW_TO_HIGH;
  size_t n_out = 0;
};

int32_t instance_message(ja_ctx* c, Instance& in, const FrH& prev, Coeffs* uni) {
  uint64_t ev[32 * 4];
  const uint64_t* aux = in.gammas.empty() ? nullptr : reinterpret_cast<const uint64_t*>(in.gammas.data());
  RoundEvalPending pend;
  int32_t st = ja_round_eval_launch(c, in.kind, in.polys.data(), in.polys.size(), in.eq, aux, in.gammas.size(), in.pow_d,
                                    in.n_out, &pend);
  if (st) return st;
  g_trace.lap(0);
  // while the kernel runs: the one field division of the round (it depends on the eq state only)
  FrH cs = host::FR_ONE, cw = host::FR_ZERO, div = host::FR_ZERO;
  if (in.eq) {
    uint64_t t[4];
    ja_spliteq_current_scalar(in.eq, t); cs = host::from_limbs(t);
    if ((st = ja_spliteq_current_w(in.eq, t))) return st;
    cw = host::from_limbs(t);
    if (in.kind == JA_EVAL_PROD || in.kind == JA_EVAL_POW) div = host::inv(host::sub(host::FR_ONE, cw));
    else div = host::inv(host::gruen_eq1(cs, cw));
  }
  g_trace.lap(1);
  if ((st = ja_round_eval_collect(c, pend, ev))) return st;
  g_trace.lap(2);
  std::vector<FrH> e(in.n_out);
  for (size_t k = 0; k < in.n_out; k++) e[k] = host::from_limbs(ev + 4 * k);
  switch (in.kind) {
    case JA_EVAL_ADD: case JA_EVAL_SUB: case JA_EVAL_IDENT:
      *uni = host::gruen_poly_deg_2(cs, cw, e[0], prev, div); break;               // ops/add.rs:297-304
    case JA_EVAL_MUL: case JA_EVAL_SQUARE:
      *uni = host::gruen_poly_deg_3(cs, cw, e[0], e[1], prev, div); break;         // ops/mul.rs:177
    case JA_EVAL_PROD: case JA_EVAL_POW:
      for (auto& x : e) x = host::mul(x, cs);                                      // mles_product_sum.rs:120-128
      *uni = host::finish_mles_product_sum_from_evals(e, prev, cw, div); break;
    default:
      *uni = host::from_evals_and_hint(prev, e); break;                            // einsum/dot.rs:304,349; hamming_weight.rs:137
  }
  return JA_OK;
}

}  // namespace

extern "C" {

int32_t ja_sumcheck_prove(ja_ctx* c, int32_t kind, ja_poly* const* polys, size_t n_polys, const uint64_t* eq_w, size_t eq_m,
                          const uint64_t* aux_fr, size_t n_aux, uint32_t aux_u32, const uint64_t claim[4],
                          uint8_t transcript_state[32], uint32_t* n_rounds_io, size_t max_coeffs, uint64_t* out_coeffs,
                          uint32_t* out_ncoeffs, uint64_t* out_challenges, uint64_t* out_final_claims) {
  JA_REQUIRE(c && polys && n_polys && claim && transcript_state && n_rounds_io && out_coeffs && out_ncoeffs && out_challenges,
             "ja_sumcheck_prove: null argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  Instance in;
  in.kind = kind;
  in.polys.assign(polys, polys + n_polys);
  in.pow_d = aux_u32;
  const size_t len = ja_poly_len(polys[0]);
  JA_REQUIRE(len >= 2 && (len & (len - 1)) == 0, "ja_sumcheck_prove: polynomial length must be a power of two >= 2");
  size_t rounds = 0;
  while ((size_t(1) << rounds) < len) rounds++;
  bool family_s = false;
  switch (kind) {
    case JA_EVAL_ADD: case JA_EVAL_SUB: case JA_EVAL_IDENT: in.n_out = 1; family_s = true; break;
    case JA_EVAL_MUL: case JA_EVAL_SQUARE: in.n_out = 2; family_s = true; break;
    case JA_EVAL_PROD: in.n_out = n_polys; family_s = true; break;
    case JA_EVAL_POW: in.n_out = aux_u32; family_s = true; break;
    case JA_EVAL_DOT2: in.n_out = 2; in.order = JA_HIGH_TO_LOW; break;
    case JA_EVAL_DOT3: in.n_out = 3; in.order = JA_HIGH_TO_LOW; break;
    case JA_EVAL_SUM1: in.n_out = 1; break;
    case JA_EVAL_SUMHI: in.n_out = 1; in.order = JA_HIGH_TO_LOW; break;
    default: return fail(JA_ERR_UNSUPPORTED, "ja_sumcheck_prove: kind not implemented");
  }
  if (kind == JA_EVAL_SUM1 && aux_fr) {
    JA_REQUIRE(n_aux == n_polys, "ja_sumcheck_prove: SUM1 takes one gamma per polynomial");
    in.gammas.resize(n_aux);
    for (size_t i = 0; i < n_aux; i++) in.gammas[i] = host::from_limbs(aux_fr + 4 * i);
  }
  int32_t st;
  if (family_s) {
    JA_REQUIRE(eq_w && eq_m == rounds, "ja_sumcheck_prove: family S needs one eq point coordinate per round");
    if ((st = ja_spliteq_new(c, eq_w, eq_m, JA_LOW_TO_HIGH, nullptr, &in.eq))) return st;
  }
  host::Blake2bTranscript t(transcript_state, *n_rounds_io);
  FrH prev = host::from_limbs(claim);
  t.append_scalar(prev);                                           // sumcheck.rs:574
  g_trace.start();
  for (size_t round = 0; round < rounds; round++) {
    Coeffs uni;
    if ((st = instance_message(c, in, prev, &uni))) { ja_spliteq_free(c, in.eq); return st; }
    g_trace.lap(3);
    const Coeffs cp = host::compress(uni);
    if (cp.size() > max_coeffs) { ja_spliteq_free(c, in.eq); return fail(JA_ERR_INVALID, "ja_sumcheck_prove: max_coeffs too small"); }
    t.append_message("UniPoly_begin");                             // unipoly.rs:550-558
    for (auto& x : cp) t.append_scalar(x);
    t.append_message("UniPoly_end");
    uint64_t ch[4];
    t.challenge_scalar_optimized(ch);                              // sumcheck.rs:586
    prev = host::evaluate(uni, host::from_limbs(ch));              // sumcheck.rs:589
    g_trace.lap(4);
    if (in.eq && (st = ja_spliteq_bind(c, in.eq, ch))) { ja_spliteq_free(c, in.eq); return st; }
    if ((st = ja_bind_many(c, in.polys.data(), in.polys.size(), ch, in.order))) { ja_spliteq_free(c, in.eq); return st; }
    g_trace.lap(5);
    out_ncoeffs[round] = (uint32_t)cp.size();
    for (size_t k = 0; k < cp.size(); k++) memcpy(out_coeffs + 4 * (round * max_coeffs + k), cp[k].l, 32);
    memcpy(out_challenges + 4 * round, ch, 32);
  }
  if (out_final_claims)
    for (size_t i = 0; i < n_polys; i++)
      if ((st = ja_final_claim(c, polys[i], out_final_claims + 4 * i))) { ja_spliteq_free(c, in.eq); return st; }
  ja_spliteq_free(c, in.eq);
  if (g_trace.on)
    fprintf(stderr, "[sc kind=%d n_polys=%zu rounds=%zu] cumulative us: launch=%.0f inv=%.0f wait=%.0f interp=%.0f hash+eval=%.0f bind=%.0f\n",
            kind, n_polys, rounds, g_trace.t[0], g_trace.t[1], g_trace.t[2], g_trace.t[3], g_trace.t[4], g_trace.t[5]);
  memcpy(transcript_state, t.state, 32);
  *n_rounds_io = t.n_rounds;
  return (int32_t)JA_OK;
}

}  // extern "C"
