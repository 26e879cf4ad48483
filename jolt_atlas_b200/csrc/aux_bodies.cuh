// Round-evaluation bodies whose weight comes from a SMALLER table indexed by the high bits of the pair index, and the
// LowToHigh / eq-scheduled plain products (SURVEY 8a addendum, family D).  These are the two-phase provers of the softmax
// stages, the selector / gather sums and the eq schedules of the einsum and mean-of-squares provers; the caller (which owns the
// phase logic and the transcript) drives them round by round through ja_round_eval + ja_bind_many.
//   JA_EVAL_WSUM        [sum_g p[2g] T[g >> s]]                                  softmax_last_axis/exp_sum.rs:146-158 (phase 1, degree 1)
//   JA_EVAL_WDOT2       [sum_g T[g >> s] X(k) e(k), k = 0, 2, 3]  LowToHigh      softmax_last_axis/max.rs:185-206 (phase 1)
//   JA_EVAL_DOT2_L2H    [sum_g a(0) b(0), sum_g a(2) b(2)]  LowToHigh            ops/slice.rs:254-272, reshape.rs:286, concat.rs:290,
//                                                                                gather/mod.rs:232-258 (b = table + gamma identity, linear)
//   JA_EVAL_SQ_EQHI     [sum_i l(k)^2 e(k), k = 0, 2, 3]  HighToLow, e from the eq table at i >> s (a constant once it is bound to
//                       one entry)                                              ops/mean_of_squares.rs:363-386
//   JA_EVAL_DOT2_EQHI   [sum_i l(k) r(k) e(k)]  the same with two operands        ops/einsum/dot.rs:306-326 (EqSchedule::High)
//   JA_EVAL_DOT2_EQLOW  [sum_i l(k) r(k) T[i & (2^s - 1)]]                        ops/einsum/dot.rs:328-347 (EqSchedule::Low, rounds < log_k)
// (JA_EVAL_WIDENT, the split-eq weighted [p[2g] T[g >> s]] of softmax_last_axis/recip_mult.rs:196-216, is body 12 of k_round_eval_s.)
// sumcheck_evals (multilinear_polynomial.rs:873-905): e(0) = a, e(k) = b + (k - 1)(b - a).
#pragma once
#include "poly_kernels.cuh"

namespace ja {

enum { W_SUM = 0, W_DOT2 = 1, W_DOT2_L2H = 2, W_SQ_EQHI = 3, W_DOT2_EQHI = 4, W_DOT2_EQLOW = 5 };
struct WArgs {
  const Fr* p[2];                // operands
  const Fr* tab;                 // the smaller table / eq polynomial
  unsigned long long tab_len;
  unsigned int shift;
};
template <int MODE> struct WOut { static constexpr int N = MODE == W_SUM ? 1 : (MODE == W_DOT2_L2H ? 2 : 3); };

template <int MODE>
__global__ void __launch_bounds__(kBlock)
k_round_eval_w(WArgs a, size_t half, Fr* partials, unsigned int* counter, Fr* out) {
  constexpr int NOUT = WOut<MODE>::N;
  Fr acc[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; k++) acc[k] = fp_zero<FrParams>();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    if (MODE == W_SUM) {
      acc[0] = fp_add<FrParams>(acc[0], fp_mul<FrParams>(fp_load(a.p[0] + 2 * i), fp_load(a.tab + (i >> a.shift))));
      continue;
    }
    constexpr bool L2H = MODE == W_DOT2 || MODE == W_DOT2_L2H;
    constexpr int NOP = MODE == W_SQ_EQHI ? 1 : 2;
    Fr prod[3];
#pragma unroll
    for (int q = 0; q < NOP; q++) {
      const Fr x0 = fp_load(a.p[q] + (L2H ? 2 * i : i)), x1 = fp_load(a.p[q] + (L2H ? 2 * i + 1 : i + half));
      const Fr m = fp_sub<FrParams>(x1, x0);
      const Fr e2 = fp_add<FrParams>(x1, m), e3 = fp_add<FrParams>(e2, m);
      if (q == 0) { prod[0] = x0; prod[1] = e2; prod[2] = e3; }
      else { prod[0] = fp_mul<FrParams>(prod[0], x0); prod[1] = fp_mul<FrParams>(prod[1], e2); if (NOUT == 3) prod[2] = fp_mul<FrParams>(prod[2], e3); }
    }
    if (MODE == W_SQ_EQHI) { prod[0] = fp_sqr<FrParams>(prod[0]); prod[1] = fp_sqr<FrParams>(prod[1]); prod[2] = fp_sqr<FrParams>(prod[2]); }
    if (MODE == W_DOT2 || MODE == W_DOT2_EQLOW) {
      const Fr t = fp_load(a.tab + (MODE == W_DOT2 ? (i >> a.shift) : (i & ((size_t(1) << a.shift) - 1))));
#pragma unroll
      for (int k = 0; k < 3; k++) prod[k] = fp_mul<FrParams>(prod[k], t);
    } else if (MODE == W_SQ_EQHI || MODE == W_DOT2_EQHI) {
      Fr w0, w2, w3;
      if (a.tab_len == 1) { w0 = fp_load(a.tab); w2 = w0; w3 = w0; }             // the eq polynomial is fully bound: its final claim
      else {
        const size_t j = i >> a.shift;
        const Fr t0 = fp_load(a.tab + j), t1 = fp_load(a.tab + j + a.tab_len / 2);
        const Fr m = fp_sub<FrParams>(t1, t0);
        w0 = t0; w2 = fp_add<FrParams>(t1, m); w3 = fp_add<FrParams>(w2, m);
      }
      prod[0] = fp_mul<FrParams>(prod[0], w0); prod[1] = fp_mul<FrParams>(prod[1], w2); prod[2] = fp_mul<FrParams>(prod[2], w3);
    }
#pragma unroll
    for (int k = 0; k < NOUT; k++) acc[k] = fp_add<FrParams>(acc[k], prod[k]);
  }
  grid_sum<NOUT>(acc, partials, counter, out);
}

}  // namespace ja
