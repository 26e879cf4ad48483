// Sumcheck-side kernels: variable binding, eq tables, split-eq round evaluation, dot-style rounds,
// i32 tensor folds.  Every kernel states the reference loop it replaces.
// All arithmetic is canonical BN254-Fr Montgomery (fp.cuh); sums are exact field sums, hence
// independent of reduction order / grid shape (bit-identical to the CPU prover).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "fp.cuh"

namespace ja {

#ifndef JA_COMMON_CONSTS
#define JA_COMMON_CONSTS
constexpr int kSMs = 148;             // B200
constexpr int kBlock = 256;
#endif

// ---- small helpers ---------------------------------------------------------------------------
// Montgomery form of a signed 32-bit integer with ONE product row: |v| * 2^288 * 2^-32 = |v| * R.
// (field/ark.rs:125-136 from_i32)
JA_DEV Fr fr_from_i32(int v) {
  const uint32_t c288[8] = {0x15b8b9dau, 0x93e78865u, 0xb05ea154u, 0x16df2426u,
                            0x302ab839u, 0x1271b743u, 0xec6c226eu, 0x06bc037eu};  // 2^288 mod p
  uint32_t mag = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
  Fr r;
  fp_mont_rows<FrParams, 1>(r.l, c288, &mag);
  fp_final_sub<FrParams>(r.l);
  return v < 0 ? fp_neg<FrParams>(r) : r;
}
// |v| up to 2^32 (difference of two i32): sign + 33-bit magnitude handled with two rows of 2^320.
JA_DEV Fr fr_from_i64_small(long long v) {
  const uint32_t c320[8] = {0x7c5fb586u, 0xb4c6edf9u, 0xbfeb93beu, 0x708c8d50u,
                            0x04f7e0efu, 0x9ffd1de4u, 0x9a392866u, 0x215b02acu};  // 2^320 mod p
  unsigned long long mag = v < 0 ? (unsigned long long)(-v) : (unsigned long long)v;
  uint32_t b[2] = {(uint32_t)mag, (uint32_t)(mag >> 32)};
  Fr r;
  fp_mont_rows<FrParams, 2>(r.l, c320, b);
  fp_final_sub<FrParams>(r.l);
  return v < 0 ? fp_neg<FrParams>(r) : r;
}

JA_DEV Fr fr_shfl_down(const Fr& a, int delta) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, a.l[i], delta);
  return r;
}
JA_DEV Fr fr_warp_sum(Fr a) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) a = fp_add<FrParams>(a, fr_shfl_down(a, d));
  return a;  // lane 0 holds the sum
}

// Publication of a field element into host-mapped memory WITHOUT a system-scope fence (the fence + flag protocol costs
// ~2 us per round: scripts/micro/latency_probe.cu).  Element k of a slot is three 16-byte vectors
// [l0 l1 l2 tag] [l3 l4 l5 tag] [l6 l7 chk tag]; an aligned 16-byte store reaches host memory as a unit, so every vector
// validates itself and the host waits until all vectors it expects carry the round's tag (common.hpp: wait_tagged).  chk = xor of
// the limbs and the tag: belt and braces - an element that fails it is simply read again.
JA_DEV void store_tagged(Fr* slot_base, int k, const Fr& v, unsigned int tag) {
  uint4* p = reinterpret_cast<uint4*>(slot_base) + 3 * k;
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.l[0]), "r"(v.l[1]), "r"(v.l[2]), "r"(tag) : "memory");
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p + 1), "r"(v.l[3]), "r"(v.l[4]), "r"(v.l[5]), "r"(tag) : "memory");
  const unsigned int chk = v.l[0] ^ v.l[1] ^ v.l[2] ^ v.l[3] ^ v.l[4] ^ v.l[5] ^ v.l[6] ^ v.l[7] ^ tag;   // the host re-reads an element whose checksum fails
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p + 2), "r"(v.l[6]), "r"(v.l[7]), "r"(chk), "r"(tag) : "memory");
}

// Block-wide exact field sum of NOUT values per thread, then grid-wide via per-block partials and a
// "last block" pass.  `partials` holds gridDim.x * NOUT Fr; `counter` must be 0 on entry and is reset.
// Returns true (block-uniform) in the one block that wrote `out`.  bx / nb = index of this block and number of blocks
// taking part (a kernel that hosts several independent reductions passes its own sub-grid).
// tag != 0: `out` is a host-mapped result slot and the sums are published with store_tagged.
template <int NOUT>
JA_DEV bool grid_sum_ex(Fr (&acc)[NOUT], Fr* partials, unsigned int* counter, Fr* out, unsigned int bx, unsigned int nb,
                        unsigned int tag = 0) {
  __shared__ Fr s_part[kBlock / 32][NOUT];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NOUT; k++) {
    Fr v = fr_warp_sum(acc[k]);
    if (lane == 0) s_part[warp][k] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
      Fr v = lane < (blockDim.x >> 5) ? s_part[lane][k] : fp_zero<FrParams>();
      v = fr_warp_sum(v);
      if (lane == 0) {
        if (nb != 1) fp_store(&partials[(size_t)bx * NOUT + k], v);
        else if (tag) store_tagged(out, k, v, tag);
        else fp_store(&out[k], v);
      }
    }
  }
  if (nb == 1) return true;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicInc(counter, nb - 1);   // wraps back to 0 after the last block
    s_last = (t == nb - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NOUT; k++) {
    Fr v = fp_zero<FrParams>();
    for (unsigned b = threadIdx.x; b < nb; b += blockDim.x) {
      const volatile uint32_t* p = reinterpret_cast<const volatile uint32_t*>(&partials[(size_t)b * NOUT + k]);
      Fr t;
#pragma unroll
      for (int i = 0; i < 8; i++) t.l[i] = p[i];
      v = fp_add<FrParams>(v, t);
    }
    v = fr_warp_sum(v);
    __syncthreads();
    if (lane == 0) s_part[warp][k] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
      Fr v = lane < (blockDim.x >> 5) ? s_part[lane][k] : fp_zero<FrParams>();
      v = fr_warp_sum(v);
      if (lane == 0) {
        if (tag) store_tagged(out, k, v, tag);
        else fp_store(&out[k], v);
      }
    }
  }
  return true;
}

template <int NOUT>
JA_DEV bool grid_sum(Fr (&acc)[NOUT], Fr* partials, unsigned int* counter, Fr* out) {
  return grid_sum_ex<NOUT>(acc, partials, counter, out, blockIdx.x, gridDim.x);
}

// ---- variable binding (the "fold") ------------------------------------------------------------
// LowToHigh: out[i] = a[2i] + r*(a[2i+1]-a[2i])      dense_mlpoly.rs:219-239
// HighToLow: out[i] = a[i] + r*(a[i+half]-a[i])      dense_mlpoly.rs:126-141
// Up to kMaxBindPolys polynomials of the same length per launch (one `ingest_challenge`).
constexpr int kMaxBindPolys = 32;
struct BindArgs {
  const Fr* in[kMaxBindPolys];
  Fr* out[kMaxBindPolys];
};

template <bool LOW_TO_HIGH>
__global__ void __launch_bounds__(kBlock) k_bind(BindArgs args, Challenge r, size_t half) {
  const Fr* __restrict__ in = args.in[blockIdx.y];
  Fr* __restrict__ out = args.out[blockIdx.y];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    Fr a, b;
    if (LOW_TO_HIGH) { a = fp_load(in + 2 * i); b = fp_load(in + 2 * i + 1); }
    else             { a = fp_load_rw(in + i);  b = fp_load_rw(in + i + half); }   // HighToLow binds run in place
    Fr m = fp_sub<FrParams>(b, a);
    fp_store(out + i, fp_add<FrParams>(a, fp_mul_challenge<FrParams>(m, r)));
  }
}

// First bind of a compact i32 polynomial (compact_polynomial.rs:272-353): 4 B/coeff in, 32 B/coeff out.
template <bool LOW_TO_HIGH>
__global__ void __launch_bounds__(kBlock) k_bind_i32(const int* __restrict__ in, Fr* __restrict__ out,
                                                    Challenge r, size_t half) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    int a, b;
    if (LOW_TO_HIGH) { int2 v = __ldg(reinterpret_cast<const int2*>(in) + i); a = v.x; b = v.y; }
    else             { a = __ldg(in + i); b = __ldg(in + i + half); }
    Fr fa = fr_from_i32(a);
    Fr m = fr_from_i64_small((long long)b - (long long)a);
    fp_store(out + i, fp_add<FrParams>(fa, fp_mul_challenge<FrParams>(m, r)));
  }
}

static __global__ void __launch_bounds__(kBlock) k_i32_to_fr(const int* __restrict__ in, Fr* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    fp_store(out + i, fr_from_i32(__ldg(in + i)));
}

// RaPolynomial materialisation (joltworks/src/poly/ra_poly.rs:31-81, shout.rs:549-598 compute_ra_evals): out[t] =
// table[idx[t]] for a small table (K <= 2^16 eq evaluations); idx == 0xFFFFFFFF (None) gives 0.
static __global__ void __launch_bounds__(kBlock)
k_gather_small_table(const Fr* __restrict__ table, const uint32_t* __restrict__ idx, size_t n, Fr* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const uint32_t k = __ldg(idx + t);
    fp_store(out + t, k == 0xffffffffu ? fp_zero<FrParams>() : fp_load(table + k));
  }
}

// ---- eq tables ---------------------------------------------------------------------------------
// All prefix tables of eq(w, .) in one buffer: level j (2^j entries) at offset 2^j - 1.
//   rev == 0: EqPolynomial::evals_cached      (eq_poly.rs:174-194)   level j+1: [2i+1] = prev[i]*w[j], [2i] = prev[i]-[2i+1]
//   rev == 1: EqPolynomial::evals_cached_rev  (eq_poly.rs:198-217)   level j+1: [i+2^j] = prev[i]*w[m-1-j], [i] = prev[i]-that
// One block per table (blockIdx.x selects the descriptor): the levels are tiny (<= 2^13 products).
struct EqLevelsArgs {
  const Fr* w[2];
  int m[2];
  int rev[2];
  Fr* buf[2];
  Fr scale[2];
};
static __global__ void __launch_bounds__(1024) k_eq_levels(EqLevelsArgs a) {
  const int t = blockIdx.x;
  const Fr* w = a.w[t];
  const int m = a.m[t], rev = a.rev[t];
  Fr* buf = a.buf[t];
  if (threadIdx.x == 0) fp_store(buf, a.scale[t]);
  __syncthreads();
  for (int j = 0; j < m; j++) {
    const Fr wj = rev ? w[m - 1 - j] : w[j];
    const Fr* prev = buf + ((size_t(1) << j) - 1);
    Fr* cur = buf + ((size_t(1) << (j + 1)) - 1);
    for (size_t i = threadIdx.x; i < (size_t(1) << j); i += blockDim.x) {
      Fr s = prev[i];
      Fr hi = fp_mul<FrParams>(s, wj);
      Fr lo = fp_sub<FrParams>(s, hi);
      if (rev) { fp_store(cur + i + (size_t(1) << j), hi); fp_store(cur + i, lo); }
      else     { fp_store(cur + 2 * i + 1, hi); fp_store(cur + 2 * i, lo); }
    }
    __syncthreads();
  }
}

// The same tables with the point passed BY VALUE in the kernel parameters (no staged H2D copy of w ahead of the launch: a
// copy is one more ~3 us operation on the stream, and every sumcheck instance starts with one of these) and the previous
// level read from shared memory instead of through L2 (levels up to 2^9 entries; larger ones fall back to global).
constexpr int kEqValMax = 16;
struct EqLevelsValArgs {
  Fr w[2][kEqValMax];
  int m[2];
  int rev[2];
  Fr* buf[2];
  Fr scale[2];
};
static __global__ void __launch_bounds__(kBlock) k_eq_levels_val(const __grid_constant__ EqLevelsValArgs a) {
  __shared__ Fr sm[2][512];
  const int t = blockIdx.x;
  const int m = a.m[t], rev = a.rev[t];
  Fr* buf = a.buf[t];
  if (threadIdx.x == 0) { fp_store(buf, a.scale[t]); sm[0][0] = a.scale[t]; }
  __syncthreads();
  for (int j = 0; j < m; j++) {
    const Fr wj = rev ? a.w[t][m - 1 - j] : a.w[t][j];
    const size_t n = size_t(1) << j;
    const Fr* prev = buf + (n - 1);
    Fr* cur = buf + (2 * n - 1);
    const Fr* sp = sm[j & 1];
    Fr* sc = sm[(j + 1) & 1];
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
      const Fr s = n <= 512 ? sp[i] : prev[i];
      const Fr hi = fp_mul<FrParams>(s, wj);
      const Fr lo = fp_sub<FrParams>(s, hi);
      const size_t ih = rev ? i + n : 2 * i + 1, il = rev ? i : 2 * i;
      fp_store(cur + ih, hi); fp_store(cur + il, lo);
      if (2 * n <= 512) { sc[ih] = hi; sc[il] = lo; }
    }
    __syncthreads();
  }
}

// out[x] = hi[x >> bits_lo] * lo[x & mask]   — EqPolynomial::evals (eq_poly.rs:77-101) as an outer product
static __global__ void __launch_bounds__(kBlock) k_eq_expand(const Fr* __restrict__ hi, const Fr* __restrict__ lo,
                                                     int bits_lo, size_t n, Fr* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t mask = (size_t(1) << bits_lo) - 1;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride)
    fp_store(out + x, fp_mul<FrParams>(fp_load(hi + (x >> bits_lo)), fp_load(lo + (x & mask))));
}

// ---- family S: split-eq weighted round evaluation, LowToHigh ------------------------------------
// sum_{x_out} E_out[x_out] * ( sum_{x_in} E_in[x_in] * f(g) ),  g = (x_out << bits_in) | x_in
// (GruenSplitEqPolynomial::par_fold_out_in_unreduced, split_eq_poly.rs:526-597).
// Blocks own contiguous runs of g so a thread keeps one x_out for as long as possible and applies
// E_out once per run (the reference's delayed outer product).
struct EvalPolys {
  const Fr* p[6];
  const Fr* aux;                 // scalars of the body (device memory): RSQRT {gamma, S^3}, LIN3 {tau}
  unsigned int shift;            // WIDENT: the table p[1] is indexed by g >> shift
};

// Bodies with three to five operands (SURVEY 8a addendum, family S), written once on the pair values (x0 = value at 2g,
// dx = x[2g+1] - x[2g]) and shared by the un-fused kernel below and the fused / round-resident kernels (fused_kernels.cuh):
//   8  IFF    [m0 a0 + (1 - m0) b0,  dm da - dm db]                         ops/iff.rs:189-216          (mask, a, b)
//   9  DIV    [r0 q0 + R0 - l0,  dr dq]                                     ops/div.rs:329-347          (l, r, q, R)
//   10 RSQRT  [x0 quot0 + dr0 - S^3 + gamma (out0^2 + sr0 - quot0),  dx dquot + gamma dout^2]   ops/rsqrt.rs:390-418
//                                                                         (x, quotient, output, div_remainder, sqrt_remainder; aux = gamma, S^3)
//   11 LIN3   [tau q0 + r0 - in0]                                           neural_teleport/division.rs:231-246   (input, quotient, remainder; aux = tau)
// (ScalarConstDiv's [l0 - R0], ops/scalar_const_div.rs:227-239, is body 1 = SUB.)
template <int KID> struct SGen;
template <> struct SGen<8> {
  static constexpr int NP = 3, NOUT = 2;
  JA_DEV static void eval(const Fr* lo, const Fr* hi, const Fr*, Fr* v) {
    const Fr dm = fp_sub<FrParams>(hi[0], lo[0]), da = fp_sub<FrParams>(hi[1], lo[1]), db = fp_sub<FrParams>(hi[2], lo[2]);
    v[0] = fp_add<FrParams>(lo[2], fp_mul<FrParams>(lo[0], fp_sub<FrParams>(lo[1], lo[2])));     // m a + (1 - m) b = b + m (a - b)
    v[1] = fp_mul<FrParams>(dm, fp_sub<FrParams>(da, db));
  }
};
template <> struct SGen<9> {
  static constexpr int NP = 4, NOUT = 2;
  JA_DEV static void eval(const Fr* lo, const Fr* hi, const Fr*, Fr* v) {
    v[0] = fp_sub<FrParams>(fp_add<FrParams>(fp_mul<FrParams>(lo[1], lo[2]), lo[3]), lo[0]);
    v[1] = fp_mul<FrParams>(fp_sub<FrParams>(hi[1], lo[1]), fp_sub<FrParams>(hi[2], lo[2]));
  }
};
template <> struct SGen<10> {
  static constexpr int NP = 5, NOUT = 2;
  JA_DEV static void eval(const Fr* lo, const Fr* hi, const Fr* aux, Fr* v) {
    const Fr gamma = fp_load(aux), s3 = fp_load(aux + 1);
    const Fr div0 = fp_sub<FrParams>(fp_add<FrParams>(fp_mul<FrParams>(lo[0], lo[1]), lo[3]), s3);
    const Fr sqrt0 = fp_sub<FrParams>(fp_add<FrParams>(fp_sqr<FrParams>(lo[2]), lo[4]), lo[1]);
    const Fr dout = fp_sub<FrParams>(hi[2], lo[2]);
    v[0] = fp_add<FrParams>(div0, fp_mul<FrParams>(gamma, sqrt0));
    v[1] = fp_add<FrParams>(fp_mul<FrParams>(fp_sub<FrParams>(hi[0], lo[0]), fp_sub<FrParams>(hi[1], lo[1])), fp_mul<FrParams>(gamma, fp_sqr<FrParams>(dout)));
  }
};
template <> struct SGen<11> {
  static constexpr int NP = 3, NOUT = 1;
  JA_DEV static void eval(const Fr* lo, const Fr*, const Fr* aux, Fr* v) {
    v[0] = fp_sub<FrParams>(fp_add<FrParams>(fp_mul<FrParams>(fp_load(aux), lo[1]), lo[2]), lo[0]);
  }
};

template <int KID> struct SBody;
template <> struct SBody<0> {  // ADD  ops/add.rs:283-296
  static constexpr int NOUT = 1;
  JA_DEV static void eval(const EvalPolys& P, size_t g, Fr (&v)[1]) {
    v[0] = fp_add<FrParams>(fp_load(P.p[0] + 2 * g), fp_load(P.p[1] + 2 * g)); }
};
template <> struct SBody<1> {  // SUB  ops/sub.rs:267
  static constexpr int NOUT = 1;
  JA_DEV static void eval(const EvalPolys& P, size_t g, Fr (&v)[1]) {
    v[0] = fp_sub<FrParams>(fp_load(P.p[0] + 2 * g), fp_load(P.p[1] + 2 * g)); }
};
template <> struct SBody<2> {  // MUL  ops/mul.rs:160-176
  static constexpr int NOUT = 2;
  JA_DEV static void eval(const EvalPolys& P, size_t g, Fr (&v)[2]) {
    Fr l0 = fp_load(P.p[0] + 2 * g), l1 = fp_load(P.p[0] + 2 * g + 1);
    Fr r0 = fp_load(P.p[1] + 2 * g), r1 = fp_load(P.p[1] + 2 * g + 1);
    v[0] = fp_mul<FrParams>(l0, r0);
    v[1] = fp_mul<FrParams>(fp_sub<FrParams>(l1, l0), fp_sub<FrParams>(r1, r0)); }
};
template <> struct SBody<3> {  // SQUARE  ops/square.rs:163
  static constexpr int NOUT = 2;
  JA_DEV static void eval(const EvalPolys& P, size_t g, Fr (&v)[2]) {
    Fr o0 = fp_load(P.p[0] + 2 * g), o1 = fp_load(P.p[0] + 2 * g + 1);
    Fr d = fp_sub<FrParams>(o1, o0);
    v[0] = fp_sqr<FrParams>(o0);
    v[1] = fp_sqr<FrParams>(d); }
};
template <> struct SBody<6> {  // IDENT  ps_shout/mod.rs:464-488, opening_reduction.rs:355-403
  static constexpr int NOUT = 1;
  JA_DEV static void eval(const EvalPolys& P, size_t g, Fr (&v)[1]) { v[0] = fp_load(P.p[0] + 2 * g); }
};

template <int KID> struct SBodyGen {
  static constexpr int NOUT = SGen<KID>::NOUT;
  JA_DEV static void eval(const EvalPolys& P, size_t g, Fr (&v)[NOUT]) {
    Fr lo[SGen<KID>::NP], hi[SGen<KID>::NP];
#pragma unroll
    for (int q = 0; q < SGen<KID>::NP; q++) { lo[q] = fp_load(P.p[q] + 2 * g); hi[q] = fp_load(P.p[q] + 2 * g + 1); }
    SGen<KID>::eval(lo, hi, P.aux, v);
  }
};
template <> struct SBody<12> {  // WIDENT  softmax_last_axis/recip_mult.rs:196-216 (phase 1): [exp_q(k, 0) * inv_sum(k)], k = g >> shift
  static constexpr int NOUT = 1;
  JA_DEV static void eval(const EvalPolys& P, size_t g, Fr (&v)[1]) { v[0] = fp_mul<FrParams>(fp_load(P.p[0] + 2 * g), fp_load(P.p[1] + (g >> P.shift))); }
};
template <> struct SBody<8> : SBodyGen<8> {};
template <> struct SBody<9> : SBodyGen<9> {};
template <> struct SBody<10> : SBodyGen<10> {};
template <> struct SBody<11> : SBodyGen<11> {};

template <int KID>
__global__ void __launch_bounds__(kBlock)
k_round_eval_s(EvalPolys P, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in, int bits_in, size_t G,
               size_t tiles_per_block, Fr* partials, unsigned int* counter, Fr* out, size_t g_off = 0) {
  // g_off: global index of this GPU's first pair when the arrays are one hypercube slice of a sharded MLE (shard.cu);
  // the polynomials are indexed locally, the replicated eq tables globally
  constexpr int NOUT = SBody<KID>::NOUT;
  Fr outer[NOUT], inner[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; k++) { outer[k] = fp_zero<FrParams>(); inner[k] = fp_zero<FrParams>(); }
  const size_t mask_in = (size_t(1) << bits_in) - 1;
  const size_t g_begin = (size_t)blockIdx.x * tiles_per_block * kBlock;
  size_t g_end = g_begin + tiles_per_block * kBlock;
  if (g_end > G) g_end = G;
  size_t cur_xout = ~size_t(0);
  for (size_t g = g_begin + threadIdx.x; g < g_end; g += kBlock) {
    const size_t x_out = (g + g_off) >> bits_in;
    if (x_out != cur_xout) {
      if (cur_xout != ~size_t(0)) {
        Fr eo = fp_load(e_out + cur_xout);
#pragma unroll
        for (int k = 0; k < NOUT; k++) {
          outer[k] = fp_add<FrParams>(outer[k], fp_mul<FrParams>(eo, inner[k]));
          inner[k] = fp_zero<FrParams>();
        }
      }
      cur_xout = x_out;
    }
    Fr v[NOUT];
    SBody<KID>::eval(P, g, v);
    Fr ei = fp_load(e_in + ((g + g_off) & mask_in));
#pragma unroll
    for (int k = 0; k < NOUT; k++) inner[k] = fp_add<FrParams>(inner[k], fp_mul<FrParams>(ei, v[k]));
  }
  if (cur_xout != ~size_t(0)) {
    Fr eo = fp_load(e_out + cur_xout);
#pragma unroll
    for (int k = 0; k < NOUT; k++) outer[k] = fp_add<FrParams>(outer[k], fp_mul<FrParams>(eo, inner[k]));
  }
  grid_sum<NOUT>(outer, partials, counter, out);
}

// ---- family S, product of d linear factors (RA virtualisation; Cube as a same-MLE power) -----------------------
// compute_mles_product_sum (mles_product_sum.rs:15-129): per pair g evaluate prod_i (p_i0 + X*dp_i) on the grid
// X in {1, ..., d-1, inf}, weight by the split eq tables, sum -> d field elements (current_scalar and the Toom
// interpolation stay with the caller, :330-376).  blockIdx.y selects a chunk of KC consecutive grid points so the
// per-thread state is KC products + 2*KC accumulators regardless of d; the factors are re-read per chunk (L2 hits).
// SAME == true: all d factors are polys[0] (compute_mle_product_sum, :41-55,:135-205, Cube x^3).
constexpr int kMaxProdPolys = 32;
struct ProdPolys { const Fr* p[kMaxProdPolys]; };
constexpr int kProdChunk = 4;

template <bool SAME>
__global__ void __launch_bounds__(kBlock)
k_round_eval_prod(ProdPolys P, int d, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in, int bits_in, size_t G,
                  size_t tiles_per_block, Fr* partials /* [gridDim.y][gridDim.x][KC] */, Fr* block_out /* [gridDim.y][KC] */,
                  unsigned int* counters /* gridDim.y */) {
  constexpr int KC = kProdChunk;
  const int k0 = blockIdx.y * KC;                 // grid point index of this chunk's first output
  Fr outer[KC], inner[KC];
#pragma unroll
  for (int k = 0; k < KC; k++) { outer[k] = fp_zero<FrParams>(); inner[k] = fp_zero<FrParams>(); }
  const size_t mask_in = (size_t(1) << bits_in) - 1;
  const size_t g_begin = (size_t)blockIdx.x * tiles_per_block * kBlock;
  size_t g_end = g_begin + tiles_per_block * kBlock;
  if (g_end > G) g_end = G;
  size_t cur_xout = ~size_t(0);
  // Montgomery form of the small integer k0 + 1 (first grid point X of the chunk)
  const Fr x0 = fr_from_i32(k0 + 1);
  for (size_t g = g_begin + threadIdx.x; g < g_end; g += kBlock) {
    const size_t x_out = g >> bits_in;
    if (x_out != cur_xout) {
      if (cur_xout != ~size_t(0)) {
        const Fr eo = fp_load(e_out + cur_xout);
#pragma unroll
        for (int k = 0; k < KC; k++) {
          outer[k] = fp_add<FrParams>(outer[k], fp_mul<FrParams>(eo, inner[k]));
          inner[k] = fp_zero<FrParams>();
        }
      }
      cur_xout = x_out;
    }
    Fr v[KC];
    for (int i = 0; i < d; i++) {
      const Fr* z = SAME ? P.p[0] : P.p[i];
      const Fr p0 = fp_load(z + 2 * g);
      const Fr dp = fp_sub<FrParams>(fp_load(z + 2 * g + 1), p0);
      Fr cur = fp_add<FrParams>(p0, fp_mul<FrParams>(dp, x0));     // factor at X = k0 + 1
#pragma unroll
      for (int k = 0; k < KC; k++) {
        const int idx = k0 + k;
        const Fr val = idx == d - 1 ? dp : cur;                     // last grid point is X = inf (leading coefficient)
        v[k] = i == 0 ? val : fp_mul<FrParams>(v[k], val);
        cur = fp_add<FrParams>(cur, dp);
      }
    }
    const Fr ei = fp_load(e_in + (g & mask_in));
#pragma unroll
    for (int k = 0; k < KC; k++) inner[k] = fp_add<FrParams>(inner[k], fp_mul<FrParams>(ei, v[k]));
  }
  if (cur_xout != ~size_t(0)) {
    const Fr eo = fp_load(e_out + cur_xout);
#pragma unroll
    for (int k = 0; k < KC; k++) outer[k] = fp_add<FrParams>(outer[k], fp_mul<FrParams>(eo, inner[k]));
  }
  grid_sum<KC>(outer, partials + (size_t)blockIdx.y * gridDim.x * KC, counters + blockIdx.y, block_out + (size_t)blockIdx.y * KC);
}

// ---- product of d <= 16 linear factors, warp-transposed --------------------------------------------------------------
// Same sums as k_round_eval_prod, laid out for latency: a group of L = next_pow2(d) lanes owns one pair g.  Lane i
// loads factor i (p_i[2g], p_i[2g+1]) and tabulates it on the L grid points with additions only
// (X = 1..L-1 at indices 0..L-2, X = inf at index L-1).  The d-fold product is then a log2(L)-level exchange in which
// each lane keeps half of its points and multiplies them with its partner's values at the same points
// (L-1 products per lane in total, L/2 + L/4 + ... + 1 on the critical path), so lane j ends with the product at grid
// point j.  The caller reads points 0..d-2 and L-1 (inf).  Lanes >= d carry the constant factor 1.
JA_DEV Fr fr_select(bool c, const Fr& a, const Fr& b) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = c ? a.l[i] : b.l[i];
  return r;
}
JA_DEV Fr fr_shfl_xor(const Fr& a, int mask) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(0xffffffffu, a.l[i], mask);
  return r;
}
template <int M> struct LaneProduct {     // M = number of values currently held per lane
  JA_DEV static void run(Fr (&v)[M], int lane) {
    constexpr int H = M / 2;
    const bool hi = (lane & H) != 0;
    Fr nv[H];
#pragma unroll
    for (int j = 0; j < H; j++) {
      const Fr send = fr_select(hi, v[j], v[j + H]);
      const Fr keep = fr_select(hi, v[j + H], v[j]);
      nv[j] = fp_mul<FrParams>(keep, fr_shfl_xor(send, H));
    }
    LaneProduct<H>::run(nv, lane);
    v[0] = nv[0];
  }
};
template <> struct LaneProduct<1> { JA_DEV static void run(Fr (&)[1], int) {} };

// Value of the product of the L lanes' linear factors p0 + X dp at this lane's evaluation point (points X = 1 .. L-1 and the
// leading coefficient "at infinity" in slot L-1; a pad lane is the constant 1).  Generic form: L values per lane, log2 L
// butterfly stages, L - 1 products per lane.  L = 16: the FIRST stage multiplies two LINEAR factors, whose product is the
// quadratic c0 + c1 X + c2 X^2 - three products (c0 = p0 q0, c2 = dp dq, q(1) = (p0 + dp)(q0 + dq)) and a second-difference
// recurrence (two additions per point) instead of eight products, and two exchanged field elements instead of eight:
// 10 products per lane instead of 15 (mles_product_sum.rs:61-129 computes the same values; the field is exact, the order free).
template <int L>
JA_DEV Fr lane_product(const Fr& p0, const Fr& dp, bool pad, int d, int li) {
  Fr v[L];
  Fr cur = p0;
#pragma unroll
  for (int k = 0; k < L - 1; k++) { cur = fp_add<FrParams>(cur, dp); v[k] = cur; }
  v[L - 1] = pad ? p0 : dp;
  LaneProduct<L>::run(v, li);
  return v[0];
}
template <>
JA_DEV Fr lane_product<16>(const Fr& p0, const Fr& dp, bool pad, int d, int li) {
  const Fr q0 = fr_shfl_xor(p0, 8), dq = fr_shfl_xor(dp, 8);
  const bool partner_pad = (li ^ 8) >= d;
  const Fr c0 = fp_mul<FrParams>(p0, q0), c2 = fp_mul<FrParams>(dp, dq);
  Fr q = fp_mul<FrParams>(fp_add<FrParams>(p0, dp), fp_add<FrParams>(q0, dq));                  // q(1) = c0 + c1 + c2
  const Fr two_c2 = fp_add<FrParams>(c2, c2);
  Fr dlt = fp_add<FrParams>(fp_sub<FrParams>(q, c0), two_c2);                                  // q(2) - q(1) = c1 + 3 c2
  // slot 15 of the pair: product of the two leading coefficients, a pad factor counting as 1
  const Fr inf = pad ? (partner_pad ? p0 : dq) : (partner_pad ? dp : c2);
  const bool hi = (li & 8) != 0;
  Fr nv[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { nv[j] = q; q = fp_add<FrParams>(q, dlt); dlt = fp_add<FrParams>(dlt, two_c2); }      // X = 1 .. 8 (low lanes)
#pragma unroll
  for (int j = 0; j < 7; j++) {                                                                                       // X = 9 .. 15 (high lanes)
    nv[j] = fr_select(hi, q, nv[j]);
    if (j < 6) { q = fp_add<FrParams>(q, dlt); dlt = fp_add<FrParams>(dlt, two_c2); }
  }
  nv[7] = fr_select(hi, inf, nv[7]);
  LaneProduct<8>::run(nv, li);
  return nv[0];
}

// sum of `v` over all threads of the block that share (threadIdx.x % L); valid in threads t < L afterwards
template <int L, int BLOCK = kBlock>
JA_DEV Fr block_sum_by_lane(Fr v) {
  __shared__ Fr s_lane[BLOCK / 32][L > 32 ? 32 : L];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int dlt = 16; dlt >= L; dlt >>= 1) v = fp_add<FrParams>(v, fr_shfl_down(v, dlt));
  __syncthreads();
  if (lane < L) s_lane[warp][lane] = v;
  __syncthreads();
  Fr t = fp_zero<FrParams>();
  if (threadIdx.x < L)
    for (int w = 0; w < BLOCK / 32; w++) t = fp_add<FrParams>(t, s_lane[w][threadIdx.x]);
  return t;
}

template <int L, bool SAME>
__global__ void __launch_bounds__(kBlock)
k_round_eval_prod_t(ProdPolys P, int d, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in, int bits_in, size_t G,
                    size_t pairs_per_block, Fr* partials /* [gridDim.x][L] */, Fr* out /* [L] */, unsigned int* counter,
                    size_t g_off = 0) {
  constexpr int GPB = kBlock / L;                     // lane groups (pairs in flight) per block
  const int li = threadIdx.x & (L - 1);
  const int group = threadIdx.x / L;
  const bool pad = li >= d;
  const Fr* __restrict__ z = P.p[(SAME || pad) ? 0 : li];
  const size_t mask_in = (size_t(1) << bits_in) - 1;
  const size_t g_begin = (size_t)blockIdx.x * pairs_per_block;
  size_t g_end = g_begin + pairs_per_block;
  if (g_end > G) g_end = G;
  Fr outer = fp_zero<FrParams>(), inner = fp_zero<FrParams>();
  size_t cur_xout = ~size_t(0);
  for (size_t base = g_begin; base < g_end; base += GPB) {      // uniform trip count: every lane joins the shuffles
    const size_t g = base + group;
    const bool active = g < g_end;
    const size_t gl = active ? g : g_begin;
    Fr p0, dp;
    if (pad) { p0 = fp_one<FrParams>(); dp = fp_zero<FrParams>(); }
    else { p0 = fp_load(z + 2 * gl); dp = fp_sub<FrParams>(fp_load(z + 2 * gl + 1), p0); }
    const Fr pv = lane_product<L>(p0, dp, pad, d, li);
    if (active) {
      const size_t x_out = (g + g_off) >> bits_in;
      if (x_out != cur_xout) {
        if (cur_xout != ~size_t(0)) {
          outer = fp_add<FrParams>(outer, fp_mul<FrParams>(fp_load(e_out + cur_xout), inner));
          inner = fp_zero<FrParams>();
        }
        cur_xout = x_out;
      }
      inner = fp_add<FrParams>(inner, fp_mul<FrParams>(fp_load(e_in + ((g + g_off) & mask_in)), pv));
    }
  }
  if (cur_xout != ~size_t(0)) outer = fp_add<FrParams>(outer, fp_mul<FrParams>(fp_load(e_out + cur_xout), inner));
  Fr tot = block_sum_by_lane<L>(outer);
  if (gridDim.x == 1) {
    if (threadIdx.x < L) fp_store(out + threadIdx.x, tot);
    return;
  }
  if (threadIdx.x < L) fp_store(partials + (size_t)blockIdx.x * L + threadIdx.x, tot);
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  Fr acc = fp_zero<FrParams>();
  for (unsigned b = threadIdx.x / L; b < gridDim.x; b += GPB) {
    const volatile uint32_t* q = reinterpret_cast<const volatile uint32_t*>(partials + (size_t)b * L + li);
    Fr t;
#pragma unroll
    for (int i = 0; i < 8; i++) t.l[i] = q[i];
    acc = fp_add<FrParams>(acc, t);
  }
  tot = block_sum_by_lane<L>(acc);
  if (threadIdx.x < L) fp_store(out + threadIdx.x, tot);
}

// ---- plain sums (no eq): Hamming weight (sum_j p_i[2j], LowToHigh; hamming_weight.rs:118-139) and Sum over an axis
// (sum_{j<n/2} p[j], HighToLow; ops/sum/axis.rs:220-233).  blockIdx.y = polynomial; out[i] = the sum of poly i.
// The gamma combination of the Hamming instance is O(d) host work on the returned sums.
struct SumPolys { const Fr* p[kMaxProdPolys]; };
template <int STRIDE>
__global__ void __launch_bounds__(kBlock)
k_round_sum(SumPolys P, size_t half, Fr* partials, Fr* out, unsigned int* counters) {
  const Fr* __restrict__ z = P.p[blockIdx.y];
  Fr acc[1];
  acc[0] = fp_zero<FrParams>();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += stride)
    acc[0] = fp_add<FrParams>(acc[0], fp_load(z + STRIDE * j));
  grid_sum<1>(acc, partials + (size_t)blockIdx.y * gridDim.x, counters + blockIdx.y, out + blockIdx.y);
}

// ---- family D: plain products at X in {0,2,3}, HighToLow ----------------------------------------
// sumcheck_evals (multilinear_polynomial.rs:873-905): e0 = a, e_k = b + (k-1)(b-a).
template <int NPOLY>   // 2: einsum/dot.rs:292-303 ; 3: dot.rs:330-350 with eq as a third MLE of equal length
__global__ void __launch_bounds__(kBlock)
k_round_eval_dot(EvalPolys P, size_t half, Fr* partials, unsigned int* counter, Fr* out) {
  constexpr int NOUT = NPOLY;
  Fr acc[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; k++) acc[k] = fp_zero<FrParams>();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    Fr prod[NOUT];
#pragma unroll
    for (int q = 0; q < NPOLY; q++) {
      Fr a = fp_load(P.p[q] + i), b = fp_load(P.p[q] + i + half);
      Fr m = fp_sub<FrParams>(b, a);
      Fr e = fp_add<FrParams>(b, m);   // X = 2
      if (q == 0) { prod[0] = a; prod[1] = e; } else { prod[0] = fp_mul<FrParams>(prod[0], a); prod[1] = fp_mul<FrParams>(prod[1], e); }
      if (NOUT == 3) {
        Fr e3 = fp_add<FrParams>(e, m);  // X = 3
        if (q == 0) prod[2] = e3; else prod[2] = fp_mul<FrParams>(prod[2], e3);
      }
    }
#pragma unroll
    for (int k = 0; k < NOUT; k++) acc[k] = fp_add<FrParams>(acc[k], prod[k]);
  }
  grid_sum<NOUT>(acc, partials, counter, out);
}

// plain sum_i a[i]*b[i] over n entries (MLE evaluation against a dense eq table)
static __global__ void __launch_bounds__(kBlock) k_dot_full(const Fr* __restrict__ a, const Fr* __restrict__ b, size_t n,
                                                    Fr* partials, unsigned int* counter, Fr* out) {
  Fr acc[1];
  acc[0] = fp_zero<FrParams>();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    acc[0] = fp_add<FrParams>(acc[0], fp_mul<FrParams>(fp_load(a + i), fp_load(b + i)));
  grid_sum<1>(acc, partials, counter, out);
}

// ---- i32 tensor folds (einsum operand fold, ops/einsum/mk_kn_mn.rs:47-79) ------------------------
// transpose == 0: out[j] = sum_i from_i32(A[i*cols+j]) * eq[i]; thread per column, rows split over grid.y,
//                 partial[y][j] summed by k_fold_cols_finish.
// Delayed reduction: from_i32(v) * e = v * e (e is a Montgomery residue, so v * e is the Montgomery form of the product) is
// accumulated as a plain 320-bit integer, positive and negative v apart - 8 IMAD.WIDE per element instead of a Montgomery
// row for from_i32 plus a full product (~150) - and folded back into the field ONCE per output:
//   X mod p = mont(1_mont, lo) + mont(R^2, hi)   for X = lo + hi 2^256
// (the unreduced value is the per-row multiplier of fp_mont_rows, whose bound "< 2p" needs only the OTHER operand < p).
// |v| <= 2^31 and < 2^32 terms keep X below 2^317.
struct Acc320 { uint32_t w[10]; };
JA_DEV Acc320 acc320_zero() { Acc320 a;
#pragma unroll
  for (int i = 0; i < 10; i++) a.w[i] = 0; return a; }
JA_DEV void acc320_mad(Acc320& a, const Fr& e, uint32_t v) {
  unsigned long long c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (unsigned long long)e.l[i] * v + a.w[i];       // <= (2^32-1)^2 + 2 (2^32-1) = 2^64 - 1
    a.w[i] = (uint32_t)c; c >>= 32;
  }
  c += a.w[8]; a.w[8] = (uint32_t)c; c >>= 32;
  a.w[9] += (uint32_t)c;
}
JA_DEV Fr acc320_reduce(const Acc320& a) {
  Fr lo, hi = fp_zero<FrParams>();
#pragma unroll
  for (int i = 0; i < 8; i++) lo.l[i] = a.w[i];
  hi.l[0] = a.w[8]; hi.l[1] = a.w[9];
  return fp_add<FrParams>(fp_mul<FrParams>(fp_one<FrParams>(), lo), fp_mul<FrParams>(fp_r2<FrParams>(), hi));
}
JA_DEV void acc320_mad_i32(Acc320& pos, Acc320& neg, const Fr& e, int v) {
  if (v > 0) acc320_mad(pos, e, (uint32_t)v);
  else if (v < 0) acc320_mad(neg, e, (uint32_t)(-(long long)v));
}

static __global__ void __launch_bounds__(kBlock)
k_fold_cols(const int* __restrict__ A, size_t rows, size_t cols, const Fr* __restrict__ eq,
            size_t rows_per_slice, Fr* __restrict__ partial) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  size_t i0 = (size_t)blockIdx.y * rows_per_slice, i1 = i0 + rows_per_slice;
  if (i1 > rows) i1 = rows;
  Acc320 pos = acc320_zero(), neg = acc320_zero();
  for (size_t i = i0; i < i1; i++) {
    const int v = __ldg(A + i * cols + j);
    if (v != 0) acc320_mad_i32(pos, neg, fp_load(eq + i), v);
  }
  fp_store(partial + (size_t)blockIdx.y * cols + j, fp_sub<FrParams>(acc320_reduce(pos), acc320_reduce(neg)));
}
static __global__ void __launch_bounds__(kBlock)
k_fold_cols_finish(const Fr* __restrict__ partial, size_t slices, size_t cols, Fr* __restrict__ out) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  Fr acc = fp_zero<FrParams>();
  for (size_t s = 0; s < slices; s++) acc = fp_add<FrParams>(acc, fp_load(partial + s * cols + j));
  fp_store(out + j, acc);
}
// transpose == 1: out[i] = sum_j from_i32(A[i*cols+j]) * eq[j]; one warp per row.
static __global__ void __launch_bounds__(kBlock)
k_fold_rows(const int* __restrict__ A, size_t rows, size_t cols, const Fr* __restrict__ eq, Fr* __restrict__ out) {
  const size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  Acc320 pos = acc320_zero(), neg = acc320_zero();
  for (size_t j = lane; j < cols; j += 32) {
    const int v = __ldg(A + row * cols + j);
    if (v != 0) acc320_mad_i32(pos, neg, fp_load(eq + j), v);
  }
  Fr acc = fp_sub<FrParams>(acc320_reduce(pos), acc320_reduce(neg));
  acc = fr_warp_sum(acc);
  if (lane == 0) fp_store(out + row, acc);
}

// ---- calibration: register-resident Montgomery products (roofline denominator for integer-bound kernels)
template <class M>
__global__ void __launch_bounds__(kBlock) k_calib_mul(Fp<M>* out, int iters) {
  // four independent product chains, every operand thread-variant (no uniform-datapath shortcut)
  Fp<M> a = fp_one<M>(), b = fp_r2<M>(), c = fp_one<M>(), d = fp_r2<M>();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  a.l[0] += t; b.l[1] ^= t * 2654435761u; c.l[2] += t * 40503u; d.l[3] ^= t;
  for (int i = 0; i < iters; i++) {
    a = fp_mul<M>(a, b);
    c = fp_mul<M>(c, d);
    b = fp_mul<M>(b, a);
    d = fp_mul<M>(d, c);
  }
  Fp<M> s = fp_add<M>(fp_add<M>(a, b), fp_add<M>(c, d));
  if (s.l[7] == 0xdeadbeefu) fp_store(out + threadIdx.x, s);   // never true for canonical values; keeps the loop alive
}

}  // namespace ja
