// HyperKZG::open on device (joltworks/src/poly/commitment/hyperkzg/mod.rs:400-447 and :192-280).
// The l-1 folds reuse k_bind<LowToHigh> (poly_kernels.cuh) and the commitments the batched MSM
// (msm_kernels.cuh); this file holds what is specific to the opening:
//   k_univariate_eval3   f(u), f(-u), f(u^2) in ONE pass over f             unipoly.rs:247-305
//   k_hkzg_lincomb       B = sum_k q^k * polys[k]                            dense_mlpoly.rs:444-499
//   k_witness_*          h = (f - f(u)) / (x - u) as a three-level suffix scan of the reference's
//                        serial recurrence h[i-1] = f[i] + h[i]*u             hyperkzg/mod.rs:213-229
// All sums are exact field sums, so the regrouping is bit-identical to the serial loops.
// `tab` arguments hold u^(2^j), j = 0..39, for each evaluation point (built on the host, 40 squarings).
#pragma once
#include "poly_kernels.cuh"

namespace ja {

constexpr int kPowTab = 40;

// x^e from the table of x^(2^j)
JA_DEV Fr pow_from_tab(const Fr* __restrict__ tab, unsigned long long e) {
  Fr r = fp_one<FrParams>();
  bool first = true;
  for (int j = 0; e; j++, e >>= 1) {
    if (e & 1) { Fr t = fp_load(tab + j); r = first ? t : fp_mul<FrParams>(r, t); first = false; }
  }
  return r;
}

// out[p] = sum_i f[i] * u_p^i, p = 0..2.  Thread g owns i = g, g+S, g+2S, ... (S = total threads, a power of two,
// log_s its log): Horner in u^S, then one multiplication by u^g.  Coalesced, each coefficient read once.
static __global__ void __launch_bounds__(kBlock)
k_univariate_eval3(const Fr* __restrict__ f, size_t n, const Fr* __restrict__ tab /*3 x kPowTab*/, int log_s,
                   Fr* partials, unsigned int* counter, Fr* out) {
  const size_t S = (size_t)gridDim.x * blockDim.x;
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fr acc[3];
#pragma unroll
  for (int p = 0; p < 3; p++) acc[p] = fp_zero<FrParams>();
  if (g < n) {
    Fr step[3];
#pragma unroll
    for (int p = 0; p < 3; p++) step[p] = fp_load(tab + p * kPowTab + log_s);
    const size_t last = g + ((n - 1 - g) / S) * S;
    for (size_t i = last;; i -= S) {
      const Fr c = fp_load(f + i);
#pragma unroll
      for (int p = 0; p < 3; p++) acc[p] = fp_add<FrParams>(fp_mul<FrParams>(acc[p], step[p]), c);
      if (i == g) break;
    }
    if (g) {
#pragma unroll
      for (int p = 0; p < 3; p++) acc[p] = fp_mul<FrParams>(acc[p], pow_from_tab(tab + p * kPowTab, g));
    }
  }
  grid_sum<3>(acc, partials, counter, out);
}

// B[j] = sum_{k : len_k > j} q[k] * P_k[j]; P = [poly_0 (n) | poly_1 (n/2) | ... | poly_{ell-1} (2)]
static __global__ void __launch_bounds__(kBlock)
k_hkzg_lincomb(const Fr* __restrict__ P, size_t n, int ell, const Fr* __restrict__ q, Fr* __restrict__ B) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    Fr acc = fp_mul<FrParams>(fp_load(q), fp_load(P + j));
    size_t off = n, len = n >> 1;
    for (int k = 1; k < ell && j < len; k++) {
      acc = fp_add<FrParams>(acc, fp_mul<FrParams>(fp_load(q + k), fp_load(P + off + j)));
      off += len; len >>= 1;
    }
    fp_store(B + j, acc);
  }
}

// ---- witness polynomial -----------------------------------------------------------------------------------
// s(a) = sum_{j >= a} f[j] u^(j-a);  h[i] = s(i+1).  Levels: thread chunk of L = 2^log_l coefficients, block of
// 256 chunks, grid.  tab = u^(2^j).
constexpr int kWitBlock = 256;

// local Horner value of thread chunk [a, a+L): E = sum_j f[a+j] u^j
JA_DEV Fr wit_chunk_eval(const Fr* __restrict__ f, size_t a, size_t L, const Fr& u) {
  Fr acc = fp_load(f + a + L - 1);
  for (size_t j = L - 1; j-- > 0;) acc = fp_add<FrParams>(fp_mul<FrParams>(acc, u), fp_load(f + a + j));
  return acc;
}
// in-block suffix scan: on return s_e[t] = sum_{t' >= t} E_t' (u^L)^(t'-t)
JA_DEV void wit_block_scan(Fr* s_e, const Fr* __restrict__ tab, int log_l) {
  for (int d = 0; (1 << d) < kWitBlock; d++) {
    const Fr mult = fp_load(tab + log_l + d);
    Fr add = fp_zero<FrParams>();
    const int o = threadIdx.x + (1 << d);
    if (o < kWitBlock) add = fp_mul<FrParams>(s_e[o], mult);
    __syncthreads();
    s_e[threadIdx.x] = fp_add<FrParams>(s_e[threadIdx.x], add);
    __syncthreads();
  }
}
// pass 1: block value F_b = s_blk(first coefficient of block b)
static __global__ void __launch_bounds__(kWitBlock)
k_witness_block_sums(const Fr* __restrict__ f, size_t nchunks, int log_l, const Fr* __restrict__ tab, Fr* __restrict__ F) {
  __shared__ Fr s_e[kWitBlock];
  const size_t t = (size_t)blockIdx.x * kWitBlock + threadIdx.x;
  const size_t L = size_t(1) << log_l;
  const Fr u = fp_load(tab);
  s_e[threadIdx.x] = t < nchunks ? wit_chunk_eval(f, t * L, L, u) : fp_zero<FrParams>();
  __syncthreads();
  wit_block_scan(s_e, tab, log_l);
  if (threadIdx.x == 0) fp_store(F + blockIdx.x, s_e[0]);
}
// pass 2 (one block): C[b] = s(first coefficient of block b) = F_b + M * C[b+1], M = u^(L*256); C[nb] = 0.
// Thread t owns `per` = 2^log_per consecutive blocks; Kogge-Stone across threads.
static __global__ void __launch_bounds__(1024)
k_witness_carry(const Fr* __restrict__ F, size_t nb, int log_per, int log_m, const Fr* __restrict__ tab, Fr* __restrict__ C) {
  __shared__ Fr s_v[1024];
  const size_t per = size_t(1) << log_per;
  const size_t b0 = (size_t)threadIdx.x * per;
  const Fr M = fp_load(tab + log_m);
  Fr acc = fp_zero<FrParams>();       // local suffix value at b0: sum_{k<per} F[b0+k] M^k
  for (size_t k = per; k-- > 0;) {
    const Fr v = b0 + k < nb ? fp_load(F + b0 + k) : fp_zero<FrParams>();
    acc = fp_add<FrParams>(fp_mul<FrParams>(acc, M), v);
  }
  s_v[threadIdx.x] = acc;
  __syncthreads();
  for (int d = 0; (1 << d) < 1024; d++) {
    const Fr mult = fp_load(tab + log_m + log_per + d);
    Fr add = fp_zero<FrParams>();
    const int o = threadIdx.x + (1 << d);
    if (o < 1024) add = fp_mul<FrParams>(s_v[o], mult);
    __syncthreads();
    s_v[threadIdx.x] = fp_add<FrParams>(s_v[threadIdx.x], add);
    __syncthreads();
  }
  // s_v[t] = s(block b0); walk the thread's own blocks from the top with the carry of the next thread
  Fr cur = threadIdx.x + 1 < 1024 ? s_v[threadIdx.x + 1] : fp_zero<FrParams>();
  for (size_t k = per; k-- > 0;) {
    if (b0 + k < nb) {
      cur = fp_add<FrParams>(fp_mul<FrParams>(cur, M), fp_load(F + b0 + k));
      fp_store(C + b0 + k, cur);
    }
  }
  if (threadIdx.x == 0) fp_store(C + nb, fp_zero<FrParams>());
}
// pass 3: h[i] = s(i+1) for every coefficient of the block, from the block carry C[b+1]
static __global__ void __launch_bounds__(kWitBlock)
k_witness_write(const Fr* __restrict__ f, size_t nchunks, int log_l, const Fr* __restrict__ tab, const Fr* __restrict__ C,
                Fr* __restrict__ h) {
  __shared__ Fr s_e[kWitBlock];
  const size_t t = (size_t)blockIdx.x * kWitBlock + threadIdx.x;
  const size_t L = size_t(1) << log_l;
  const Fr u = fp_load(tab);
  const bool live = t < nchunks;
  s_e[threadIdx.x] = live ? wit_chunk_eval(f, t * L, L, u) : fp_zero<FrParams>();
  __syncthreads();
  wit_block_scan(s_e, tab, log_l);
  if (!live) return;
  // s(a_{t+1}) = s_blk(a_{t+1}) + (u^L)^(255 - threadIdx) * C[b+1]      (s_blk of the chunk after the block's last is 0)
  const Fr carry_blk = fp_load(C + blockIdx.x + 1);
  Fr cur = fp_mul<FrParams>(carry_blk, pow_from_tab(tab + log_l, (unsigned long long)(kWitBlock - 1 - threadIdx.x)));
  if (threadIdx.x + 1 < kWitBlock) cur = fp_add<FrParams>(cur, s_e[threadIdx.x + 1]);
  const size_t a = t * L;
  for (size_t j = L; j-- > 0;) {
    fp_store(h + a + j, cur);
    cur = fp_add<FrParams>(fp_mul<FrParams>(cur, u), fp_load(f + a + j));
  }
}

}  // namespace ja
