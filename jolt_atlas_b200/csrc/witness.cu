// Witness generation of a fused node ON the device (jolt-atlas-core/src/onnx_proof/witness.rs:142-214 generate_node_witnesses):
// the committed one-hot polynomials of a fused node are the 4-bit chunks of two integer tensors the prover re-derives from the
// node's i32 operands,
//   acc       = einsum_acc_i64 / mul_acc_i64 / left +- right in i64            atlas-onnx-tracer/src/ops/einsum.rs:248-258, mul.rs:51-60
//   quotient  = acc.div_euclid(2^S)   (the pre-clamp value, the lookup index of the 64-bit saturating-clamp read-raf)   ops/mod.rs:224-232
//   remainder = acc.rem_euclid(2^S)   (range-checked to [0, 2^S))                                                       ops/mod.rs:237-249
//   ClampRaD(d)[t]            = (quotient[t] as u64 >> 4 (15 - d)) & 15,  d < 16   clamp_lookups/mod.rs:245-252, joltworks/src/config.rs:75-77
//   RescaleRemainderRaD(d)[t] = (remainder[t] >> 4 (D - 1 - d)) & 15,   D = ceil(S / 4)   witness.rs:601-617
// padded with zeros to the power-of-two cycle domain T.  The reference computes them on the host and the one-hot index arrays
// (d x T entries per node: 37 MB for a nanoGPT proof, 1.09 GB at GPT-2 size) would cross PCIe; here the operands (weights are
// resident anyway) go in and the address batches, the lookup indices of ps_shout and the clamped i32 output are BORN in HBM.
// Integer work on CUDA cores: exact i32 x i32 -> i64 (IMAD.WIDE); no tensor cores (an int8 / bf16 GEMM would not be exact).
#include <memory>

#include "common.hpp"

namespace {

constexpr int kWitTile = 16;

// acc[i][j] = sum_k A[i][k] * B[k][j], i64.  16 x 16 output tile per block, operands staged through shared memory.
__global__ void __launch_bounds__(kWitTile * kWitTile)
k_wit_matmul(const int* __restrict__ A, const int* __restrict__ B, size_t m, size_t k, size_t n, long long* __restrict__ acc) {
  __shared__ int sa[kWitTile][kWitTile + 1], sb[kWitTile][kWitTile + 1];
  const int tx = threadIdx.x % kWitTile, ty = threadIdx.x / kWitTile;
  const size_t row = (size_t)blockIdx.y * kWitTile + ty, col = (size_t)blockIdx.x * kWitTile + tx;
  long long s = 0;
  for (size_t k0 = 0; k0 < k; k0 += kWitTile) {
    sa[ty][tx] = (row < m && k0 + tx < k) ? __ldg(A + row * k + k0 + tx) : 0;
    sb[ty][tx] = (k0 + ty < k && col < n) ? __ldg(B + (k0 + ty) * n + col) : 0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kWitTile; q++) s += (long long)sa[ty][q] * (long long)sb[q][tx];
    __syncthreads();
  }
  if (row < m && col < n) acc[row * n + col] = s;
}
// element-wise accumulations: op 1 = a * b (mul_acc_i64), 2 = a + b, 3 = a - b (sat_binop_intermediate)
__global__ void __launch_bounds__(kBlock) k_wit_elementwise(int op, const int* __restrict__ a, const int* __restrict__ b, size_t n, long long* __restrict__ acc) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long x = __ldg(a + i), y = __ldg(b + i);
    acc[i] = op == 1 ? x * y : (op == 2 ? x + y : x - y);
  }
}
// quotient / remainder / chunks / clamped output for entry t < T (entries >= n_valid are the zero padding)
__global__ void __launch_bounds__(kBlock)
k_wit_chunks(const long long* __restrict__ acc, size_t n_valid, size_t T, uint32_t scale_bits, uint32_t d_rem,
             unsigned long long* __restrict__ idx /* T */, uint32_t* __restrict__ clamp_k /* 16 x T */, uint32_t* __restrict__ rem_k /* d_rem x T */,
             int* __restrict__ out_i32 /* T */, unsigned long long* __restrict__ rem_idx /* T, or null */) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += stride) {
    const long long v = t < n_valid ? acc[t] : 0ll;
    const long long q = v >> scale_bits;                               // div_euclid by a power of two = arithmetic shift (floor)
    const unsigned long long r = (unsigned long long)v & ((1ull << scale_bits) - 1);     // rem_euclid
    const unsigned long long uq = (unsigned long long)q;
    idx[t] = uq;
    if (rem_idx) rem_idx[t] = r;
#pragma unroll
    for (int d = 0; d < 16; d++) clamp_k[(size_t)d * T + t] = (uint32_t)((uq >> (4 * (15 - d))) & 15ull);
    for (uint32_t d = 0; d < d_rem; d++) rem_k[(size_t)d * T + t] = (uint32_t)((r >> (4 * (d_rem - 1 - d))) & 15ull);
    const long long lo = -2147483648ll, hi = 2147483647ll;
    out_i32[t] = (int)(q < lo ? lo : (q > hi ? hi : q));               // clamp_to_i32: the node's output
  }
}

}  // namespace

struct ja_witness {
  unsigned long long* d_idx = nullptr;     // T lookup indices of the clamp read-raf (quotient as u64)
  int* d_out = nullptr;                    // T clamped outputs
  unsigned long long* d_rem_idx = nullptr; // T rescale remainders: the lookup indices of the remainder range check (absent for scale 0)
  uint32_t scale_bits = 0;
  ja_addr* clamp = nullptr;                // 16 x T, K = 16
  ja_addr* rem = nullptr;                  // ceil(S / 4) x T, K = 16 (absent for scale 0)
  size_t T = 0;
};

extern "C" {

int32_t ja_witness_fused(ja_ctx* c, int32_t op, const ja_tensor_i32* A, const ja_tensor_i32* B, uint32_t scale_bits, size_t T, ja_witness** out) {
  JA_REQUIRE(c && A && B && out && op >= JA_WIT_EINSUM_MK_KN && op <= JA_WIT_SUB && is_pow2(T) && scale_bits < 31, "ja_witness_fused: bad argument");
  size_t m = 0, k = 0, n = 0, n_valid = 0;
  if (op == JA_WIT_EINSUM_MK_KN) {
    JA_REQUIRE(A->cols == B->rows, "ja_witness_fused: contraction lengths differ (A is m x k, B is k x n)");
    m = A->rows; k = A->cols; n = B->cols; n_valid = m * n;
  } else {
    JA_REQUIRE(A->rows == B->rows && A->cols == B->cols, "ja_witness_fused: element-wise operands must have the same shape");
    n_valid = A->rows * A->cols;
  }
  JA_REQUIRE(n_valid <= T, "ja_witness_fused: T smaller than the number of outputs");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  const uint32_t d_rem = (scale_bits + 3) / 4;
  std::unique_ptr<ja_witness> w(new ja_witness());
  w->T = T; w->scale_bits = scale_bits;
  long long* d_acc = nullptr;
  int32_t st;
  if ((st = dev_alloc(c, n_valid * 8, (void**)&d_acc))) return st;
  uint32_t *d_ck = nullptr, *d_rk = nullptr;
  if ((st = dev_alloc(c, T * 8, (void**)&w->d_idx)) || (st = dev_alloc(c, T * 4, (void**)&w->d_out)) ||
      (st = dev_alloc(c, 16 * T * 4, (void**)&d_ck)) || (d_rem && (st = dev_alloc(c, (size_t)d_rem * T * 4, (void**)&d_rk))) ||
      (d_rem && (st = dev_alloc(c, T * 8, (void**)&w->d_rem_idx)))) {
    dev_free(c, d_acc); dev_free(c, w->d_idx); dev_free(c, w->d_out); dev_free(c, d_ck); dev_free(c, d_rk); dev_free(c, w->d_rem_idx);
    return st;
  }
  if (op == JA_WIT_EINSUM_MK_KN) {
    const dim3 grid((unsigned)((n + kWitTile - 1) / kWitTile), (unsigned)((m + kWitTile - 1) / kWitTile));
    JA_LAUNCH(c, KC_TENSOR_FOLD, k_wit_matmul<<<grid, kWitTile * kWitTile, 0, c->stream>>>(A->data, B->data, m, k, n, d_acc));
  } else {
    JA_LAUNCH(c, KC_TENSOR_FOLD, k_wit_elementwise<<<grid_for(n_valid), kBlock, 0, c->stream>>>(op, A->data, B->data, n_valid, d_acc));
  }
  JA_LAUNCH(c, KC_CONVERT, k_wit_chunks<<<grid_for(T), kBlock, 0, c->stream>>>(d_acc, n_valid, T, scale_bits, d_rem, w->d_idx, d_ck, d_rk, w->d_out, w->d_rem_idx));
  JA_CUDA(cudaGetLastError());
  dev_free(c, d_acc);                                  // stream-ordered reuse
  w->clamp = new ja_addr();
  w->clamp->d_k = d_ck; w->clamp->d = 16; w->clamp->T = T; w->clamp->K = 16;
  if (d_rem) { w->rem = new ja_addr(); w->rem->d_k = d_rk; w->rem->d = d_rem; w->rem->T = T; w->rem->K = 16; }
  *out = w.release();
  return JA_OK;
}

const ja_addr* ja_witness_clamp_addr(const ja_witness* w) { return w ? w->clamp : nullptr; }
const ja_addr* ja_witness_rem_addr(const ja_witness* w) { return w ? w->rem : nullptr; }

int32_t ja_witness_to_host(ja_ctx* c, const ja_witness* w, uint64_t* out_idx, int32_t* out_i32, uint32_t* out_clamp_k, uint32_t* out_rem_k) {
  JA_REQUIRE(c && w, "ja_witness_to_host: null argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  if (out_idx) JA_CUDA(cudaMemcpyAsync(out_idx, w->d_idx, w->T * 8, cudaMemcpyDeviceToHost, c->stream));
  if (out_i32) JA_CUDA(cudaMemcpyAsync(out_i32, w->d_out, w->T * 4, cudaMemcpyDeviceToHost, c->stream));
  if (out_clamp_k) JA_CUDA(cudaMemcpyAsync(out_clamp_k, w->clamp->d_k, 16 * w->T * 4, cudaMemcpyDeviceToHost, c->stream));
  if (out_rem_k && w->rem) JA_CUDA(cudaMemcpyAsync(out_rem_k, w->rem->d_k, w->rem->d * w->T * 4, cudaMemcpyDeviceToHost, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));
  return JA_OK;
}

// ps_shout state over the witness' lookup indices (device to device: nothing crosses PCIe)
int32_t ja_psshout_from_witness(ja_ctx* c, const ja_witness* w, const uint64_t* r_cycle, size_t log_t, uint32_t log_k, uint32_t phases, ja_psshout** out) {
  JA_REQUIRE(c && w && out && (size_t(1) << log_t) == w->T, "ja_psshout_from_witness: T = 2^log_t");
  return ja_psshout_new_dev(c, w->d_idx, w->T, r_cycle, log_t, log_k, phases, out);
}

// ps_shout state of the remainder range check (IdentityRCProver over the rescale remainders): LOG_K = the rescale bits
int32_t ja_psshout_from_witness_rem(ja_ctx* c, const ja_witness* w, const uint64_t* r_cycle, size_t log_t, uint32_t phases, ja_psshout** out) {
  JA_REQUIRE(c && w && out && (size_t(1) << log_t) == w->T, "ja_psshout_from_witness_rem: T = 2^log_t");
  JA_REQUIRE(w->d_rem_idx, "ja_psshout_from_witness_rem: the node has no rescale remainder (scale 0)");
  return ja_psshout_new_dev(c, w->d_rem_idx, w->T, r_cycle, log_t, w->scale_bits, phases, out);
}

void ja_witness_free(ja_ctx* c, ja_witness* w) {
  if (!c || !w) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  dev_free(c, w->d_idx); dev_free(c, w->d_out); dev_free(c, w->d_rem_idx);
  if (w->clamp) { dev_free(c, w->clamp->d_k); delete w->clamp; }
  if (w->rem) { dev_free(c, w->rem->d_k); delete w->rem; }
  delete w;
}

}  // extern "C"
