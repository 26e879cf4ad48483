// Test hooks of the C ABI: the device field arithmetic (fp.cuh) applied element-wise to host arrays, so that the golden
// edge vectors of tests/golden/field.json reach fp_mul / fp_add / fp_sub / fp_mul_challenge / fp_from_i64 and the delayed
// reduction (fpw_mul_acc + fpw_reduce) DIRECTLY instead of only through bind / round-evaluation kernels.
// Reference semantics: joltworks/src/field/ark.rs:76-297 (JoltField for ark_bn254::Fr), field/challenge/macros.rs:274-286
// (F * MontU128Challenge), field/mod.rs:286-310 (mul_unreduced::<9> + from_montgomery_reduce).
#include "common.hpp"

namespace {

__global__ void __launch_bounds__(kBlock) k_test_field(int op, const Fr* __restrict__ a, const Fr* __restrict__ b, Fr* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr x = fp_load(a + i), y = fp_load(b + i);
  Fr r;
  switch (op) {
    case JA_TEST_FR_MUL: r = fp_mul<FrParams>(x, y); break;
    case JA_TEST_FR_ADD: r = fp_add<FrParams>(x, y); break;
    case JA_TEST_FR_SUB: r = fp_sub<FrParams>(x, y); break;
    case JA_TEST_FR_MUL_CHALLENGE: {
      Challenge c; c.c[0] = y.l[4]; c.c[1] = y.l[5]; c.c[2] = y.l[6]; c.c[3] = y.l[7];
      r = fp_mul_challenge<FrParams>(x, c); break;
    }
    case JA_TEST_FR_FROM_I64: {
      const long long v = (long long)((unsigned long long)x.l[0] | ((unsigned long long)x.l[1] << 32));
      r = fp_from_i64<FrParams>(v); break;
    }
    case JA_TEST_FR_MUL_WIDE: {            // delayed reduction: 16 copies of the product accumulated as 512-bit integers, one reduction
      FpWide w = fpw_zero();
#pragma unroll 1
      for (int k = 0; k < 16; k++) fpw_mul_acc<FrParams>(w, x, y);
      r = fpw_reduce<FrParams>(w); break;   // == 16 * x * y
    }
    case JA_TEST_FR_NEG: r = fp_neg<FrParams>(x); break;
    case JA_TEST_FR_SQR: r = fp_sqr<FrParams>(x); break;
    default: r = fp_zero<FrParams>(); break;
  }
  fp_store(out + i, r);
}

}  // namespace

extern "C" int32_t ja_test_field_ops(ja_ctx* c, int32_t op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
  JA_REQUIRE(c && a && b && out && n >= 1 && n <= (size_t(1) << 20) && op >= 0 && op <= JA_TEST_FR_SQR, "ja_test_field_ops: bad argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  Fr *da = nullptr, *db = nullptr, *dout = nullptr;
  int32_t st;
  if ((st = dev_alloc(c, n * sizeof(Fr), (void**)&da))) return st;
  if ((st = dev_alloc(c, n * sizeof(Fr), (void**)&db))) { dev_free(c, da); return st; }
  if ((st = dev_alloc(c, n * sizeof(Fr), (void**)&dout))) { dev_free(c, da); dev_free(c, db); return st; }
  cudaError_t e = cudaMemcpyAsync(da, a, n * sizeof(Fr), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(db, b, n * sizeof(Fr), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    JA_LAUNCH(c, KC_MISC, k_test_field<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, c->stream>>>(op, da, db, dout, n));
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, n * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dev_free(c, da); dev_free(c, db); dev_free(c, dout);
  if (e != cudaSuccess) return fail(JA_ERR_CUDA, std::string("ja_test_field_ops: ") + cudaGetErrorString(e));
  return JA_OK;
}
