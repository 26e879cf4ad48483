// TMA-staged variant of the streaming family-S round kernels (ADD / SUB / IDENT: fused bind + split-eq evaluation,
// LowToHigh) for polynomial slabs that do not fit the cache hierarchy.
//
// Why: the register-staged kernel (fused_kernels.cuh: k_round_s) issues the eight 16-byte loads of a pair's 128-byte
// row, waits, then runs ~500 integer instructions per row; at 80 registers only ~6 warps per scheduler are resident
// and neither DRAM (43 %) nor the integer pipe (57 %) saturates (profiles/r1_ncu_fused_kernels.md).  Here the loads
// leave the register file: one elected lane per warp streams 4 KB slabs (32 rows of 4 Fr = one fused pair per
// thread) through a ring of shared-memory slots with cp.async.bulk.tensor (TMA, SASS UTMALDG) completing on an
// mbarrier per slot; a warp only ever waits for a slab it requested two slabs earlier (no block-wide barrier), so it is
// compute-ready almost always and the pipe that remains is the Montgomery arithmetic.
//
// Layout: a polynomial of current length n is described to TMA as a 2-D tensor [n/4 rows][32 x u32] (row = the 128
// bytes a0..a3 that produce one pair (lo, hi) of the bound array) with SWIZZLE_128B: 16-byte chunk c of row t lands at
// chunk position c ^ (t & 7), so the per-thread row reads (8 x LDS.128, one row per thread) are bank-conflict free.
// Reference semantics are those of k_round_s (split_eq_poly.rs:526-597 with the Add/Sub/identity closures, and
// dense_mlpoly.rs:219-239 for the bind); results are the same canonical field elements.
#pragma once
#include <cuda.h>
#include "fused_kernels.cuh"

namespace ja {

constexpr int kTmaRows = 256;                      // rows (= fused pairs) per block step == threads per block
constexpr int kTmaWarps = kTmaRows / 32;
constexpr int kTmaSlabBytes = 32 * 128;            // one warp's slab: 32 rows, 4 KB
constexpr int kTmaSlots = 3;                       // ring depth per warp
constexpr size_t kTmaSmemBytes = (size_t)kTmaWarps * kTmaSlots * kTmaSlabBytes + 1024;   // + alignment slack (SWIZZLE_128B: 1024 B)

JA_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
JA_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
JA_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
JA_DEV bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
JA_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// one slab: rows [row0, row0 + 32) of the tensor -> smem slot, completion counted in bytes on `bar`
JA_DEV void tma_load_slab(void* dst, const CUtensorMap* tm, int row0, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(0), "r"(row0), "r"(smem_u32(bar))
      : "memory");
}
// 16-byte chunk `c` of row `t` of a SWIZZLE_128B slab
JA_DEV uint4 slab_chunk(const uint8_t* slab, int t, int c) {
  return *reinterpret_cast<const uint4*>(slab + t * 128 + ((c ^ (t & 7)) << 4));
}
JA_DEV Fr slab_fr(const uint8_t* slab, int t, int e) {
  const uint4 lo = slab_chunk(slab, t, 2 * e), hi = slab_chunk(slab, t, 2 * e + 1);
  Fr r;
  r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
  r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
  return r;
}

// KID: 0 ADD, 1 SUB, 6 IDENT.  G (pairs of the bound array) must be a multiple of 256.
template <int KID>
__global__ void __launch_bounds__(kTmaRows, 2)
k_round_s_tma(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1, Fr* __restrict__ out0,
              Fr* __restrict__ out1, Challenge r, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in, int bits_in, size_t G,
              size_t slabs_per_block, Fr* partials, unsigned int* counter, Publish pub, size_t g_off) {
  constexpr int NP = (KID == 6) ? 1 : 2;
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // every warp streams its own 32 rows of each 256-row step through a private ring: no block-wide synchronisation
  uint8_t* slabs = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023)) +
                   (size_t)warp * kTmaSlots * kTmaSlabBytes;
  __shared__ __align__(8) uint64_t full_all[kTmaWarps][kTmaSlots];
  uint64_t* full = full_all[warp];
  // Work split: block b owns the contiguous steps [s_begin, s_end) of 256 rows; inside it warp w owns a CONTIGUOUS run of
  // 32-row slabs, so that consecutive rows of a thread are 32 apart and stay under one split-eq outer index for
  // 2^bits_in / 32 iterations: the eq-weighted sum accumulates unreduced (delayed Montgomery reduction) and is only
  // reduced every 16 products or when the outer index changes.
  const size_t n_steps = G / kTmaRows;
  const size_t s_begin = (size_t)blockIdx.x * slabs_per_block;
  size_t s_end = s_begin + slabs_per_block;
  if (s_end > n_steps) s_end = n_steps;
  const size_t my_steps = s_begin < s_end ? s_end - s_begin : 0;       // this warp: my_steps slabs of 32 rows
  const size_t row_first = s_begin * kTmaRows + (size_t)warp * my_steps * 32;
  const int n_items = (int)my_steps * NP;                              // (slab, polynomial) units, in consumption order
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kTmaSlots; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int q) {
    const int slot = q % kTmaSlots;
    mbar_expect_tx(&full[slot], kTmaSlabBytes);
    tma_load_slab(slabs + (size_t)slot * kTmaSlabBytes, (NP == 2 && (q % NP) == 1) ? &tm1 : &tm0, (int)(row_first + (size_t)(q / NP) * 32), &full[slot]);
  };
  if (lane == 0)
    for (int q = 0; q < kTmaSlots && q < n_items; q++) issue(q);

  Fr outer = fp_zero<FrParams>(), inner = fp_zero<FrParams>();
  FpWide wide = fpw_zero();
  int pending = 0;
  const size_t mask_in = (size_t(1) << bits_in) - 1;
  size_t cur_xout = ~size_t(0);
  int q = 0;
  for (size_t it = 0; it < my_steps; it++) {
    const size_t g = row_first + it * 32 + lane;
    Fr lo[NP];
#pragma unroll
    for (int p = 0; p < NP; p++, q++) {
      const int slot = q % kTmaSlots;
      const uint8_t* sl = slabs + (size_t)slot * kTmaSlabBytes;
      mbar_wait(&full[slot], (uint32_t)(q / kTmaSlots) & 1u);
      const Fr a0 = slab_fr(sl, lane, 0), a1 = slab_fr(sl, lane, 1), a2 = slab_fr(sl, lane, 2), a3 = slab_fr(sl, lane, 3);
      __syncwarp();                                           // every row of the slot is in registers: refill it
      if (lane == 0 && q + kTmaSlots < n_items) issue(q + kTmaSlots);
      lo[p] = fp_add<FrParams>(a0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a1, a0), r));
      const Fr hi = fp_add<FrParams>(a2, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a3, a2), r));
      Fr* o = (p == 0 ? out0 : out1) + 2 * g;
      fp_store(o, lo[p]);
      fp_store(o + 1, hi);
    }
    const size_t x_out = (g + g_off) >> bits_in;
    if (x_out != cur_xout || pending == 16) {
      if (pending) { inner = fp_add<FrParams>(inner, fpw_reduce<FrParams>(wide)); wide = fpw_zero(); pending = 0; }
      if (x_out != cur_xout) {
        if (cur_xout != ~size_t(0)) {
          outer = fp_add<FrParams>(outer, fp_mul<FrParams>(fp_load(e_out + cur_xout), inner));
          inner = fp_zero<FrParams>();
        }
        cur_xout = x_out;
      }
    }
    Fr v;
    if (KID == 6) v = lo[0];
    else if (KID == 0) v = fp_add<FrParams>(lo[0], lo[NP - 1]);
    else v = fp_sub<FrParams>(lo[0], lo[NP - 1]);
    fpw_mul_acc<FrParams>(wide, fp_load(e_in + ((g + g_off) & mask_in)), v);
    pending++;
  }
  if (pending) inner = fp_add<FrParams>(inner, fpw_reduce<FrParams>(wide));
  if (cur_xout != ~size_t(0)) outer = fp_add<FrParams>(outer, fp_mul<FrParams>(fp_load(e_out + cur_xout), inner));
  Fr acc[1] = {outer};
  grid_sum_ex<1>(acc, partials, counter, pub.vals, blockIdx.x, gridDim.x, pub.value);
}

}  // namespace ja
