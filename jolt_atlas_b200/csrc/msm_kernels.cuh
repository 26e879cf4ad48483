// Multi-scalar multiplication over BN254 G1: batched Pippenger (bucket method) for sm_100a.
//
// Replaces the MSM family the reference reaches through joltworks/src/msm/mod.rs:27-181
// (VariableBaseMSM::msm by scalar width -> ark-ec, external), :309-318 (batch_msm), and
// joltworks/src/poly/commitment/hyperkzg/mod.rs:520-596 (commit_one_hot / batch_commit_one_hot ->
// jolt_optimizations::batch_g1_additions_multi, external).  The result of an MSM is a group element, so
// any complete algorithm yields the reference's affine point bit for bit.
//
// One pipeline serves a whole BATCH of MSMs that share the SRS (HyperKZG::open commits l-1 folded
// polynomials at once; witness commitment sums hundreds of one-hot index lists):
//   1  k_msm_hist      signed-digit recoding of every scalar, histogram over (msm, window, bucket)
//   2  k_scan*         exclusive scan of the histogram -> bucket offsets
//   3  k_msm_scatter   counting-sort scatter of (base index | sign) by bucket
//   4  k_msm_accumulate  every thread sums a FIXED-SIZE run of T sorted entries (mixed XYZZ additions);
//                      runs are cut at bucket boundaries, so load balance does not depend on how the
//                      digits are distributed (one-hot commits are a single bucket, small scalars are
//                      skewed).  Buckets inside one run are final; buckets spanning runs leave one
//                      partial per run (head / tail).
//   5  k_msm_combine   per bucket: add the partials of the runs it spans (wide buckets go to a
//                      block-cooperative kernel, k_msm_combine_big)
//   6  k_msm_bucket_reduce / k_msm_part_sum   sum_b (b+1) * B_b per window: per-thread running sums over
//                      segments of buckets + one small scalar multiple, then a two-stage block tree per window
//   7  k_msm_final     Horner over the windows (c doublings each) and conversion to affine
// Integer-throughput bound (8M+2S per pair and window); the 64 B base gathers are hidden behind it.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "ec.cuh"

namespace ja {

enum MsmKind : uint32_t {
  MSM_FR = 0,      // Fr scalars, Montgomery form (LargeScalars)
  MSM_U8 = 1, MSM_U16 = 2, MSM_U32 = 3, MSM_U64 = 4,   // msm_u8/u16/u32/u64
  MSM_I32 = 5, MSM_I64 = 6,                            // I32Scalars / I64Scalars: msm(pos) - msm(neg) == signed digits
  MSM_INDEXED = 7  // sum of selected bases (one-hot commit): scalars = u64 base indices, every digit is 1
};

struct MsmDesc {
  const void* scalars;
  uint32_t n;            // pairs in this MSM
  uint32_t kind;
  uint32_t c;            // window bits (signed digits in [-2^(c-1), 2^(c-1)])
  uint32_t nwin;         // windows
  uint32_t nb;           // buckets per window = 2^(c-1)
  uint32_t bucket_base;  // first global bucket id of this MSM
  uint32_t win_base;     // first global window id of this MSM
  uint32_t entry_base;   // prefix sum of n over the batch (thread -> msm lookup)
  uint32_t base_offset;  // SRS index of base 0 (ignored for MSM_INDEXED)
  uint32_t fixed_stride; // != 0: fixed-base window table in use (stride = SRS length): every window shares ONE bucket set and
                         // digit w of scalar j selects base table_off + w * stride + j (= 2^(c w) * G_j); 0: classic per-window buckets
  uint32_t table_off;    // first entry of the window table this job uses (0 unless fixed_stride != 0)
  uint32_t sub;          // fixed_stride != 0: sub-windows per table window (table window bits / c), >= 1.  sub > 1 = a SHORT job
                         // on the table: digit v uses base table window v / sub and bucket set v % sub, so the final Horner
                         // runs (sub - 1) * c doublings instead of one doubling per scalar bit
};

constexpr uint32_t kSignBit = 0x80000000u;

// msm id of global scalar position g: largest m with entry_base[m] <= g
JA_DEV uint32_t msm_find(const MsmDesc* __restrict__ d, uint32_t count, uint32_t g) {
  uint32_t lo = 0, hi = count - 1;
  while (lo < hi) {
    uint32_t mid = (lo + hi + 1) >> 1;
    if (d[mid].entry_base <= g) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// magnitude limbs (little-endian u32, zero padded to 10 words) and sign of scalar j of MSM d
JA_DEV bool msm_load_scalar(const MsmDesc& d, uint32_t j, uint32_t (&s)[10]) {
#pragma unroll
  for (int i = 0; i < 10; i++) s[i] = 0;
  bool neg = false;
  switch (d.kind) {
    case MSM_FR: {
      Fr a = fp_load(reinterpret_cast<const Fr*>(d.scalars) + j);
      const uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
      Fr r;
      fp_mont_rows<FrParams, 8>(r.l, a.l, one);     // a * R^-1: Montgomery -> canonical integer
      fp_final_sub<FrParams>(r.l);
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = r.l[i];
      break; }
    case MSM_U8: s[0] = reinterpret_cast<const uint8_t*>(d.scalars)[j]; break;
    case MSM_U16: s[0] = reinterpret_cast<const uint16_t*>(d.scalars)[j]; break;
    case MSM_U32: s[0] = reinterpret_cast<const uint32_t*>(d.scalars)[j]; break;
    case MSM_U64: { unsigned long long v = reinterpret_cast<const unsigned long long*>(d.scalars)[j];
      s[0] = (uint32_t)v; s[1] = (uint32_t)(v >> 32); break; }
    case MSM_I32: { int v = reinterpret_cast<const int*>(d.scalars)[j];
      neg = v < 0; s[0] = neg ? (uint32_t)(-(long long)v) : (uint32_t)v; break; }
    case MSM_I64: { long long v = reinterpret_cast<const long long*>(d.scalars)[j];
      neg = v < 0; unsigned long long m = neg ? (unsigned long long)(-(v + 1)) + 1ull : (unsigned long long)v;
      s[0] = (uint32_t)m; s[1] = (uint32_t)(m >> 32); break; }
    default: break;
  }
  return neg;
}

// Warp-aggregated atomics: lanes that hit the same bucket elect a leader which issues ONE atomicAdd for the
// group (degenerate top windows, small scalars and one-hot sums put millions of entries on a few counters).
JA_DEV uint32_t warp_agg_atomic_add(uint32_t* ctr, uint32_t key, bool active) {
  // returns this lane's slot: base (leader's atomicAdd result) + rank among same-key lanes; undefined if !active
  const uint32_t amask = __ballot_sync(0xffffffffu, active);
  uint32_t res = 0;
  if (active) {
    const uint32_t peers = __match_any_sync(amask, key);
    const int leader = __ffs(peers) - 1;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(ctr + key, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    res = base + __popc(peers & ((1u << lane) - 1));
  }
  return res;
}

// Signed digit w of scalar j (0 = none): key = global bucket id, payload = base index | sign.
struct MsmDigitIter {
  uint32_t s[10]; bool neg; uint32_t carry; uint32_t base;
};
JA_DEV bool msm_digit(const MsmDesc& d, MsmDigitIter& it, uint32_t w, uint32_t& key, uint32_t& payload) {
  const uint32_t c = d.c, nb = d.nb, full = 1u << c;
  const uint32_t bit = w * c, word = bit >> 5, sh = bit & 31;
  unsigned long long two = ((unsigned long long)it.s[word + 1] << 32) | it.s[word];
  uint32_t raw = (uint32_t)((two >> sh) & (full - 1)) + it.carry;
  bool dneg = false;
  if (raw > nb) { raw = full - raw; it.carry = 1; dneg = true; } else it.carry = 0;
  if (!raw) return false;
  key = d.bucket_base + (d.fixed_stride ? (w % d.sub) * nb : w * nb) + (raw - 1);
  payload = (it.base + d.table_off + (d.fixed_stride ? w / d.sub : 0u) * d.fixed_stride) | ((dneg != it.neg) ? kSignBit : 0u);
  return true;
}

// SCATTER == false: histogram; SCATTER == true: counting-sort scatter through `ctr` (a copy of the offsets).
// Whole warps stay converged (threads past the end are inactive lanes) so the aggregated atomics are legal.
template <bool SCATTER>
__global__ void __launch_bounds__(256)
k_msm_digits(const MsmDesc* __restrict__ descs, uint32_t count, uint32_t total_n, uint32_t max_nwin,
             uint32_t* __restrict__ ctr, uint32_t* __restrict__ entries, bool plain = false /* full-width field scalars only:
             digits of random 254-bit scalars almost never collide inside a warp, so the match_any aggregation is pure overhead */) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = g < total_n;
  MsmDesc d; d.kind = MSM_INDEXED; d.nwin = 0; d.entry_base = 0; d.bucket_base = 0;
  if (live) d = descs[msm_find(descs, count, g)];
  const uint32_t j = g - d.entry_base;
  if (__all_sync(0xffffffffu, !live || d.kind == MSM_INDEXED)) {
    uint32_t payload = 0;
    if (live) payload = (uint32_t)reinterpret_cast<const unsigned long long*>(d.scalars)[j];
    const uint32_t slot = warp_agg_atomic_add(ctr, d.bucket_base, live);
    if (SCATTER && live) entries[slot] = payload;
    return;
  }
  MsmDigitIter it;
  it.neg = false; it.carry = 0; it.base = d.base_offset + j;
  if (live && d.kind != MSM_INDEXED) it.neg = msm_load_scalar(d, j, it.s);
  else {
#pragma unroll
    for (int i = 0; i < 10; i++) it.s[i] = 0;
  }
  if (live && d.kind == MSM_INDEXED) {      // mixed warp (batch boundary): plain atomics for the indexed lanes
    const uint32_t slot = atomicAdd(ctr + d.bucket_base, 1u);
    if (SCATTER) entries[slot] = (uint32_t)reinterpret_cast<const unsigned long long*>(d.scalars)[j];
  }
  const bool digits = live && d.kind != MSM_INDEXED;
  for (uint32_t w = 0; w < max_nwin; w++) {
    uint32_t key = 0, payload = 0;
    const bool has = digits && w < d.nwin && msm_digit(d, it, w, key, payload);
    if (plain) {
      if (has) {
        if (SCATTER) entries[atomicAdd(ctr + key, 1u)] = payload;
        else atomicAdd(ctr + key, 1u);                 // result unused: a reduction
      }
    } else {
      const uint32_t slot = warp_agg_atomic_add(ctr, key, has);
      if (SCATTER && has) entries[slot] = payload;
    }
  }
}

// ---- exclusive scan of a u32 array (histogram -> offsets), 3 launches ------------------------------
constexpr int kScanBlock = 512, kScanItems = 8, kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0, winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += t; }
    s_warp[lane] = winc - w;          // exclusive prefix of warp totals
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  uint32_t res = inc - v + s_warp[warp];
  if (total) *total = s_warp[32];
  __syncthreads();
  return res;
}

// in-place: data[i] <- exclusive prefix within its tile; tile_sums[tile] = tile total
static __global__ void __launch_bounds__(kScanBlock) k_scan_tiles(uint32_t* data, size_t n, uint32_t* tile_sums) {
  __shared__ uint32_t s_warp[33];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems], sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) { v[i] = base + i < n ? data[base + i] : 0; sum += v[i]; }
  uint32_t total;
  uint32_t pre = block_excl_scan(sum, s_warp, &total);
#pragma unroll
  for (int i = 0; i < kScanItems; i++) { if (base + i < n) data[base + i] = pre; pre += v[i]; }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// single block: exclusive scan of tile_sums in place
static __global__ void __launch_bounds__(kScanBlock) k_scan_top(uint32_t* tile_sums, size_t ntiles) {
  __shared__ uint32_t s_warp[33];
  uint32_t running = 0;
  for (size_t base = 0; base < ntiles; base += kScanBlock) {
    const size_t i = base + threadIdx.x;
    uint32_t v = i < ntiles ? tile_sums[i] : 0, total;
    uint32_t pre = block_excl_scan(v, s_warp, &total);
    if (i < ntiles) tile_sums[i] = running + pre;
    running += total;
  }
}
static __global__ void __launch_bounds__(kScanBlock) k_scan_add(uint32_t* data, size_t n, const uint32_t* tile_sums) {
  const uint32_t add = tile_sums[blockIdx.x];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) if (base + i < n) data[base + i] += add;
}

// ---- pass A: fixed-size runs of the sorted entry list ------------------------------------------------
// offsets has nbt + 1 entries (offsets[nbt] = number of entries E).  Thread t owns entries [t*T, (t+1)*T).
// MINB = resident blocks per SM the register allocation is capped for (4: 128 regs, 5: 96, 6: 80 + small spills)
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k_msm_accumulate(const uint32_t* __restrict__ offsets, uint32_t nbt, const uint32_t* __restrict__ entries,
                 const G1Aff* __restrict__ bases, uint32_t T, G1X* __restrict__ bucket_sums,
                 G1X* __restrict__ head, G1X* __restrict__ tail) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t E = offsets[nbt];
  const unsigned long long start64 = (unsigned long long)t * T;
  if (start64 >= E) return;
  const uint32_t start = (uint32_t)start64;
  const uint32_t end = (E - start < T) ? E : start + T;
  // bucket containing `start`: largest b with offsets[b] <= start (skipping empty buckets to the right)
  uint32_t lo = 0, hi = nbt - 1;
  while (lo < hi) {
    uint32_t mid = (lo + hi + 1) >> 1;
    if (__ldg(offsets + mid) <= start) lo = mid; else hi = mid - 1;
  }
  uint32_t b = lo;
  uint32_t next_off = __ldg(offsets + b + 1);
  uint32_t seg_first = start;
  bool begins_here = __ldg(offsets + b) == start;
  uint32_t e = __ldg(entries + start);
  G1Aff pt = g1aff_load(bases + (e & ~kSignBit));
  G1X acc = g1x_inf();
  // ONE flat loop: every iteration is a uniform mixed addition for all 32 lanes; bucket boundaries only add a
  // short divergent flush (the accumulator restarts from infinity, which g1x_madd treats as a copy).
  for (uint32_t p = start; p < end; p++) {
    if (p == next_off) {
      G1X* dst = begins_here ? bucket_sums + b : (seg_first == start ? head + t : tail + t);
      g1x_store(dst, acc);
      acc = g1x_inf();
      seg_first = p; begins_here = true;
      do { b++; next_off = __ldg(offsets + b + 1); } while (next_off <= p);
    }
    const G1Aff cur = pt;
    const bool neg = (e & kSignBit) != 0;
    if (p + 1 < end) { e = __ldg(entries + p + 1); pt = g1aff_load(bases + (e & ~kSignBit)); }
    g1x_madd(acc, cur, neg);
  }
  const bool complete = begins_here && next_off <= end;
  G1X* dst = complete ? bucket_sums + b : (seg_first == start ? head + t : tail + t);
  g1x_store(dst, acc);
}

// ---- pass B: per bucket, add the partials of the runs it spans -------------------------------------
constexpr uint32_t kBigSpan = 48;
JA_DEV G1X msm_piece(const G1X* head, const G1X* tail, uint32_t c, uint32_t c0, bool starts_on_run) {
  return (c == c0 && !starts_on_run) ? g1x_load(tail + c) : g1x_load(head + c);
}
static __global__ void __launch_bounds__(128)
k_msm_combine(const uint32_t* __restrict__ offsets, uint32_t nbt, uint32_t T, const G1X* __restrict__ head,
              const G1X* __restrict__ tail, G1X* __restrict__ bucket_sums, uint32_t* __restrict__ big_list,
              uint32_t* __restrict__ big_count) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbt) return;
  const uint32_t o0 = offsets[b], o1 = offsets[b + 1];
  if (o0 == o1) { g1x_store(bucket_sums + b, g1x_inf()); return; }
  const uint32_t c0 = o0 / T, c1 = (o1 - 1) / T;
  if (c0 == c1) return;                       // final value written by pass A
  if (c1 - c0 > kBigSpan) { big_list[atomicAdd(big_count, 1u)] = b; return; }
  const bool on_run = (o0 % T) == 0;
  G1X acc = msm_piece(head, tail, c0, c0, on_run);
  for (uint32_t c = c0 + 1; c <= c1; c++) g1x_add(acc, g1x_load(head + c));
  g1x_store(bucket_sums + b, acc);
}
// wide buckets (skewed digits, one-hot sums): one block per bucket, strided partial sums + shared-memory tree
static __global__ void __launch_bounds__(128)
k_msm_combine_big(const uint32_t* __restrict__ offsets, uint32_t T, const G1X* __restrict__ head,
                  const G1X* __restrict__ tail, G1X* __restrict__ bucket_sums, const uint32_t* __restrict__ big_list,
                  const uint32_t* __restrict__ big_count) {
  __shared__ G1X s_acc[128];
  const uint32_t nbig = *big_count;
  for (uint32_t i = blockIdx.x; i < nbig; i += gridDim.x) {
    const uint32_t b = big_list[i];
    const uint32_t o0 = offsets[b], o1 = offsets[b + 1];
    const uint32_t c0 = o0 / T, c1 = (o1 - 1) / T;
    const bool on_run = (o0 % T) == 0;
    G1X acc = g1x_inf();
    for (uint32_t c = c0 + threadIdx.x; c <= c1; c += blockDim.x) g1x_add(acc, msm_piece(head, tail, c, c0, on_run));
    s_acc[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
      if (threadIdx.x < s) { G1X a = s_acc[threadIdx.x]; g1x_add(a, s_acc[threadIdx.x + s]); s_acc[threadIdx.x] = a; }
      __syncthreads();
    }
    if (threadIdx.x == 0) g1x_store(bucket_sums + b, s_acc[0]);
    __syncthreads();
  }
}

// ---- bucket reduction: W = sum_b (b+1) * B_b per window ------------------------------------------------
struct MsmWindow { uint32_t bucket_base, nb, c, msm; };
constexpr uint32_t kSegBuckets = 8;       // buckets per thread: the running sums are ONE dependent chain of 2 x kSegBuckets full additions
constexpr uint32_t kSegSpan = 1024;       // segment partials summed per block of k_msm_part_sum (stage A)
// grid (ceil(max_segs/128), n_windows); seg_part[w * max_segs + seg]
static __global__ void __launch_bounds__(128)
k_msm_bucket_reduce(const MsmWindow* __restrict__ wins, const G1X* __restrict__ bucket_sums, uint32_t max_segs,
                    G1X* __restrict__ seg_part) {
  const MsmWindow w = wins[blockIdx.y];
  const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lo = seg * kSegBuckets;
  if (lo >= w.nb) return;
  const uint32_t hi = lo + kSegBuckets < w.nb ? lo + kSegBuckets : w.nb;
  G1X run = g1x_inf(), tot = g1x_inf();
  for (uint32_t b = hi; b-- > lo;) {
    g1x_add(run, g1x_load(bucket_sums + w.bucket_base + b));
    g1x_add(tot, run);
  }
  // sum_b (b+1) B_b = sum_b (b-lo+1) B_b + lo * sum_b B_b
  if (lo) g1x_add(tot, g1x_mul_small(run, lo));
  g1x_store(seg_part + (size_t)blockIdx.y * max_segs + seg, tot);
}
// Tree sum of partial points, two stages so that a window of 2^19 buckets is not summed by one block:
//   A  grid (ceil(max_in / span), n_windows): block (x, w) sums inputs [x span, (x+1) span) of window w -> out[w * out_stride + x]
//   B  grid (1, n_windows) with span >= the number of stage-A blocks -> the window sum
// `per` = buckets covered by one input element (kSegBuckets for the segment partials, kSegBuckets * span for stage A's output).
static __global__ void __launch_bounds__(128)
k_msm_part_sum(const MsmWindow* __restrict__ wins, const G1X* __restrict__ in, uint32_t in_stride, uint32_t per, uint32_t span,
               G1X* __restrict__ out, uint32_t out_stride) {
  __shared__ G1X s_acc[128];
  const MsmWindow w = wins[blockIdx.y];
  const uint32_t n_in = (w.nb + per - 1) / per;
  const uint32_t lo = blockIdx.x * span;
  if (lo >= n_in) return;
  const uint32_t hi = lo + span < n_in ? lo + span : n_in;
  G1X acc = g1x_inf();
  for (uint32_t s = lo + threadIdx.x; s < hi; s += blockDim.x) g1x_add(acc, g1x_load(in + (size_t)blockIdx.y * in_stride + s));
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  const uint32_t cnt = hi - lo;
  for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (threadIdx.x < s && threadIdx.x + s < cnt) { G1X a = s_acc[threadIdx.x]; g1x_add(a, s_acc[threadIdx.x + s]); s_acc[threadIdx.x] = a; }
    __syncthreads();
  }
  if (threadIdx.x == 0) g1x_store(out + (size_t)blockIdx.y * out_stride + blockIdx.x, s_acc[0]);
}

// ---- indexed point sums (one-hot commitments), dedicated path ---------------------------------------------------
// HyperKZG::commit_one_hot / batch_commit_one_hot (hyperkzg/mod.rs:520-596): C_j = sum_t G[idx_j[t]], no scalars, no
// buckets.  Blocks own kIdxBlock * kIdxRun consecutive entries of ONE list: every thread chains kIdxRun mixed
// additions (strided so the index loads coalesce), the block folds its 128 partial sums in shared memory, and a second
// launch folds the per-block partials of each list.  The XYZZ results go to the host, which normalises the whole
// batch with one inversion (fq_host.hpp).
// kIdxRun: the block's closing tree is 7 levels of FULL additions (~98 products on the critical path, most threads idle), as
// much as a chain of 8 mixed additions (80 products): at kIdxRun = 8 the kernel ran at 0.46 of the field-mul peak.  32 puts
// 320 products of useful chain against the same tree.
constexpr int kIdxBlock = 128, kIdxRun = 32;
struct IdxJob { const unsigned long long* idx; uint32_t n; uint32_t first_block; };

JA_DEV G1X block_point_sum(G1X acc, G1X* s_acc /* kIdxBlock */) {
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t s = kIdxBlock >> 1; s > 0; s >>= 1) {
    if (threadIdx.x < s) { G1X a = s_acc[threadIdx.x]; g1x_add(a, s_acc[threadIdx.x + s]); s_acc[threadIdx.x] = a; }
    __syncthreads();
  }
  return s_acc[0];
}

static __global__ void __launch_bounds__(kIdxBlock)
k_indexed_partial(const IdxJob* __restrict__ jobs, uint32_t njobs, const G1Aff* __restrict__ bases, G1X* __restrict__ partial) {
  __shared__ G1X s_acc[kIdxBlock];
  uint32_t lo = 0, hi = njobs - 1;
  while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (jobs[mid].first_block <= blockIdx.x) lo = mid; else hi = mid - 1; }
  const IdxJob job = jobs[lo];
  const uint32_t base = (blockIdx.x - job.first_block) * (kIdxBlock * kIdxRun);
  G1X acc = g1x_inf();
  uint32_t e = base + threadIdx.x;
  G1Aff pt;
  if (e < job.n) pt = g1aff_load(bases + __ldg(job.idx + e));
#pragma unroll 1
  for (int r = 0; r < kIdxRun; r++) {
    if (e >= job.n) break;
    const G1Aff cur = pt;
    e += kIdxBlock;
    if (r + 1 < kIdxRun && e < job.n) pt = g1aff_load(bases + __ldg(job.idx + e));    // prefetch the next base
    g1x_madd(acc, cur, false);
  }
  const G1X tot = block_point_sum(acc, s_acc);
  if (threadIdx.x == 0) g1x_store(partial + blockIdx.x, tot);
}
// one block per list: fold its per-block partials
static __global__ void __launch_bounds__(kIdxBlock)
k_indexed_final(const IdxJob* __restrict__ jobs, uint32_t njobs, const G1X* __restrict__ partial, uint32_t total_blocks,
                G1X* __restrict__ out) {
  __shared__ G1X s_acc[kIdxBlock];
  const IdxJob job = jobs[blockIdx.x];
  const uint32_t end = blockIdx.x + 1 < njobs ? jobs[blockIdx.x + 1].first_block : total_blocks;
  G1X acc = g1x_inf();
  for (uint32_t b = job.first_block + threadIdx.x; b < end; b += kIdxBlock) g1x_add(acc, g1x_load(partial + b));
  const G1X tot = block_point_sum(acc, s_acc);
  if (threadIdx.x == 0) g1x_store(out + blockIdx.x, tot);
}

// ---- final: Horner over windows + affine conversion; one thread per MSM -------------------------------
struct MsmResult { Fq x, y; uint32_t inf; uint32_t pad[3]; };   // 80 B
static __global__ void __launch_bounds__(32)
k_msm_final(const MsmDesc* __restrict__ descs, uint32_t count, const G1X* __restrict__ window_sums,
            MsmResult* __restrict__ out) {
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= count) return;
  const MsmDesc d = descs[m];
  G1X acc = g1x_inf();
  for (uint32_t w = d.fixed_stride ? d.sub : d.nwin; w-- > 0;) {
    if (!g1x_is_inf(acc)) for (uint32_t k = 0; k < d.c; k++) acc = g1x_dbl(acc);
    g1x_add(acc, g1x_load(window_sums + d.win_base + w));
  }
  MsmResult r;
  if (g1x_is_inf(acc)) { r.x = fp_zero<FqParams>(); r.y = fp_zero<FqParams>(); r.inf = 1; }
  else { G1Aff a = g1x_to_aff(acc); r.x = a.x; r.y = a.y; r.inf = 0; }
  r.pad[0] = r.pad[1] = r.pad[2] = 0;
  out[m] = r;
}

}  // namespace ja

// ---- SRS generation: g1_powers[i] = beta^i * g1  (SRS::setup, hyperkzg/kzg.rs:26-93; FixedBase::msm) ----------
// Fixed-base: table[k] = 2^k * g1 (affine, built once by k_srs_table), then every point is a sum of <= 254 table
// entries (mixed additions only, no doublings) followed by one inversion.
namespace ja {

static __global__ void k_srs_table(G1Aff g1, G1Aff* __restrict__ table /*254*/) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  G1X cur = g1x_from_aff(g1);
  for (int k = 0; k < 254; k++) {
    table[k] = g1x_to_aff(cur);
    cur = g1x_dbl(cur);
  }
}

JA_DEV Fr fr_pow_u32(const Fr& b, uint32_t e) {
  Fr r = fp_one<FrParams>();
  if (e == 0) return r;
  for (int bit = 31 - __clz(e); bit >= 0; bit--) {
    r = fp_sqr<FrParams>(r);
    if ((e >> bit) & 1) r = fp_mul<FrParams>(r, b);
  }
  return r;
}

// Fixed-base window table: table[w * n + i] = 2^(c w) * P_i, w < nwin <= 22.  One thread per point: (nwin - 1) x c doublings
// in XYZZ, then the multiples are normalised with one shared inversion (Montgomery's trick).
static __global__ void __launch_bounds__(128)
k_srs_window_table(const G1Aff* __restrict__ points, uint32_t n, G1Aff* __restrict__ table, int c, int nwin) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const G1Aff p = g1aff_load(points + i);
  fp_store(&table[i].x, p.x); fp_store(&table[i].y, p.y);
  G1X cur = g1x_from_aff(p);
  // pass 1: walk the doubling chain; the XYZZ multiples wait in their table slots (x <- X * ZZZ, y <- Y * ZZ) and the prefix
  // products of their denominators ZZ * ZZZ in local arrays
  Fq pre[21], den[21];
  Fq run = fp_one<FqParams>();
#pragma unroll 1
  for (int w = 1; w < nwin; w++) {
#pragma unroll 1
    for (int k = 0; k < c; k++) cur = g1x_dbl(cur);
    // X / ZZ and Y / ZZZ need 1 / (ZZ * ZZZ): x = X * ZZZ * inv, y = Y * ZZ * inv
    fp_store(&table[(size_t)w * n + i].x, fq_mul(cur.X, cur.ZZZ));
    fp_store(&table[(size_t)w * n + i].y, fq_mul(cur.Y, cur.ZZ));
    den[w - 1] = fq_mul(cur.ZZ, cur.ZZZ);
    pre[w - 1] = run;
    run = fq_mul(run, den[w - 1]);
  }
  Fq inv = fq_inv(run);
#pragma unroll 1
  for (int w = nwin - 1; w >= 1; w--) {
    const Fq di = fq_mul(inv, pre[w - 1]);
    inv = fq_mul(inv, den[w - 1]);
    G1Aff* dst = table + (size_t)w * n + i;
    const Fq x = fq_mul(fp_load(&dst->x), di), y = fq_mul(fp_load(&dst->y), di);
    fp_store(&dst->x, x); fp_store(&dst->y, y);
  }
}

static __global__ void __launch_bounds__(128)
k_srs_powers(const G1Aff* __restrict__ table, Fr beta, uint32_t n, G1Aff* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = fr_pow_u32(beta, i);
  const uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  Fr canon;
  fp_mont_rows<FrParams, 8>(canon.l, s.l, one);
  fp_final_sub<FrParams>(canon.l);
  G1X acc = g1x_inf();
#pragma unroll 1
  for (int w = 0; w < 8; w++) {
    uint32_t bits = canon.l[w];
    while (bits) {
      const int k = __ffs(bits) - 1;
      bits &= bits - 1;
      g1x_madd(acc, g1aff_load(table + 32 * w + k), false);
    }
  }
  // beta^i != 0 => acc is never the identity for a generator of the prime-order group
  G1Aff a = g1x_to_aff(acc);
  fp_store(&out[i].x, a.x); fp_store(&out[i].y, a.y);
}

}  // namespace ja
