// One-hot address batches resident on the device: the d chunk-address lists of a node (Option<u8/u16> per cycle,
// K = 2^log_k_chunk = 16) are uploaded ONCE as u32 (0xFFFFFFFF = None, 4 B per entry) and every consumer derives what
// it needs from them on the device:
//   ja_addr_commit      HyperKZG::batch_commit_one_hot (hyperkzg/mod.rs:558-596): C_i = sum_t G[k_i[t] * T + t]
//   ja_addr_gather      RaPolynomial materialisation (poly/ra_poly.rs:31-81): ra_i[t] = table_i[k_i[t]]
//   ja_addr_ra_evals    compute_ra_evals (subprotocols/shout.rs:549-598): G_i[k] = sum_{t: k_i[t] = k} eq(r_cycle, t)
// No CPU fallback.
#include "common.hpp"
#include "fq_host.hpp"
#include "msm_kernels.cuh"
#include "poly_kernels.cuh"

namespace ja {

// point sums straight from the address lists: entry t of a list selects base k*T + t.  One job per LIST (all lists of
// all nodes of a proof in one launch: the witness commitments do not depend on the transcript, prover.rs:72-87).
struct AddrJob { const uint32_t* k; uint32_t T; uint32_t first_block; uint32_t pad; };
static __global__ void __launch_bounds__(kIdxBlock)
k_addr_partial(const AddrJob* __restrict__ jobs, uint32_t njobs, const G1Aff* __restrict__ bases, G1X* __restrict__ partial) {
  __shared__ G1X s_acc[kIdxBlock];
  uint32_t lo = 0, hi = njobs - 1;
  while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (jobs[mid].first_block <= blockIdx.x) lo = mid; else hi = mid - 1; }
  const AddrJob job = jobs[lo];
  const uint32_t* __restrict__ k = job.k;
  const uint32_t T = job.T;
  const uint32_t base = (blockIdx.x - job.first_block) * (kIdxBlock * kIdxRun);
  G1X acc = g1x_inf();
  uint32_t e = base + threadIdx.x;
  uint32_t kk = e < T ? __ldg(k + e) : 0xffffffffu;
  G1Aff pt;
  if (kk != 0xffffffffu) pt = g1aff_load(bases + (size_t)kk * T + e);
#pragma unroll 1
  for (int r = 0; r < kIdxRun; r++) {
    if (e >= T) break;
    const G1Aff cur = pt;
    const bool have = kk != 0xffffffffu;
    e += kIdxBlock;
    if (r + 1 < kIdxRun && e < T) {
      kk = __ldg(k + e);
      if (kk != 0xffffffffu) pt = g1aff_load(bases + (size_t)kk * T + e);
    }
    if (have) g1x_madd(acc, cur, false);
  }
  const G1X tot = block_point_sum(acc, s_acc);
  if (threadIdx.x == 0) g1x_store(partial + blockIdx.x, tot);
}
static __global__ void __launch_bounds__(kIdxBlock)
k_addr_final(const AddrJob* __restrict__ jobs, uint32_t njobs, uint32_t total_blocks, const G1X* __restrict__ partial, G1X* __restrict__ out) {
  __shared__ G1X s_acc[kIdxBlock];
  const uint32_t first = jobs[blockIdx.x].first_block;
  const uint32_t end = blockIdx.x + 1 < njobs ? jobs[blockIdx.x + 1].first_block : total_blocks;
  G1X acc = g1x_inf();
  for (uint32_t b = first + threadIdx.x; b < end; b += kIdxBlock) g1x_add(acc, g1x_load(partial + b));
  const G1X tot = block_point_sum(acc, s_acc);
  if (threadIdx.x == 0) g1x_store(out + blockIdx.x, tot);
}

// ra_i[t] = table_i[k_i[t]] for all d lists in one launch (blockIdx.y = list)
struct GatherOut { Fr* p[kMaxProdPolys]; };
static __global__ void __launch_bounds__(kBlock)
k_addr_gather(const uint32_t* __restrict__ k_all, size_t T, const Fr* __restrict__ tables, uint32_t K, GatherOut out) {
  const uint32_t* __restrict__ k = k_all + (size_t)blockIdx.y * T;
  const Fr* __restrict__ tab = tables + (size_t)blockIdx.y * K;
  Fr* __restrict__ dst = out.p[blockIdx.y];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += stride) {
    const uint32_t kk = __ldg(k + t);
    fp_store(dst + t, kk == 0xffffffffu ? fp_zero<FrParams>() : fp_load(tab + kk));
  }
}

// G_i[k] partial sums.  Block = 16 bins x 16 entry lanes (K <= 16 per pass over bins); grid (tiles, d).
// Lane l of a block owns the CONTIGUOUS run of kRaTile / 16 entries [t0 + 64 l, t0 + 64 l + 64): its addresses arrive as
// 16 independent 128-bit loads issued up front (the 16 bin-threads of a lane read the same words: one L1 broadcast),
// and only the ~1/16 of the entries that hit this thread's bin load their eq value.  (The first version walked the
// entries one dependent 4-byte load at a time: 34 us per launch at T = 2^14, latency-bound.)
constexpr int kRaTile = 1024;      // entries per block
constexpr int kRaRun = kRaTile / 16;
static __global__ void __launch_bounds__(256)
k_ra_evals_partial(const uint32_t* __restrict__ k_all, size_t T, const Fr* __restrict__ eq, uint32_t K, uint32_t k_base,
                   Fr* __restrict__ partial /* [d][tiles][16] */) {
  __shared__ Fr s_bin[16][16];
  const uint32_t bin = threadIdx.x & 15, lane = threadIdx.x >> 4;
  const uint32_t want = k_base + bin;
  const uint32_t* __restrict__ k = k_all + (size_t)blockIdx.y * T;
  const size_t t0 = (size_t)blockIdx.x * kRaTile + (size_t)lane * kRaRun;
  Fr acc = fp_zero<FrParams>();
  if (t0 + kRaRun <= T && (T & 3) == 0) {
    uint4 kk[kRaRun / 4];
#pragma unroll
    for (int q = 0; q < kRaRun / 4; q++) kk[q] = __ldg(reinterpret_cast<const uint4*>(k + t0) + q);
#pragma unroll
    for (int q = 0; q < kRaRun / 4; q++) {
      const bool h0 = kk[q].x == want, h1 = kk[q].y == want, h2 = kk[q].z == want, h3 = kk[q].w == want;
      if (h0 | h1 | h2 | h3) {
        const Fr* e = eq + t0 + 4 * q;
        Fr v0 = fp_zero<FrParams>(), v1 = v0, v2 = v0, v3 = v0;
        if (h0) v0 = fp_load(e);
        if (h1) v1 = fp_load(e + 1);
        if (h2) v2 = fp_load(e + 2);
        if (h3) v3 = fp_load(e + 3);
        if (h0) acc = fp_add<FrParams>(acc, v0);
        if (h1) acc = fp_add<FrParams>(acc, v1);
        if (h2) acc = fp_add<FrParams>(acc, v2);
        if (h3) acc = fp_add<FrParams>(acc, v3);
      }
    }
  } else {
    for (size_t t = t0; t < t0 + kRaRun && t < T; t++)
      if (__ldg(k + t) == want) acc = fp_add<FrParams>(acc, fp_load(eq + t));
  }
  s_bin[lane][bin] = acc;
  __syncthreads();
  if (lane == 0) {
    Fr tot = s_bin[0][bin];
    for (int l = 1; l < 16; l++) tot = fp_add<FrParams>(tot, s_bin[l][bin]);
    fp_store(partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + bin, tot);
  }
}
static __global__ void __launch_bounds__(256)
k_ra_evals_final(const Fr* __restrict__ partial, uint32_t tiles, uint32_t d, uint32_t K, uint32_t k_base, Fr* __restrict__ out /* [d][K] */) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= d * 16) return;
  const uint32_t i = id >> 4, bin = id & 15;
  if (k_base + bin >= K) return;
  Fr tot = fp_zero<FrParams>();
  for (uint32_t b = 0; b < tiles; b++) tot = fp_add<FrParams>(tot, fp_load(partial + ((size_t)i * tiles + b) * 16 + bin));
  fp_store(out + (size_t)i * K + k_base + bin, tot);
}

// One-launch form for a single job whose results the host is waiting for (the RA checks of a node call compute_ra_evals
// right before their sumcheck): eq(r_cycle, t) is formed on the fly from the two half tables (eq[t] = hi[t >> bits_lo] *
// lo[t & mask], the same product k_eq_expand stores), the block that finishes the LAST tile of a list adds the list's tile
// partials, and the sums go to the host with store_tagged (no D2H copy, no stream synchronisation).  K <= 16.
static __global__ void __launch_bounds__(256)
k_ra_evals_fused(const uint32_t* __restrict__ k_all, size_t T, const Fr* __restrict__ eq_hi, const Fr* __restrict__ eq_lo, int bits_lo,
                 uint32_t K, Fr* __restrict__ partial /* [d][tiles][16] */, unsigned int* counters /* [d], zero on entry, reset on exit */,
                 Fr* out /* d x K: tagged host-mapped elements (tag != 0) or plain device elements (tag == 0) */, unsigned int tag) {
  __shared__ Fr s_eq[kRaTile];            // the tile's eq values: 4 products per thread, all lanes busy (forming them at the hits
  __shared__ Fr s_bin[16][16];            // would run the product once per ENTRY per warp: every entry hits exactly one bin lane)
  __shared__ bool s_last;
  const uint32_t bin = threadIdx.x & 15, lane = threadIdx.x >> 4;
  const uint32_t* __restrict__ k = k_all + (size_t)blockIdx.y * T;
  const size_t tile0 = (size_t)blockIdx.x * kRaTile;
  const size_t t0 = tile0 + (size_t)lane * kRaRun;
  const size_t mask_lo = (size_t(1) << bits_lo) - 1;
  for (uint32_t i = threadIdx.x; i < (uint32_t)kRaTile; i += blockDim.x) {
    const size_t t = tile0 + i;
    if (t < T) s_eq[i] = fp_mul<FrParams>(fp_load(eq_hi + (t >> bits_lo)), fp_load(eq_lo + (t & mask_lo)));
  }
  __syncthreads();
  Fr acc = fp_zero<FrParams>();
  for (size_t t = t0; t < t0 + kRaRun && t < T; t += 4) {
    uint32_t kk[4];
    if (t + 4 <= T && (T & 3) == 0) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(k + t));
      kk[0] = v.x; kk[1] = v.y; kk[2] = v.z; kk[3] = v.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; q++) kk[q] = t + q < T ? __ldg(k + t + q) : 0xffffffffu;
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (kk[q] == bin) acc = fp_add<FrParams>(acc, s_eq[t + q - tile0]);
  }
  s_bin[lane][bin] = acc;
  __syncthreads();
  const uint32_t tiles = gridDim.x;
  if (lane == 0) {
    Fr tot = s_bin[0][bin];
    for (int l = 1; l < 16; l++) tot = fp_add<FrParams>(tot, s_bin[l][bin]);
    if (tiles == 1) {
      if (bin < K) { if (tag) store_tagged(out, (int)(blockIdx.y * K + bin), tot, tag); else fp_store(out + (size_t)blockIdx.y * K + bin, tot); }
    }
    else fp_store(partial + ((size_t)blockIdx.y * tiles + blockIdx.x) * 16 + bin, tot);
  }
  if (tiles == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicInc(counters + blockIdx.y, tiles - 1) == tiles - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  acc = fp_zero<FrParams>();
  for (uint32_t b = lane; b < tiles; b += 16) {
    const uint4* q = reinterpret_cast<const uint4*>(partial + ((size_t)blockIdx.y * tiles + b) * 16 + bin);
    const uint4 lo = __ldcg(q), hi = __ldcg(q + 1);
    Fr v;
    v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w; v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
    acc = fp_add<FrParams>(acc, v);
  }
  s_bin[lane][bin] = acc;
  __syncthreads();
  if (lane == 0 && bin < K) {
    Fr tot = s_bin[0][bin];
    for (int l = 1; l < 16; l++) tot = fp_add<FrParams>(tot, s_bin[l][bin]);
    if (tag) store_tagged(out, (int)(blockIdx.y * K + bin), tot, tag); else fp_store(out + (size_t)blockIdx.y * K + bin, tot);
  }
}

// build_materialized_rlc, sparse half (poly/rlc_polynomial.rs:59-74): joint[k_i[t] * T + t] += coeff_i for the d one-hot
// polynomials of an address batch.  Thread t owns column t (all its targets are congruent to t mod T), so the d updates
// of a column are sequential in one thread and no atomics are needed; different batches are separate launches.
static __global__ void __launch_bounds__(kBlock)
k_rlc_add_onehot(const uint32_t* __restrict__ k_all, size_t T, uint32_t d, const Fr* __restrict__ coeffs, Fr* __restrict__ joint) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += stride) {
    for (uint32_t i = 0; i < d; i++) {
      const uint32_t kk = __ldg(k_all + (size_t)i * T + t);
      if (kk == 0xffffffffu) continue;
      Fr* dst = joint + (size_t)kk * T + t;
      Fr cur;                                  // plain loads: the same thread may have just written this element
      const uint4* q = reinterpret_cast<const uint4*>(dst);
      const uint4 lo = q[0], hi = q[1];
      cur.l[0] = lo.x; cur.l[1] = lo.y; cur.l[2] = lo.z; cur.l[3] = lo.w; cur.l[4] = hi.x; cur.l[5] = hi.y; cur.l[6] = hi.z; cur.l[7] = hi.w;
      fp_store(dst, fp_add<FrParams>(cur, fp_load(coeffs + i)));
    }
  }
}
// The same for d <= 16 lists with a group of 16 lanes per column: lane i owns list i; lanes of a column that hit the same row
// (equal k) elect a leader, which adds their coefficients and performs ONE read-modify-write.  All read-modify-writes of a
// column go to distinct rows and are independent (one memory round trip instead of d dependent ones: 15 -> ~4 us per launch
// at T = 2^14), and the coefficients travel in the kernel parameters (no staged copy per call).
struct RlcCoeffs { Fr c[16]; };
static __global__ void __launch_bounds__(kBlock)
k_rlc_add_onehot_lanes(const uint32_t* __restrict__ k_all, size_t T, uint32_t d, const __grid_constant__ RlcCoeffs co, Fr* __restrict__ joint) {
  const uint32_t lane = threadIdx.x & 31, li = threadIdx.x & 15;
  const size_t col = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  uint32_t kk = 0xffffffffu;
  if (col < T && li < d) kk = __ldg(k_all + (size_t)li * T + col);
  const bool act = kk != 0xffffffffu;
  // match key: the row within this column's half-warp; idle lanes get keys that match nothing
  const uint32_t key = act ? (kk | ((lane & 16u) << 16)) : (0x80000000u | lane);
  const uint32_t peers = __match_any_sync(0xffffffffu, key);
  const bool leader = act && (uint32_t)(__ffs(peers) - 1) == lane;
  Fr mine = fp_zero<FrParams>();
  if (act) mine = co.c[li];
  Fr acc = mine;
#pragma unroll 1
  for (uint32_t o = 1; o < 16; o++) {
    const uint32_t src = (lane & 16u) | ((li + o) & 15u);
    Fr v;
#pragma unroll
    for (int q = 0; q < 8; q++) v.l[q] = __shfl_sync(0xffffffffu, mine.l[q], src);
    if (leader && ((peers >> src) & 1u)) acc = fp_add<FrParams>(acc, v);
  }
  if (leader) {
    Fr* dst = joint + (size_t)kk * T + col;
    fp_store(dst, fp_add<FrParams>(fp_load(dst), acc));
  }
}
// dense half (:42-57): joint[i] += coeff * poly[i] for i < len
static __global__ void __launch_bounds__(kBlock)
k_rlc_add_dense(const Fr* __restrict__ poly, size_t len, Fr coeff, Fr* __restrict__ joint) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    Fr cur;
    const uint4* q = reinterpret_cast<const uint4*>(joint + i);
    const uint4 lo = q[0], hi = q[1];
    cur.l[0] = lo.x; cur.l[1] = lo.y; cur.l[2] = lo.z; cur.l[3] = lo.w; cur.l[4] = hi.x; cur.l[5] = hi.y; cur.l[6] = hi.z; cur.l[7] = hi.w;
    fp_store(joint + i, fp_add<FrParams>(cur, fp_mul<FrParams>(coeff, fp_load(poly + i))));
  }
}

// address validation on the device (every entry must be None = 0xFFFFFFFF or < K): flags[slot] |= 1 on a violation.
// The host-side scan it replaces read the whole index array a second time on one core (1.1 GB for a GPT-2-shaped proof).
static __global__ void __launch_bounds__(kBlock)
k_addr_validate(const uint32_t* __restrict__ k, size_t n, uint32_t K, unsigned int* flags, unsigned int slot) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned int bad = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t v = k[i];
    bad |= (v != 0xffffffffu) & (v >= K);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flags + slot, 1u);
}

}  // namespace ja

int32_t eq_evals_device_pub(ja_ctx* c, const uint64_t* r, size_t m, Fr* out);   // capi.cu
int32_t eq_halves_device_pub(ja_ctx* c, const uint64_t* r, size_t m, Fr** lv_hi, Fr** lv_lo, const Fr** hi, const Fr** lo, int* bits_lo);   // capi.cu

extern "C" {

// Upload n address batches back to back: every copy and its validation kernel are enqueued, then ONE synchronisation
// (the source buffers may be reused when the call returns).  A proof uploads all of its one-hot index arrays at once
// (commit_witness_polynomials runs before the IOP), so this replaces one blocking copy + host scan per node.
int32_t ja_addr_upload_many(ja_ctx* c, const uint32_t* const* ks, const size_t* ds, const size_t* Ts, const size_t* Ks, size_t n,
                            ja_addr** outs) {
  JA_REQUIRE(c && ks && ds && Ts && Ks && outs && n > 0, "ja_addr_upload: null or empty argument");
  JA_REQUIRE(n * sizeof(unsigned int) <= kPinnedBytes, "ja_addr_upload: too many batches in one call");
  for (size_t b = 0; b < n; b++) {
    JA_REQUIRE(ks[b] && ds[b] > 0 && Ts[b] > 0 && Ks[b] > 0, "ja_addr_upload: null or empty argument");
    JA_REQUIRE(ds[b] <= (size_t)kMaxProdPolys, "ja_addr_upload: at most 32 lists per batch");
    JA_REQUIRE(Ts[b] < (size_t(1) << 31) && Ks[b] <= 65536, "ja_addr_upload: T or K too large");
  }
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  unsigned int* d_flags = nullptr;
  int32_t st = dev_alloc(c, n * sizeof(unsigned int), (void**)&d_flags);
  if (st) return st;
  JA_CUDA(cudaMemsetAsync(d_flags, 0, n * sizeof(unsigned int), c->stream));
  for (size_t b = 0; b < n; b++) outs[b] = nullptr;
  for (size_t b = 0; b < n && !st; b++) {
    ja_addr* a = new ja_addr();
    a->d = ds[b]; a->T = Ts[b]; a->K = Ks[b];
    outs[b] = a;
    const size_t cnt = ds[b] * Ts[b];
    if ((st = dev_alloc(c, cnt * sizeof(uint32_t), (void**)&a->d_k))) break;
    if (cudaMemcpyAsync(a->d_k, ks[b], cnt * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = fail(JA_ERR_CUDA, "ja_addr_upload: copy failed"); break; }
    JA_LAUNCH(c, KC_CONVERT, k_addr_validate<<<grid_for(cnt), kBlock, 0, c->stream>>>(a->d_k, cnt, (uint32_t)Ks[b], d_flags, (unsigned int)b));
  }
  unsigned int* h_flags = reinterpret_cast<unsigned int*>(c->h_pinned);
  if (!st && cudaMemcpyAsync(h_flags, d_flags, n * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) st = fail(JA_ERR_CUDA, "ja_addr_upload: copy failed");
  const cudaError_t se = cudaStreamSynchronize(c->stream);
  if (!st && se != cudaSuccess) st = fail(JA_ERR_CUDA, std::string("ja_addr_upload: ") + cudaGetErrorString(se));
  if (!st) for (size_t b = 0; b < n; b++) if (h_flags[b]) { st = fail(JA_ERR_INVALID, "ja_addr_upload: address outside [0, K)"); break; }
  dev_free(c, d_flags);
  if (st) {
    for (size_t b = 0; b < n; b++) if (outs[b]) { dev_free(c, outs[b]->d_k); delete outs[b]; outs[b] = nullptr; }
    return st;
  }
  return JA_OK;
}

int32_t ja_addr_upload(ja_ctx* c, const uint32_t* k, size_t d, size_t T, size_t K, ja_addr** out) {
  JA_REQUIRE(out, "ja_addr_upload: null or empty argument");
  return ja_addr_upload_many(c, &k, &d, &T, &K, 1, out);
}

void ja_addr_free(ja_ctx* c, ja_addr* a) {
  if (!c || !a) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  dev_free(c, a->d_k);
  delete a;
}

int32_t ja_addr_commit_many(ja_ctx* c, const ja_srs* srs, const ja_addr* const* batches, size_t n_batches, uint64_t* out_xy,
                            int32_t* is_inf) {
  JA_REQUIRE(c && srs && batches && n_batches && out_xy, "ja_addr_commit: null argument");
  std::vector<AddrJob> jobs;
  uint64_t blocks = 0;
  for (size_t b = 0; b < n_batches; b++) {
    const ja_addr* a = batches[b];
    JA_REQUIRE(a, "ja_addr_commit: null batch");
    if (a->K * a->T > srs->n)
      return fail(JA_ERR_KEY_LENGTH, "KeyLengthError: SRS has " + std::to_string(srs->n) + " powers, one-hot polynomial needs " +
                                         std::to_string(a->K * a->T));
    const uint32_t bpl = (uint32_t)((a->T + kIdxBlock * kIdxRun - 1) / (kIdxBlock * kIdxRun));
    for (size_t i = 0; i < a->d; i++) { jobs.push_back(AddrJob{a->d_k + i * a->T, (uint32_t)a->T, (uint32_t)blocks, 0}); blocks += bpl; }
    JA_REQUIRE(blocks < (1ull << 31), "ja_addr_commit: batch too large");
  }
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  // One proof on several GPUs (ja_comm_init + ja_set_msm_shard): the lists are dealt round-robin to the ranks - the commitments
  // of a proof are independent of each other and of the transcript - and ONE all-gather makes every rank hold all of them.
  const bool dealt = c->comm && c->msm_shard_count > 1 && c->comm_world == c->msm_shard_count && jobs.size() >= c->comm_world;
  const size_t total_lists = jobs.size();
  std::vector<size_t> owner_slot;               // list i -> its position among this rank's lists
  if (dealt) {
    std::vector<AddrJob> mine;
    uint64_t nb = 0;
    for (size_t i = 0; i < jobs.size(); i++) {
      if (i % c->comm_world != c->comm_rank) continue;
      AddrJob j = jobs[i];
      const uint32_t bpl = (uint32_t)((j.T + kIdxBlock * kIdxRun - 1) / (kIdxBlock * kIdxRun));
      j.first_block = (uint32_t)nb; nb += bpl;
      mine.push_back(j);
    }
    jobs.swap(mine); blocks = nb;
  }
  const size_t count = jobs.size();
  auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t o_part = align(sizeof(AddrJob) * count), o_out = align(o_part + sizeof(G1X) * blocks);
  char* ws = nullptr;
  int32_t st = dev_alloc(c, o_out + sizeof(G1X) * count, (void**)&ws);
  if (st) return st;
  AddrJob* d_jobs = (AddrJob*)ws;
  G1X* d_part = (G1X*)(ws + o_part);
  G1X* d_out = (G1X*)(ws + o_out);
  JA_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(AddrJob) * count, cudaMemcpyHostToDevice, c->stream));
  JA_LAUNCH(c, KC_ONEHOT_SUM, k_addr_partial<<<(unsigned)blocks, kIdxBlock, 0, c->stream>>>(d_jobs, (uint32_t)count, srs->points, d_part));
  JA_LAUNCH(c, KC_ONEHOT_SUM, k_addr_final<<<(unsigned)count, kIdxBlock, 0, c->stream>>>(d_jobs, (uint32_t)count, (uint32_t)blocks, d_part, d_out));
  JA_CUDA(cudaGetLastError());
  std::vector<host::G1XH> sums(count);
  JA_CUDA(cudaMemcpyAsync(sums.data(), d_out, sizeof(G1X) * count, cudaMemcpyDeviceToHost, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));
  dev_free(c, ws);
  std::vector<int32_t> inf(count);
  if (!dealt) {
    host::xyzz_batch_to_affine(sums.data(), count, out_xy, inf.data());
    if (is_inf) memcpy(is_inf, inf.data(), sizeof(int32_t) * count);
    return JA_OK;
  }
  const size_t world = c->comm_world, per_rank = (total_lists + world - 1) / world;
  std::vector<uint64_t> mine(per_rank * 9, 0), all(per_rank * 9 * world);
  {
    std::vector<uint64_t> xy(count * 8);
    host::xyzz_batch_to_affine(sums.data(), count, xy.data(), inf.data());
    for (size_t j = 0; j < count; j++) { memcpy(&mine[9 * j], &xy[8 * j], 64); mine[9 * j + 8] = (uint64_t)inf[j]; }
  }
  if ((st = comm_allgather(c, mine.data(), mine.size() * 8, all.data()))) return st;
  for (size_t i = 0; i < total_lists; i++) {
    const uint64_t* src = &all[((i % world) * per_rank + i / world) * 9];
    memcpy(out_xy + 8 * i, src, 64);
    if (is_inf) is_inf[i] = (int32_t)src[8];
  }
  return JA_OK;
}

int32_t ja_addr_commit(ja_ctx* c, const ja_srs* srs, const ja_addr* a, uint64_t* out_xy, int32_t* is_inf) {
  const ja_addr* arr[1] = {a};
  return ja_addr_commit_many(c, srs, arr, 1, out_xy, is_inf);
}

int32_t ja_addr_gather(ja_ctx* c, const ja_addr* a, const uint64_t* tables, ja_poly** out_polys) {
  JA_REQUIRE(c && a && tables && out_polys, "ja_addr_gather: null argument");
  JA_REQUIRE(is_pow2(a->T), "ja_addr_gather: T must be a power of two");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  const size_t tab_bytes = a->d * a->K * sizeof(Fr);
  JA_REQUIRE(tab_bytes <= kRingBytes / 4, "ja_addr_gather: tables too large for the staging ring");
  Fr* d_tab = nullptr;
  int32_t st = dev_alloc(c, tab_bytes, (void**)&d_tab);
  if (st) return st;
  if ((st = stage_h2d(c, d_tab, tables, tab_bytes))) return st;
  GatherOut go;
  for (size_t i = 0; i < a->d; i++) {
    if ((st = ja_poly_alloc(c, a->T, &out_polys[i]))) return st;
    go.p[i] = out_polys[i]->buf[0];
  }
  unsigned gx = grid_for(a->T);
  if (gx > (unsigned)kSMs * 2) gx = kSMs * 2;
  JA_LAUNCH(c, KC_CONVERT, k_addr_gather<<<dim3(gx, (unsigned)a->d), kBlock, 0, c->stream>>>(a->d_k, a->T, d_tab, (uint32_t)a->K, go));
  JA_CUDA(cudaGetLastError());
  dev_free(c, d_tab);                          // stream-ordered reuse (one stream per context)
  return JA_OK;
}

// workspace of one compute_ra_evals job (Fr elements): eq table + per-tile partial bins
static size_t ra_evals_ws(const ja_addr* a) {
  const size_t tiles = (a->T + kRaTile - 1) / kRaTile;
  return a->T + a->d * tiles * 16;
}
// enqueue one job on the context's stream; results (d x K Fr) go to d_out
static int32_t ra_evals_enqueue(ja_ctx* c, const ja_addr* a, const uint64_t* r_cycle, size_t log_t, Fr* ws, Fr* d_out) {
  const uint32_t tiles = (uint32_t)((a->T + kRaTile - 1) / kRaTile);
  Fr *d_eq = ws, *d_part = ws + a->T;
  int32_t st = eq_evals_device_pub(c, r_cycle, log_t, d_eq);
  if (st) return st;
  for (uint32_t k_base = 0; k_base < a->K; k_base += 16) {
    JA_LAUNCH(c, KC_SCATTER, k_ra_evals_partial<<<dim3(tiles, (unsigned)a->d), 256, 0, c->stream>>>(a->d_k, a->T, d_eq, (uint32_t)a->K, k_base, d_part));
    JA_LAUNCH(c, KC_SCATTER, k_ra_evals_final<<<(unsigned)((a->d * 16 + 255) / 256), 256, 0, c->stream>>>(d_part, tiles, (uint32_t)a->d, (uint32_t)a->K, k_base, d_out));
  }
  JA_CUDA(cudaGetLastError());
  return JA_OK;
}

int32_t ja_addr_ra_evals(ja_ctx* c, const ja_addr* a, const uint64_t* r_cycle, size_t log_t, uint64_t* out_G) {
  JA_REQUIRE(c && a && r_cycle && out_G, "ja_addr_ra_evals: null argument");
  JA_REQUIRE((size_t(1) << log_t) == a->T, "ja_addr_ra_evals: r_cycle length does not match T");
  JA_REQUIRE(a->d * a->K * sizeof(Fr) <= kPinnedBytes, "ja_addr_ra_evals: result too large for the staging buffer");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  Fr* ws = nullptr;
  const size_t n_out = a->d * a->K;
  if (a->K <= 16 && log_t <= 32 && n_out * 48 <= kRowSeqOffset && getenv("JA_NO_RA_FUSED") == nullptr) {
    // two launches (half tables, scatter + per-list reduction) and a tagged publication through the mapped value buffer
    const uint32_t tiles = (uint32_t)((a->T + kRaTile - 1) / kRaTile);
    Fr *lv_hi = nullptr, *lv_lo = nullptr;
    const Fr *hi = nullptr, *lo = nullptr;
    int bits_lo = 0;
    int32_t st = eq_halves_device_pub(c, r_cycle, log_t, &lv_hi, &lv_lo, &hi, &lo, &bits_lo);
    if (st) return st;
    unsigned int* d_ctr = nullptr;
    if ((st = dev_alloc(c, a->d * tiles * 16 * sizeof(Fr) + a->d * sizeof(unsigned int), (void**)&ws))) return st;
    d_ctr = reinterpret_cast<unsigned int*>(ws + a->d * tiles * 16);
    JA_CUDA(cudaMemsetAsync(d_ctr, 0, a->d * sizeof(unsigned int), c->stream));
    const uint32_t tag = next_tag(c);
    JA_LAUNCH(c, KC_SCATTER, k_ra_evals_fused<<<dim3(tiles, (unsigned)a->d), 256, 0, c->stream>>>(a->d_k, a->T, hi, lo, bits_lo, (uint32_t)a->K, ws, d_ctr,
                                                                                             reinterpret_cast<Fr*>(c->d_rowvals), tag));
    JA_CUDA(cudaGetLastError());
    st = wait_tagged(c, c->h_rowvals, tag, n_out, out_G, "compute_ra_evals kernel");
    dev_free(c, ws); dev_free(c, lv_hi); dev_free(c, lv_lo);
    return st;
  }
  int32_t st = dev_alloc(c, (ra_evals_ws(a) + n_out) * sizeof(Fr), (void**)&ws);
  if (st) return st;
  Fr* d_out = ws + ra_evals_ws(a);
  if ((st = ra_evals_enqueue(c, a, r_cycle, log_t, ws, d_out))) return st;
  JA_CUDA(cudaMemcpyAsync(c->h_pinned, d_out, n_out * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(out_G, c->h_pinned, n_out * sizeof(Fr));
  dev_free(c, ws);
  return JA_OK;
}

// compute_ra_evals for MANY address batches whose points are all known (the opening reduction initialises every one-hot
// opening at once, opening_reduction.rs:532-571): the jobs are enqueued back to back and synchronised ONCE
// (218 separate calls cost ~77 us each at GPT-2 size, most of it the per-call synchronisation and copy).
// out_G[j] receives d_j x K_j Fr.
int32_t ja_addr_ra_evals_many(ja_ctx* c, const ja_addr* const* addrs, const uint64_t* const* r_cycles, const size_t* log_ts, size_t n,
                              uint64_t* const* out_G) {
  JA_REQUIRE(c && addrs && r_cycles && log_ts && out_G, "ja_addr_ra_evals_many: null argument");
  if (n == 0) return JA_OK;
  size_t ws_total = 0, out_total = 0;
  for (size_t j = 0; j < n; j++) {
    JA_REQUIRE(addrs[j] && r_cycles[j] && out_G[j], "ja_addr_ra_evals_many: null job");
    JA_REQUIRE((size_t(1) << log_ts[j]) == addrs[j]->T, "ja_addr_ra_evals_many: r_cycle length does not match T");
    ws_total += ra_evals_ws(addrs[j]);
    out_total += addrs[j]->d * addrs[j]->K;
  }
  JA_REQUIRE(out_total * sizeof(Fr) <= kPinnedBytes, "ja_addr_ra_evals_many: results too large for the staging buffer");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  Fr* ws = nullptr;
  int32_t st = dev_alloc(c, (ws_total + out_total) * sizeof(Fr), (void**)&ws);
  if (st) return st;
  Fr* d_all = ws + ws_total;                     // every job's d x K results, contiguous: ONE copy back
  size_t off = 0, ooff = 0, staged = 0;
  size_t d_total = 0;
  for (size_t j = 0; j < n; j++) d_total += addrs[j]->d;
  unsigned int* d_ctr = nullptr;                 // per-list tile counters of the one-kernel form
  if ((st = dev_alloc(c, d_total * sizeof(unsigned int), (void**)&d_ctr))) return st;
  JA_CUDA(cudaMemsetAsync(d_ctr, 0, d_total * sizeof(unsigned int), c->stream));
  const bool fused_ok = getenv("JA_NO_RA_FUSED") == nullptr;
  size_t coff = 0;
  for (size_t j = 0; j < n; j++) {
    if (fused_ok && addrs[j]->K <= 16 && log_ts[j] <= 32) {
      // half eq tables + ONE scatter kernel per job (k_ra_evals_fused, plain device output)
      const ja_addr* a = addrs[j];
      const uint32_t tiles = (uint32_t)((a->T + kRaTile - 1) / kRaTile);
      Fr *lv_hi = nullptr, *lv_lo = nullptr;
      const Fr *hi = nullptr, *lo = nullptr;
      int bits_lo = 0;
      if ((st = eq_halves_device_pub(c, r_cycles[j], log_ts[j], &lv_hi, &lv_lo, &hi, &lo, &bits_lo))) return st;
      JA_LAUNCH(c, KC_SCATTER, k_ra_evals_fused<<<dim3(tiles, (unsigned)a->d), 256, 0, c->stream>>>(a->d_k, a->T, hi, lo, bits_lo, (uint32_t)a->K, ws + off + a->T,
                                                                                               d_ctr + coff, d_all + ooff, 0u));
      JA_CUDA(cudaGetLastError());
      dev_free(c, lv_hi); dev_free(c, lv_lo);
      off += ra_evals_ws(a); ooff += a->d * a->K; coff += a->d;
      continue;
    }
    coff += addrs[j]->d;
    if (staged + 64 * 32 > kRingBytes / 2) { JA_CUDA(cudaStreamSynchronize(c->stream)); staged = 0; }   // never lap the upload ring
    if ((st = ra_evals_enqueue(c, addrs[j], r_cycles[j], log_ts[j], ws + off, d_all + ooff))) return st;
    off += ra_evals_ws(addrs[j]);
    ooff += addrs[j]->d * addrs[j]->K;
    staged += 64 * 32;
  }
  JA_CUDA(cudaMemcpyAsync(c->h_pinned, d_all, out_total * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));
  ooff = 0;
  for (size_t j = 0; j < n; j++) {
    const size_t cnt = addrs[j]->d * addrs[j]->K;
    memcpy(out_G[j], c->h_pinned + 4 * ooff, cnt * sizeof(Fr));
    ooff += cnt;
  }
  dev_free(c, ws);
  dev_free(c, d_ctr);
  return JA_OK;
}

int32_t ja_poly_zeros(ja_ctx* c, size_t n, ja_poly** out) {
  int32_t st = ja_poly_alloc(c, n, out);
  if (st) return st;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaMemsetAsync((*out)->buf[0], 0, n * sizeof(Fr), c->stream));
  return JA_OK;
}

int32_t ja_rlc_add_onehot(ja_ctx* c, ja_poly* joint, const ja_addr* a, const uint64_t* coeffs) {
  JA_REQUIRE(c && joint && a && coeffs, "ja_rlc_add_onehot: null argument");
  JA_REQUIRE(a->K * a->T <= joint->len, "ja_rlc_add_onehot: joint polynomial shorter than K * T");
  JA_REQUIRE(a->d * sizeof(Fr) <= kPinnedBytes, "ja_rlc_add_onehot: too many coefficients");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  Fr* d_co = nullptr;
  int32_t st = dev_alloc(c, a->d * sizeof(Fr), (void**)&d_co);
  if (st) return st;
  if (a->d <= 16 && getenv("JA_NO_RLC_LANES") == nullptr) {
    dev_free(c, d_co);
    RlcCoeffs co;
    memset(&co, 0, sizeof(co));
    memcpy(co.c, coeffs, a->d * sizeof(Fr));
    const size_t threads = a->T * 16;
    JA_LAUNCH(c, KC_SCATTER, k_rlc_add_onehot_lanes<<<(unsigned)((threads + kBlock - 1) / kBlock), kBlock, 0, c->stream>>>(a->d_k, a->T, (uint32_t)a->d, co, joint->data()));
    JA_CUDA(cudaGetLastError());
    return JA_OK;
  }
  if ((st = stage_h2d(c, d_co, coeffs, a->d * sizeof(Fr)))) return st;
  unsigned gx = grid_for(a->T);
  JA_LAUNCH(c, KC_SCATTER, k_rlc_add_onehot<<<gx, kBlock, 0, c->stream>>>(a->d_k, a->T, (uint32_t)a->d, d_co, joint->data()));
  JA_CUDA(cudaGetLastError());
  dev_free(c, d_co);
  return JA_OK;
}

int32_t ja_rlc_add_dense(ja_ctx* c, ja_poly* joint, const ja_poly* poly, const uint64_t coeff[4]) {
  JA_REQUIRE(c && joint && poly && coeff, "ja_rlc_add_dense: null argument");
  JA_REQUIRE(poly->len <= joint->len, "ja_rlc_add_dense: joint polynomial shorter than the summand");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  Fr co; memcpy(co.l, coeff, 32);
  JA_LAUNCH(c, KC_SCATTER, k_rlc_add_dense<<<grid_for(poly->len), kBlock, 0, c->stream>>>(poly->data(), poly->len, co, joint->data()));
  JA_CUDA(cudaGetLastError());
  return JA_OK;
}

size_t ja_addr_len(const ja_addr* a) { return a ? a->T : 0; }
size_t ja_addr_count(const ja_addr* a) { return a ? a->d : 0; }

}  // extern "C"
