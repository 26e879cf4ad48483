// Prefix-suffix Shout (joltworks/src/subprotocols/ps_shout/mod.rs): the T-sized passes of the read-raf sumcheck over a 2^LOG_K
// entry table (LOG_K = 64 for the saturating-clamp lookups, jolt-atlas-core/src/onnx_proof/clamp_lookups/mod.rs:57).  The
// protocol runs LOG_K address rounds in NUM_PHASES = 8 phases over m = 2^(LOG_K / 8) = 256-entry suffix polynomials; the rounds
// themselves are O(m) host work on tables this small (prefix MLEs with checkpoints: joltworks/src/lookup_tables/, unchanged on
// the Rust side).  What scales with T - and what this file moves to the device - is the work at every phase boundary:
//   init_phase (mod.rs:269-303)          u_evals[j] *= v[phase-1][k_bound(j)]                       (T products)
//   init_suffix_polys (mod.rs:305-335)   Q_s[y] += u_evals[j] * suffix_s(suffix_bits(k_j)), y = prefix_bits(k_j) & (m - 1),
//   RafProverState::init_Q (poly/prefix_suffix.rs:294-351)   the same scatter for the raf suffixes (One, Identity)
//   init_log_t_rounds (mod.rs:420-446)   ra[j] = prod_phase v[phase][k_bound(j, phase)]             (T x 8 gathers, 7 products)
// Scatter by an 8-bit key without atomics on field elements: a block = m key bins (one thread each) scans a tile of entries
// staged in shared memory; every entry hits exactly one bin, whose thread adds u * t for every suffix as a 320-bit integer
// (t < 2^32: acc320_mad; the 64-bit Identity suffix goes through one Montgomery product) - one reduction per bin, tile and
// suffix.  Deterministic: the columns are integer sums, exact in any order.
#include "common.hpp"
#include "poly_kernels.cuh"
#include "sumcheck_host.hpp"
#include "transcript_host.hpp"

namespace {

constexpr int kPsMaxSuffixes = 8;

// suffix MLEs of the clamp-table family (joltworks/src/lookup_tables/suffixes/{higher_all_zero,hzero_mul_lword,hone_mul_lword,one}.rs)
// and of the identity polynomial (poly/identity_poly.rs: the suffix value itself), in closed form: `bits` = the low `len` bits of
// the XLEN-bit index; the "higher" bits are those of significance >= 2^bound.
__host__ JA_DEV unsigned long long suffix_mle(uint32_t kind, unsigned long long bits, uint32_t len, uint32_t bound) {
  const unsigned long long low_mask = bound >= 64 ? ~0ull : ((1ull << bound) - 1);
  const unsigned long long low = bits & low_mask;
  const unsigned long long hi = (len > bound) ? (bits >> bound) : 0ull;
  const uint32_t hi_len = len > bound ? len - bound : 0;
  switch (kind) {
    case JA_SUF_ONE: return 1ull;
    case JA_SUF_HIGHER_ALL_ZERO: return hi == 0 ? 1ull : 0ull;
    case JA_SUF_HZERO_MUL_LWORD: return hi == 0 ? low : 0ull;
    case JA_SUF_HONE_MUL_LWORD: {
      const unsigned long long ones = hi_len >= 64 ? ~0ull : ((1ull << hi_len) - 1);
      return hi == ones ? low : 0ull;
    }
    case JA_SUF_SHIFT: return 1ull << len;   // ShiftSuffixPolynomial (poly/identity_poly.rs:160-166)
    default: return bits;   // JA_SUF_IDENTITY
  }
}

struct PsPhaseArgs {
  const unsigned long long* idx; const Fr* u; Fr* u_out; const Fr* v_prev;
  unsigned long long T;
  uint32_t prev_shift;             // k_bound of the previous phase = (k >> prev_shift) & m_mask
  uint32_t suffix_len, m_mask, bound, n_suf;
  uint32_t kinds[kPsMaxSuffixes];
  unsigned long long* gcols;       // [n_suf][m][kPsCols] global column table, zero on entry, zeroed again by the folding block
  unsigned int* counter;           // zero on entry, reset on exit
  Fr* host_out;                    // n_suf x m results, host-mapped, tagged 48-byte elements (store_tagged)
  unsigned int seq_value;
  unsigned int* sig_max;           // optional: max over the entries of the bits their SIGNED value needs (v in [-2^b, 2^b)); published as element n_suf * m
};
JA_DEV Fr fr_ld_cg(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 lo = __ldcg(q), hi = __ldcg(q + 1);
  Fr r;
  r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w; r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
  return r;
}
constexpr int kPsLimbs = 12;         // 32-bit limbs of an accumulator
constexpr int kPsCols = 2 * kPsLimbs;  // ... held as 16-bit columns in 32-bit counters (native shared-memory atomics)

// Scatter-add by an 8-bit key with INTEGER atomics on shared memory.  A (bin, suffix) accumulator is 24 columns: column c counts
// the 16-bit digits of weight 2^(16 c) of the plain integer products u * t (u a Montgomery residue < p, t < 2^64 in two 32-bit
// halves), so no carry is ever resolved atomically and one carry propagation + one Montgomery fold per accumulator finishes the
// block.
// A block accumulates kPsSufPerBlock suffixes (blockIdx.y selects which): 48 KB of static shared memory, no opt-in carve-out
// (a 147 KB dynamic allocation for all six suffixes made every launch reconfigure the SM's L1 / shared split).
constexpr int kPsSufPerBlock = 2;
__global__ void __launch_bounds__(256) k_ps_phase(const PsPhaseArgs a) {
  constexpr int NSUF = kPsSufPerBlock;
  __shared__ unsigned int s_acc[256 * kPsSufPerBlock * kPsCols];     // [m][NSUF][kPsCols]
  const uint32_t m = a.m_mask + 1;
  const uint32_t n_acc = m * NSUF;
  const uint32_t s0 = blockIdx.y * NSUF;                  // first suffix of this block
  for (uint32_t i = threadIdx.x; i < n_acc * kPsCols; i += blockDim.x) s_acc[i] = 0u;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  unsigned sig_local = 0;
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < a.T; base += (size_t)gridDim.x * blockDim.x) {
    const size_t j = base + threadIdx.x;
    Fr u = fp_zero<FrParams>();
    unsigned long long sb = 0ull;
    uint32_t key = 0xffffffffu;                                         // no entry
    if (j < a.T) {
      const unsigned long long k = a.idx[j];
      u = fp_load(a.u + j);
      if (a.v_prev) {                                                   // init_phase: u_evals[j] *= v[phase - 1][k_bound]
        u = fp_mul<FrParams>(u, fp_load(a.v_prev + ((k >> a.prev_shift) & a.m_mask)));
        if (blockIdx.y == gridDim.y - 1) fp_store(a.u_out + j, u);       // every suffix group forms the same product; one of them keeps it
      }
      sb = a.suffix_len >= 64 ? k : (k & ((1ull << a.suffix_len) - 1));
      key = (uint32_t)((a.suffix_len >= 64 ? 0ull : (k >> a.suffix_len)) & a.m_mask);
      if (a.sig_max && blockIdx.y == 0) {                               // bits of the signed value: 64 - clz(v >= 0 ? v : ~v)
        const unsigned long long mag = (k >> 63) ? ~k : k;
        sig_local = max(sig_local, 64u - (unsigned)__clzll((long long)mag));
      }
    }
    // Shared-memory atomics cost 2 cycles per LANE (64 per warp instruction, 18 of them per entry, suffix and half), and clamp
    // lookups send whole warps to the bins 0x00 / 0xff in most phases.  So the warp first settles its (up to) two most common bins
    // with FULL-mask warp reductions - lanes outside the bin contribute zeros, one lane issues the 18 atomics - and only the
    // lanes left over use their own atomics.  hot[0] = lane 0's bin, hot[1] = the first bin that differs from it.
    uint32_t hot[2];
    unsigned hot_mask[2];
    hot[0] = __shfl_sync(0xffffffffu, key, 0);
    hot_mask[0] = __ballot_sync(0xffffffffu, key == hot[0]);
    {
      const unsigned rest = ~hot_mask[0];
      const int src = rest ? __ffs(rest) - 1 : 0;
      hot[1] = __shfl_sync(0xffffffffu, key, src);
      hot_mask[1] = rest ? __ballot_sync(0xffffffffu, key == hot[1]) : (__ballot_sync(0xffffffffu, false));
    }
    int mine = -1;                                                      // which hot pass covers this lane (-1: own atomics)
#pragma unroll
    for (int g = 0; g < 2; g++)
      if (hot[g] != 0xffffffffu && __popc(hot_mask[g]) >= 4 && key == hot[g]) mine = g;
#pragma unroll 1
    for (int s = 0; s < NSUF; s++) {
      if (s0 + s >= a.n_suf) break;
      const uint32_t kind = a.kinds[s0 + s];
      const int halves = ((kind == JA_SUF_IDENTITY && a.suffix_len > 32) || (kind == JA_SUF_SHIFT && a.suffix_len >= 32)) ? 2 : 1;   // only the identity / shift suffixes exceed 32 bits
      const unsigned long long t = key != 0xffffffffu ? suffix_mle(kind, sb, a.suffix_len, a.bound) : 0ull;
#pragma unroll 1
      for (int h = 0; h < halves; h++) {
        const uint32_t th = h == 0 ? (uint32_t)t : (uint32_t)(t >> 32);
        uint32_t prod[9];                                               // u * th, 288 bits
        {
          unsigned long long c = 0;
#pragma unroll
          for (int i = 0; i < 8; i++) { c += (unsigned long long)u.l[i] * th; prod[i] = (uint32_t)c; c >>= 32; }
          prod[8] = (uint32_t)c;
        }
#pragma unroll
        for (int g = 0; g < 2; g++) {
          if (hot[g] == 0xffffffffu || __popc(hot_mask[g]) < 4) continue;          // warp-uniform
          if (__ballot_sync(0xffffffffu, mine == g && th != 0u) == 0u) continue;   // nothing to add for this bin (warp-uniform)
          unsigned int* dst = s_acc + ((size_t)hot[g] * NSUF + s) * kPsCols + 2 * h;
          const bool in = mine == g;
#pragma unroll
          for (int i = 0; i < 9; i++) {
            const uint32_t v = in ? prod[i] : 0u;
            const unsigned d0 = __reduce_add_sync(0xffffffffu, v & 0xffffu);       // 32 digits of 16 bits: below 2^21
            const unsigned d1 = __reduce_add_sync(0xffffffffu, v >> 16);
            if (lane == 0) {
              if (d0) atomicAdd(dst + 2 * i, d0);
              if (d1) atomicAdd(dst + 2 * i + 1, d1);
            }
          }
        }
        if (mine < 0 && th != 0u && key != 0xffffffffu) {
          unsigned int* dst = s_acc + ((size_t)key * NSUF + s) * kPsCols + 2 * h;
#pragma unroll
          for (int i = 0; i < 9; i++) {
            const unsigned int lo = prod[i] & 0xffffu, hi = prod[i] >> 16;
            if (lo) atomicAdd(dst + 2 * i, lo);
            if (hi) atomicAdd(dst + 2 * i + 1, hi);
          }
        }
      }
    }
  }
  if (a.sig_max && blockIdx.y == 0) {
    const unsigned mx = __reduce_max_sync(0xffffffffu, sig_local);
    if (lane == 0 && mx) atomicMax(a.sig_max, mx);
  }
  __syncthreads();
  // block -> grid: the non-zero 16-bit columns go to the phase's global column table with 64-bit reductions at L2 (fire and forget;
  // a column is a plain integer sum, so no carry is resolved here either).  No per-tile fold, no partial tables: the fold to a field
  // element happens ONCE per accumulator, in the block that finishes last.
  for (uint32_t i = threadIdx.x; i < n_acc * kPsCols; i += blockDim.x) {
    const unsigned int v = s_acc[i];
    if (v == 0u) continue;
    const uint32_t q = i / kPsCols, col = i % kPsCols;
    const uint32_t key = q / NSUF, sfx = s0 + q % NSUF;
    if (sfx < a.n_suf) atomicAdd(a.gcols + ((size_t)sfx * m + key) * kPsCols + col, (unsigned long long)v);
  }
  // the block that finishes last in its suffix group folds the group's accumulators and stores them to the host (mapped memory);
  // the group that finishes last raises the flag: one launch per phase, no D2H copy call
  __threadfence();
  __syncthreads();                                       // the shared accumulators are dead from here: word 0 carries the "last block" flag
  if (threadIdx.x == 0) s_acc[0] = atomicInc(a.counter + 1 + blockIdx.y, gridDim.x - 1) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_acc[0]) return;
  __threadfence();
  for (uint32_t q = threadIdx.x; q < n_acc; q += blockDim.x) {
    const uint32_t sfx = s0 + q / m, key = q % m;
    if (sfx >= a.n_suf) continue;
    unsigned long long* g = a.gcols + ((size_t)sfx * m + key) * kPsCols;
    unsigned long long cols[kPsCols];
#pragma unroll
    for (int i = 0; i < kPsCols; i += 2) {               // 24 independent L2 loads per accumulator, the table is zero again for the next phase
      const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(g + i));
      cols[i] = v.x; cols[i + 1] = v.y;
      *reinterpret_cast<ulonglong2*>(g + i) = make_ulonglong2(0ull, 0ull);
    }
    // columns -> integer X = lo + hi 2^256 (carry propagation), X mod p = mont(1_mont, lo) + mont(R^2, hi)
    uint32_t w[kPsLimbs + 1];
    unsigned long long c = 0;
#pragma unroll
    for (int i = 0; i < kPsLimbs; i++) {
      c += cols[2 * i];
      const uint32_t d0 = (uint32_t)(c & 0xffffu);
      c >>= 16;
      c += cols[2 * i + 1];
      const uint32_t d1 = (uint32_t)(c & 0xffffu);
      c >>= 16;
      w[i] = d0 | (d1 << 16);
    }
    w[kPsLimbs] = (uint32_t)c;
    Fr lo, hi = fp_zero<FrParams>();
#pragma unroll
    for (int i = 0; i < 8; i++) lo.l[i] = w[i];
#pragma unroll
    for (int i = 0; i < 5; i++) hi.l[i] = w[8 + i];
    const Fr val = fp_add<FrParams>(fp_mul<FrParams>(fp_one<FrParams>(), lo), fp_mul<FrParams>(fp_r2<FrParams>(), hi));
    store_tagged(a.host_out, (int)(sfx * m + key), val, a.seq_value);   // self-validating 48-byte element: no system fence, no flag
  }
  if (a.sig_max && blockIdx.y == 0 && threadIdx.x == 0) {              // every block of group 0 has added its maximum (fence + counter above)
    Fr v = fp_zero<FrParams>();
    v.l[0] = atomicExch(a.sig_max, 0u);
    store_tagged(a.host_out, (int)(a.n_suf * m), v, a.seq_value);
  }
}

// init_log_t_rounds (mod.rs:427-441): ra[j] = prod_phase v[phase][(k >> ((phases - 1 - phase) * log_m)) & m_mask]
__global__ void __launch_bounds__(kBlock) k_ps_ra(const unsigned long long* __restrict__ idx, size_t T, const Fr* __restrict__ v /* [phases][m] */,
                                                  uint32_t phases, uint32_t log_m, Fr* __restrict__ out) {
  const uint32_t m_mask = (1u << log_m) - 1;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < T; j += stride) {
    const unsigned long long k = idx[j];
    Fr acc = fp_load(v + ((k >> ((phases - 1) * log_m)) & m_mask));
    for (uint32_t ph = 1; ph < phases; ph++)
      acc = fp_mul<FrParams>(acc, fp_load(v + (size_t)ph * (m_mask + 1) + ((k >> ((phases - 1 - ph) * log_m)) & m_mask)));
    fp_store(out + j, acc);
  }
}

}  // namespace

struct ja_psshout {
  unsigned long long* d_idx = nullptr;
  Fr* d_u = nullptr;                 // u_evals: eq(r_cycle, j), then times the expanding tables of the finished phases
  Fr* d_u2 = nullptr;                // ping-pong partner (a phase reads the old values in every suffix group)
  unsigned long long* d_cols = nullptr;   // global column table of the phase pass (zero between passes)
  size_t T = 0;
  uint32_t log_k = 0, phases = 0, log_m = 0;
  uint32_t next_phase = 0;
  std::vector<ja::host::FrH> h_v;    // expanding tables of ja_psshout_prove_address (phases x m), for ja_psshout_materialize_ra
};

extern "C" {

static int32_t psshout_new_impl(ja_ctx* c, const uint64_t* lookup_indices, bool on_device, size_t T, const uint64_t* r_cycle, size_t log_t, uint32_t log_k,
                                uint32_t phases, ja_psshout** out);
int32_t ja_psshout_new(ja_ctx* c, const uint64_t* lookup_indices, size_t T, const uint64_t* r_cycle, size_t log_t, uint32_t log_k,
                       uint32_t phases, ja_psshout** out) {
  return psshout_new_impl(c, lookup_indices, false, T, r_cycle, log_t, log_k, phases, out);
}
// the same over lookup indices that already live in HBM (witness.cu): device-to-device copy
int32_t ja_psshout_new_dev(ja_ctx* c, const unsigned long long* d_indices, size_t T, const uint64_t* r_cycle, size_t log_t, uint32_t log_k,
                           uint32_t phases, ja_psshout** out) {
  return psshout_new_impl(c, reinterpret_cast<const uint64_t*>(d_indices), true, T, r_cycle, log_t, log_k, phases, out);
}
static int32_t psshout_new_impl(ja_ctx* c, const uint64_t* lookup_indices, bool on_device, size_t T, const uint64_t* r_cycle, size_t log_t, uint32_t log_k,
                                uint32_t phases, ja_psshout** out) {
  JA_REQUIRE(c && lookup_indices && out && (r_cycle || log_t == 0), "ja_psshout_new: null argument");
  JA_REQUIRE(T == (size_t(1) << log_t) && phases >= 1 && log_k >= phases && log_k <= 64 && log_k % phases == 0 && log_k / phases <= 8,
             "ja_psshout_new: T = 2^log_t, LOG_K a multiple of the number of phases, at most 8 address bits per phase");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  ja_psshout* p = new ja_psshout();
  p->T = T; p->log_k = log_k; p->phases = phases; p->log_m = log_k / phases;
  int32_t st = dev_alloc(c, T * 8, (void**)&p->d_idx);
  if (st) { delete p; return st; }
  JA_CUDA(cudaMemcpyAsync(p->d_idx, lookup_indices, T * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
  // u_evals = EqPolynomial::evals(r_node_output) (mod.rs:236)
  ja_poly* eq = nullptr;
  if ((st = ja_eq_evals(c, r_cycle, log_t, nullptr, &eq))) { dev_free(c, p->d_idx); delete p; return st; }
  p->d_u = eq->buf[eq->cur];
  eq->buf[eq->cur] = nullptr;
  ja_poly_free(c, eq);
  if (!on_device) JA_CUDA(cudaStreamSynchronize(c->stream));            // the host index array is borrowed for the duration of the call
  *out = p;
  return JA_OK;
}

static int32_t ps_init_phase(ja_ctx* c, ja_psshout* p, uint32_t phase, const uint64_t* v_prev, const uint32_t* suffix_kinds, size_t n_suffixes,
                             uint32_t bound, uint64_t* out_Q, uint32_t* out_sigbits);
int32_t ja_psshout_init_phase(ja_ctx* c, ja_psshout* p, uint32_t phase, const uint64_t* v_prev, const uint32_t* suffix_kinds, size_t n_suffixes,
                              uint32_t bound, uint64_t* out_Q) {
  JA_REQUIRE(c && p, "ja_psshout_init_phase: bad argument");
  JA_REQUIRE(phase == p->next_phase && phase < p->phases, "ja_psshout_init_phase: phases run in order 0 .. NUM_PHASES - 1");
  return ps_init_phase(c, p, phase, v_prev, suffix_kinds, n_suffixes, bound, out_Q, nullptr);
}
// out_sigbits (optional): the largest number of bits the SIGNED value of a lookup index needs (scan fused into the pass).
// `phase` may skip ahead of p->next_phase (ja_psshout_prove_address: leading phases done on the host): v_prev then stands for the
// PRODUCT of the skipped phases' expanding-table entries, indexed by the chunk of phase - 1.
static int32_t ps_init_phase(ja_ctx* c, ja_psshout* p, uint32_t phase, const uint64_t* v_prev, const uint32_t* suffix_kinds, size_t n_suffixes,
                             uint32_t bound, uint64_t* out_Q, uint32_t* out_sigbits) {
  JA_REQUIRE(c && p && suffix_kinds && out_Q && n_suffixes >= 1 && n_suffixes <= (size_t)kPsMaxSuffixes, "ja_psshout_init_phase: bad argument");
  JA_REQUIRE(phase >= p->next_phase && phase < p->phases, "ja_psshout_init_phase: phases run in order 0 .. NUM_PHASES - 1");
  JA_REQUIRE((phase == 0) == (v_prev == nullptr), "ja_psshout_init_phase: the expanding table of the previous phase is required from phase 1 on");
  for (size_t s = 0; s < n_suffixes; s++) JA_REQUIRE(suffix_kinds[s] <= JA_SUF_SHIFT, "ja_psshout_init_phase: unknown suffix kind");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  static const bool ps_trace = getenv("JA_PS_TRACE") != nullptr;
  const auto tr0 = std::chrono::steady_clock::now();
  JA_CUDA(cudaSetDevice(c->device));
  const uint32_t m = 1u << p->log_m;
  static const int tile_log = getenv("JA_PS_TILE_LOG") ? atoi(getenv("JA_PS_TILE_LOG")) : 8;
  const unsigned groups = (unsigned)((n_suffixes + kPsSufPerBlock - 1) / kPsSufPerBlock);
  // 2^tile_log entries per block (one or two per thread): many small blocks, several per SM - a block's cost is its per-entry
  // instruction chain (~1000 instructions per entry and suffix pair), so the pass wants all the warps the SMs can hold
  uint32_t tiles = (uint32_t)((p->T + (size_t(1) << tile_log) - 1) >> tile_log);
  const uint32_t max_tiles = std::max<uint32_t>(1u, 4u * (uint32_t)kSMs / groups);
  if (tiles > max_tiles) tiles = std::max<uint32_t>(max_tiles, (uint32_t)((p->T + (size_t(1) << 15) - 1) >> 15));   // a 32-bit shared column takes 2^16 digits of 16 bits
  JA_REQUIRE((p->T + tiles - 1) / tiles <= (size_t(1) << 15), "ja_psshout_init_phase: T too large for the column counters");
  const size_t n_out = n_suffixes * m;
  JA_REQUIRE((n_out + 1) * 48 <= kRowSeqOffset, "ja_psshout_init_phase: result larger than the mapped value buffer");
  Fr* d_v = nullptr;
  int32_t st;
  if (!p->d_cols) {
    const size_t bytes = (size_t)kPsMaxSuffixes * 256 * kPsCols * sizeof(unsigned long long);
    if ((st = dev_alloc(c, bytes, (void**)&p->d_cols))) return st;
    JA_CUDA(cudaMemsetAsync(p->d_cols, 0, bytes, c->stream));
  }
  if (v_prev) {
    if ((st = dev_alloc(c, m * sizeof(Fr), (void**)&d_v))) return st;
    if ((st = stage_h2d(c, d_v, v_prev, m * sizeof(Fr)))) { dev_free(c, d_v); return st; }
  }
  PsPhaseArgs a;
  memset(&a, 0, sizeof(a));
  a.idx = p->d_idx; a.u = p->d_u; a.v_prev = d_v; a.T = p->T;
  a.prev_shift = (p->phases - phase) * p->log_m;                       // mod.rs:279: k.split((phases - phase) * log_m)
  a.suffix_len = (p->phases - 1 - phase) * p->log_m;                   // mod.rs:309
  a.m_mask = m - 1; a.bound = bound; a.n_suf = (uint32_t)n_suffixes;
  for (size_t s = 0; s < n_suffixes; s++) a.kinds[s] = suffix_kinds[s];
  a.gcols = p->d_cols;
  a.sig_max = out_sigbits ? c->d_counter + 8 + 1 + kPsMaxSuffixes : nullptr;     // behind the group counters, zero between passes
  a.counter = c->d_counter + 8;                                        // [0]: groups done, [1 + y]: blocks done of group y
  a.host_out = reinterpret_cast<Fr*>(c->d_rowvals);
  a.seq_value = next_tag(c);
  // u_evals ping-pong: the suffix groups of a phase all read the OLD u_evals while one of them writes the new ones
  if (v_prev) {
    if (!p->d_u2 && (st = dev_alloc(c, p->T * sizeof(Fr), (void**)&p->d_u2))) { dev_free(c, d_v); return st; }
    a.u_out = p->d_u2;
  }
  JA_LAUNCH(c, KC_SCATTER, k_ps_phase<<<dim3(tiles, groups), 256, 0, c->stream>>>(a));
  if (v_prev) std::swap(p->d_u, p->d_u2);
  cudaError_t e = cudaGetLastError();
  dev_free(c, d_v);                                                    // stream-ordered reuse
  if (e != cudaSuccess) return fail(JA_ERR_CUDA, std::string("ja_psshout_init_phase: ") + cudaGetErrorString(e));
  const auto tr1 = std::chrono::steady_clock::now();
  {
    // the results land in mapped host memory as self-validating tagged elements (bounded wait: common.hpp wait_tagged)
    if ((st = wait_tagged(c, c->h_rowvals, a.seq_value, n_out, out_Q, "ja_psshout_init_phase"))) return st;
    if (out_sigbits) {
      uint64_t sv[4];
      if ((st = wait_tagged(c, reinterpret_cast<const char*>(c->h_rowvals) + n_out * 48, a.seq_value, 1, sv, "ja_psshout_init_phase"))) return st;
      *out_sigbits = (uint32_t)sv[0];
    }
    if (ps_trace) {
      const auto tr2 = std::chrono::steady_clock::now();
      auto us = [](std::chrono::steady_clock::time_point x, std::chrono::steady_clock::time_point y) { return std::chrono::duration<double, std::micro>(y - x).count(); };
      fprintf(stderr, "[ps_trace] phase %u T %zu tiles %u: submit %.1f us, wait %.1f us\n", phase, p->T, tiles, us(tr0, tr1), us(tr1, tr2));
    }
  }
  p->next_phase = phase + 1;
  return JA_OK;
}

int32_t ja_psshout_materialize_ra(ja_ctx* c, ja_psshout* p, const uint64_t* v, const uint64_t* scale, ja_poly** out_ra) {
  JA_REQUIRE(c && p && out_ra, "ja_psshout_materialize_ra: null argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  const uint32_t m = 1u << p->log_m;
  const size_t n_tab = (size_t)p->phases * m;
  const size_t tab_bytes = n_tab * sizeof(Fr);
  JA_REQUIRE(v || p->h_v.size() == n_tab, "ja_psshout_materialize_ra: no expanding tables (pass v, or run ja_psshout_prove_address first)");
  std::vector<ja::host::FrH> tab(n_tab);
  memcpy((void*)tab.data(), v ? (const void*)v : (const void*)p->h_v.data(), tab_bytes);
  if (scale) {
    // ra * (val + raf_val) (the constant factor of the cycle rounds, mod.rs:484-487), folded into the LAST expanding table: the
    // round polynomials of the cycle rounds over the scaled ra are the reference's gruen_poly_deg_2(eval_at_0 * (val + raf_val), claim)
    const ja::host::FrH sc = ja::host::from_limbs(scale);
    for (size_t i = n_tab - m; i < n_tab; i++) tab[i] = ja::host::mul(tab[i], sc);
  }
  Fr* d_v = nullptr;
  int32_t st = dev_alloc(c, tab_bytes, (void**)&d_v);
  if (st) return st;
  if ((st = stage_h2d(c, d_v, tab.data(), tab_bytes))) { dev_free(c, d_v); return st; }
  if ((st = ja_poly_alloc(c, p->T, out_ra))) { dev_free(c, d_v); return st; }
  JA_LAUNCH(c, KC_CONVERT, k_ps_ra<<<grid_for(p->T), kBlock, 0, c->stream>>>(p->d_idx, p->T, d_v, p->phases, p->log_m, (*out_ra)->buf[0]));
  JA_CUDA(cudaGetLastError());
  dev_free(c, d_v);
  return JA_OK;
}

// ---- the LOG_K address rounds (ps_shout/mod.rs:337-418, :491-560 under Sumcheck::prove, subprotocols/sumcheck.rs:565-599) ----------
// The tables of these rounds have m = 2^(LOG_K / phases) <= 256 entries: they stay on the HOST, next to the transcript - a
// device round would pay a host<->device round trip (3.8 us measured, profiles/r2_persist_probe.txt) for ~1 us of arithmetic.
// The T-sized work of a phase (init_phase: u_evals update + six suffix-polynomial scatters) is the kernel above; its Q rows
// arrive in mapped host memory.  The reference evaluates four prefix MLEs per (b, c) and ClampSpec::combine per b
// (lookup_tables/prefixes/{higher_all_zero,higher_all_one,lower_word,msb}.rs, clamp.rs:94-118).  Those prefixes depend on b only
// through (1) the indicator that b's "higher" bits are all zero / all one and (2) the integer value of b's lower-word bits, so
// the sum over b of combine(...) regroups EXACTLY (field arithmetic) into a handful of plain sums of the suffix polynomials over
// contiguous b ranges and index-weighted sums  sum_b b * Q[b]  (two additions per entry), times per-round scalars: a round costs
// O(m) additions and ~40 products instead of ~60 m products.  Same for the raf part: the signed-identity prefix polynomial
// (poly/signed_identity_poly.rs:183-217) is affine in b, so its H2L-bound values are  B + c * kappa + 2^s * b.
namespace {
using ja::host::FrH;
using ja::host::add; using ja::host::sub; using ja::host::mul; using ja::host::dbl; using ja::host::neg;

struct PsCp { bool has = false; FrH v; };
enum { P_HAZ = 0, P_HAO = 1, P_LW = 2, P_MSB = 3 };

// the b-independent factor / offset of SparseDensePrefix::prefix_mle at round j (r_x = the previous challenge when j is odd)
static FrH ps_prefix_scalar(int kind, const PsCp cp[4], const FrH* r_x, uint32_t c, unsigned j, unsigned bound_index, const FrH* pow2, unsigned xlen) {
  // c is 0, 1 or 2: products by c / (1 - c) are selections, additions and negations; r_x is a 125-bit challenge (mul_chal)
  auto times_c = [&](const FrH& x) { return c == 0 ? ja::host::FR_ZERO : (c == 1 ? x : dbl(x)); };
  auto times_1mc = [&](const FrH& x) { return c == 0 ? x : (c == 1 ? ja::host::FR_ZERO : neg(x)); };
  if (kind == P_MSB) return j == 0 ? times_c(ja::host::FR_ONE) : (j == 1 ? *r_x : cp[P_MSB].v);
  if (kind == P_HAZ || kind == P_HAO) {
    const bool zero = kind == P_HAZ;
    FrH r = cp[kind].has ? cp[kind].v : ja::host::FR_ONE;
    if (r_x && j > 0 && j - 1 <= bound_index) r = zero ? sub(r, ja::host::mul_chal(r, *r_x)) : ja::host::mul_chal(r, *r_x);   // r (1 - r_x) = r - r r_x
    if (j <= bound_index) r = zero ? times_1mc(r) : times_c(r);
    return r;
  }
  FrH r = cp[P_LW].has ? cp[P_LW].v : ja::host::FR_ZERO;
  if (r_x && j > 0 && j - 1 > bound_index) r = add(r, ja::host::mul_chal(pow2[xlen - (j - 1) - 1], *r_x));
  if (j > bound_index) r = add(r, times_c(pow2[xlen - j - 1]));
  return r;
}
static PsCp ps_update_checkpoint(int kind, const PsCp cp[4], const FrH& r_x, const FrH& r_y, unsigned j, unsigned bound_index, const FrH* pow2, unsigned xlen) {
  PsCp o;
  if (kind == P_MSB) {
    if (j == 0) return o;
    if (j == 1) { o.has = true; o.v = r_x; return o; }
    return cp[P_MSB];
  }
  o.has = true;
  if (kind == P_HAZ || kind == P_HAO) {
    const bool zero = kind == P_HAZ;
    FrH r = cp[kind].has ? cp[kind].v : ja::host::FR_ONE;
    if (j > 0 && j - 1 <= bound_index) r = mul(r, zero ? sub(ja::host::FR_ONE, r_x) : r_x);
    if (j <= bound_index) r = mul(r, zero ? sub(ja::host::FR_ONE, r_y) : r_y);
    o.v = r;
    return o;
  }
  FrH r = cp[P_LW].has ? cp[P_LW].v : ja::host::FR_ZERO;
  if (j > 0 && j - 1 > bound_index) r = add(r, mul(pow2[xlen - (j - 1) - 1], r_x));
  if (j > bound_index) r = add(r, mul(pow2[xlen - j - 1], r_y));
  o.v = r;
  return o;
}
// Plain sums of up to 256 field elements run as 320-bit INTEGER accumulations (4 add-with-carry per element instead of a modular
// addition with its compare-and-subtract: 14 -> 2 ns per element measured) and fold back into the field once: X mod p =
// mont(lo, 1_mont) + mont(hi, R^2) (sumcheck_host.hpp fold320).
struct PsAcc { uint64_t l[5] = {0, 0, 0, 0, 0}; };
static inline void ps_acc_add(PsAcc& a, const uint64_t* x, int limbs) {
  unsigned __int128 c = 0;
  for (int i = 0; i < 5; i++) { c += (unsigned __int128)a.l[i] + (i < limbs ? x[i] : 0ull); a.l[i] = (uint64_t)c; c >>= 64; }
}
static inline FrH ps_acc_fold(const PsAcc& a) { return ja::host::fold320(a.l, ja::host::FR_ONE, ja::host::FR_R2); }
static inline FrH ps_sum(const FrH* x, size_t n) {
  if (n <= 4) { FrH a = ja::host::FR_ZERO; for (size_t i = 0; i < n; i++) a = add(a, x[i]); return a; }
  PsAcc a;
  for (size_t i = 0; i < n; i++) ps_acc_add(a, x[i].l, 4);
  return ps_acc_fold(a);
}
// sum = sum_b x[b] and wsum = sum_b b * x[b], b < n <= 256: the running suffix sum added once per step (acc < 2^262, tot < 2^270)
static inline void ps_sum_wsum(const FrH* x, size_t n, FrH* sum, FrH* wsum) {
  if (n <= 2) {
    *wsum = n == 2 ? x[1] : ja::host::FR_ZERO;
    *sum = n == 2 ? add(x[0], x[1]) : (n == 1 ? x[0] : ja::host::FR_ZERO);
    return;
  }
  PsAcc acc, tot;
  for (size_t b = n; b-- > 1;) { ps_acc_add(acc, x[b].l, 4); ps_acc_add(tot, acc.l, 5); }
  ps_acc_add(acc, x[0].l, 4);
  *sum = ps_acc_fold(acc); *wsum = ps_acc_fold(tot);
}
static inline FrH ps_wsum(const FrH* x, size_t n) { FrH s_, w_; ps_sum_wsum(x, n, &s_, &w_); return w_; }
struct PsSide { FrH s1, r1, w_all, a_haz, a_hz, a_one, a_w, b_ho, b_one, b_w; };
}  // namespace

int32_t ja_psshout_prove_address(ja_ctx* c, ja_psshout* p, uint32_t bound, const uint64_t* gamma_in, const uint64_t* claim_in, uint8_t state[32],
                                 uint32_t* n_rounds, uint64_t* out_coeffs, uint32_t* out_ncoeffs, uint64_t* out_challenges, uint64_t* out_input_claim,
                                 uint64_t* out_val, uint64_t* out_raf_val, uint64_t* out_claim) {
  JA_REQUIRE(c && p && gamma_in && state && n_rounds && out_coeffs && out_ncoeffs && out_challenges, "ja_psshout_prove_address: null argument");
  JA_REQUIRE(p->next_phase == 0, "ja_psshout_prove_address: the address rounds start from a fresh state (phase 0)");
  const unsigned xlen = p->log_k, log_m = p->log_m, phases = p->phases;
  JA_REQUIRE(bound < xlen && xlen <= 64 && log_m >= 2, "ja_psshout_prove_address: clamp bound must be below the index width");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  const size_t m = size_t(1) << log_m;
  const unsigned bound_index = xlen - bound - 1;
  FrH pow2[66];
  pow2[0] = ja::host::FR_ONE;
  for (int i = 1; i < 66; i++) pow2[i] = dbl(pow2[i - 1]);
  const FrH gamma = ja::host::from_limbs(gamma_in);
  const FrH const_upper = sub(pow2[bound], ja::host::FR_ONE);
  const FrH lower_coeff = add(dbl(const_upper), ja::host::FR_ONE);
  const uint32_t kinds[6] = {JA_SUF_HIGHER_ALL_ZERO, JA_SUF_HZERO_MUL_LWORD, JA_SUF_HONE_MUL_LWORD, JA_SUF_ONE, JA_SUF_ONE, JA_SUF_IDENTITY};
  ja::host::Blake2bTranscript t(state, *n_rounds);
  PsCp cp[4];
  FrH cp_id = ja::host::FR_ZERO;                         // PrefixRegistry checkpoint of Prefix::SignedIdentity (None = 0)
  FrH claim = claim_in ? ja::host::from_limbs(claim_in) : ja::host::FR_ZERO;
  std::vector<FrH> Qbuf(6 * m), r_all;
  std::vector<FrH> v_cur, v_next;
  p->h_v.assign((size_t)phases * m, ja::host::FR_ZERO);
  FrH r_prev = ja::host::FR_ZERO;
  // Sign-extension phases.  Clamp lookups index the table with small SIGNED values (a rescaled accumulation): when every index needs at
  // most `sig` bits (v in [-2^sig, 2^sig), sig <= BOUND; scanned inside the phase-0 pass), the chunk of every phase p whose suffix is
  // at least sig bits long is 0x00 (v >= 0) or 0xff (v < 0) for EVERY entry.  The suffix polynomials of such a phase have two non-zero
  // entries, u_evals = eq * (product of the finished phases' v[0x00] resp. v[0xff]), and every suffix MLE is affine in v on a sign
  // class (t = alpha + beta v): Q_s[0x00] = cP (alpha S_P + beta V_P), Q_s[0xff] = cN (alpha S_N + beta V_N) with the four sums
  // S = sum eq, V = sum eq v per class - all four are entries of the phase-0 tables.  So those phases need NO pass over the T entries:
  // the host builds their tables, and the first real pass afterwards takes the product table [cP .. cN] as its "previous" one.
  // (alpha, beta) come from the suffix MLE itself at two representatives and are checked at a third: a suffix that is not affine
  // on the class switches the fast path off.  Same tables, same transcript as the reference's eight passes (tests/test_gpu_psshout.py).
  static const bool no_skip = getenv("JA_PS_NO_SKIP") != nullptr;
  unsigned n_virtual_end = 0;                              // phases 1 .. n_virtual_end - 1 are built on the host
  FrH S_cls[2], V_cls[2], c_cls[2] = {ja::host::FR_ONE, ja::host::FR_ONE};
  uint32_t sig = 64;
  auto suffix_len_of = [&](unsigned ph) { return xlen - (ph + 1) * log_m; };
  auto affine = [&](uint32_t kind, unsigned len, int cls, FrH* alpha, int* beta) -> bool {
    const unsigned long long mask = len >= 64 ? ~0ull : ((1ull << len) - 1);
    const long long v0 = cls ? -1ll : 0ll, v1 = cls ? -2ll : 1ll, v2 = cls ? -(1ll << sig) : ((1ll << sig) - 1);
    const unsigned long long t0 = suffix_mle(kind, (unsigned long long)v0 & mask, len, bound), t1 = suffix_mle(kind, (unsigned long long)v1 & mask, len, bound),
                             t2 = suffix_mle(kind, (unsigned long long)v2 & mask, len, bound);
    const __int128 b = sig == 0 ? 0 : ((__int128)t1 - (__int128)t0) / (v1 - v0);
    if (b != 0 && b != 1) return false;
    const __int128 a0 = (__int128)t0 - b * v0;
    if (a0 < 0 || a0 > (__int128)~0ull) return false;
    if (sig != 0 && (__int128)t1 != a0 + b * v1) return false;
    if ((__int128)t2 != a0 + b * v2) return false;
    *alpha = ja::host::from_u64((uint64_t)a0); *beta = (int)b;
    return true;
  };
  for (unsigned phase = 0; phase < phases; phase++) {
    int32_t st = JA_OK;
    if (phase >= 1 && phase < n_virtual_end) {
      bool ok = true;
      std::fill(Qbuf.begin(), Qbuf.end(), ja::host::FR_ZERO);
      for (int r = 0; r < 6 && ok; r++)
        for (int cls = 0; cls < 2 && ok; cls++) {
          FrH alpha; int beta = 0;
          ok = affine(kinds[r], suffix_len_of(phase), cls, &alpha, &beta);
          if (!ok) break;
          FrH a = mul(alpha, S_cls[cls]);
          if (beta) a = add(a, V_cls[cls]);
          Qbuf[(size_t)r * m + (cls ? m - 1 : 0)] = mul(c_cls[cls], a);
        }
      JA_REQUIRE(ok, "ja_psshout_prove_address: internal error (sign-extension phase lost its affine form)");   // checked for every phase up front
    } else {
      std::vector<FrH> tab;
      const uint64_t* vprev = nullptr;
      if (phase) {
        if (phase == n_virtual_end && n_virtual_end > 1) {         // first real pass after host-built phases: u_evals = eq * c[class]
          tab.assign(m, ja::host::FR_ZERO);
          tab[0] = c_cls[0]; tab[m - 1] = c_cls[1];
          vprev = reinterpret_cast<const uint64_t*>(tab.data());
        } else {
          vprev = reinterpret_cast<const uint64_t*>(p->h_v.data() + (size_t)(phase - 1) * m);
        }
      }
      st = ps_init_phase(c, p, phase, vprev, kinds, 6, bound, reinterpret_cast<uint64_t*>(Qbuf.data()), phase == 0 && !no_skip ? &sig : nullptr);
      if (st) return st;
      if (phase == 0 && !no_skip && sig <= bound && sig < 63 && m >= 2) {
        unsigned h = 0;
        while (h < phases && suffix_len_of(h) >= sig) h++;
        bool ok = h >= 2;
        for (unsigned ph = 1; ph < h && ok; ph++)
          for (int r = 0; r < 6 && ok; r++)
            for (int cls = 0; cls < 2 && ok; cls++) { FrH al; int be; ok = affine(kinds[r], suffix_len_of(ph), cls, &al, &be); }
        if (ok) {
          // the four class sums from the phase-0 tables: One suffix -> S, Identity suffix -> V (for v < 0 the suffix value is 2^len + v)
          S_cls[0] = Qbuf[3 * m]; S_cls[1] = Qbuf[3 * m + m - 1];
          V_cls[0] = Qbuf[5 * m]; V_cls[1] = sub(Qbuf[5 * m + m - 1], mul(pow2[suffix_len_of(0)], S_cls[1]));
          n_virtual_end = h;
        }
      }
    }
    // rows: 0 higher-all-zero, 1 hzero*lword, 2 hone*lword, 3 one (== row 4, the raf decomposition's One suffix), 5 identity
    FrH* Q[5] = {Qbuf.data(), Qbuf.data() + m, Qbuf.data() + 2 * m, Qbuf.data() + 3 * m, Qbuf.data() + 5 * m};
    const unsigned s_len = xlen - (phase + 1) * log_m;                 // suffix_len of the phase
    FrH bid = cp_id;
    v_cur.assign(1, ja::host::FR_ONE);
    const auto t_rounds = std::chrono::steady_clock::now();
    double tr_acc[4] = {0, 0, 0, 0};
    static const bool tr_on = getenv("JA_PS_TRACE") != nullptr;
    auto tr_now = [] { return std::chrono::steady_clock::now(); };
    auto tr_us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    for (unsigned tt = 0; tt < log_m; tt++) {
      const auto tq0 = tr_now();
      const unsigned j = phase * log_m + tt, b_len = log_m - 1 - tt;
      const size_t half = size_t(1) << b_len;
      const unsigned n_hi = bound_index >= j ? std::min<unsigned>(bound_index - j, b_len) : 0;   // b's top n_hi bits are "higher" bits
      const unsigned n_lo = b_len - n_hi;
      const size_t n_z = size_t(1) << n_lo;                            // Z = [0, n_z): higher bits of b all zero; O = [half - n_z, half): all one
      const size_t o_off = half - n_z;
      PsSide side[2];
      for (int sd = 0; sd < 2; sd++) {
        const size_t off = sd ? half : 0;
        PsSide& S = side[sd];
        ps_sum_wsum(Q[3] + off, half, &S.s1, &S.w_all);
        S.r1 = ps_sum(Q[4] + off, half);
        S.a_haz = ps_sum(Q[0] + off, n_z);
        S.a_hz = ps_sum(Q[1] + off, n_z);
        if (n_hi == 0) { S.a_one = S.s1; S.a_w = S.w_all; S.b_one = S.s1; S.b_w = S.w_all; }
        else {
          ps_sum_wsum(Q[3] + off, n_z, &S.a_one, &S.a_w);
          ps_sum_wsum(Q[3] + off + o_off, n_z, &S.b_one, &S.b_w);                              // (b & low_mask) = b - o_off on O
        }
        S.b_ho = ps_sum(Q[2] + off + o_off, n_z);
      }
      const FrH* r_x = (j & 1) ? &r_prev : nullptr;
      const FrH two_s = pow2[s_len];
      const FrH kappa = j == 0 ? neg(pow2[xlen - 1]) : pow2[b_len + s_len];        // coefficient of the current variable in the identity prefix
      // E(c, side) = sum_b combine(P_c(b), Q_side(b)); raf(c, side) = sum_b P_c(b) Q_one(b) + Q_id(b)
      const auto tq1 = tr_now();
      // products that do not depend on c, once per side: 2^s A_w - CU A_haz, 2^s B_w, 2^s W_all
      FrH zc[2], oc[2], wc[2];
      for (int sd = 0; sd < 2; sd++) {
        zc[sd] = sub(mul(two_s, side[sd].a_w), mul(const_upper, side[sd].a_haz));
        oc[sd] = n_hi == 0 ? mul(two_s, side[sd].a_w) : mul(two_s, side[sd].b_w);
        wc[sd] = n_hi == 0 ? oc[sd] : mul(two_s, side[sd].w_all);
      }
      struct PsScal { FrH hz, ho, lw, msb_term; };
      auto scalars = [&](uint32_t cc) {
        PsScal q;
        q.hz = ps_prefix_scalar(P_HAZ, cp, r_x, cc, j, bound_index, pow2, xlen); q.ho = ps_prefix_scalar(P_HAO, cp, r_x, cc, j, bound_index, pow2, xlen);
        q.lw = ps_prefix_scalar(P_LW, cp, r_x, cc, j, bound_index, pow2, xlen);
        q.msb_term = sub(const_upper, mul(ps_prefix_scalar(P_MSB, cp, r_x, cc, j, bound_index, pow2, xlen), lower_coeff));
        return q;
      };
      auto table_part = [&](const PsScal& q, int sd) {
        const PsSide& S = side[sd];
        FrH e = mul(q.msb_term, S.s1);
        const FrH zt = add(add(S.a_hz, mul(q.lw, S.a_one)), zc[sd]);
        const FrH ot = add(add(S.b_ho, mul(q.lw, S.b_one)), oc[sd]);
        e = add(e, mul(q.hz, zt));
        return add(e, mul(q.ho, ot));
      };
      auto raf_part = [&](uint32_t cc, int sd) {
        const FrH base = cc == 0 ? bid : (cc == 1 ? add(bid, kappa) : add(bid, dbl(kappa)));
        return add(add(mul(base, side[sd].s1), wc[sd]), side[sd].r1);
      };
      const PsScal q0 = scalars(0), q2 = scalars(2);
      const FrH e0 = add(table_part(q0, 0), mul(gamma, raf_part(0, 0)));
      if (j == 0) {
        // the claimed sum as the prover sees it: s(0) + s(1) = rv(r_cycle) + gamma * operand(r_cycle)
        const PsScal q1 = scalars(1);
        const FrH derived = add(e0, add(table_part(q1, 1), mul(gamma, raf_part(1, 1))));
        if (out_input_claim) memcpy(out_input_claim, derived.l, 32);
        if (!claim_in) claim = derived;
      }
      const FrH t2l = table_part(q2, 0), t2h = table_part(q2, 1);
      const FrH a2l = raf_part(2, 0), a2r = raf_part(2, 1);
      const FrH e2 = add(sub(dbl(t2h), t2l), mul(gamma, sub(dbl(a2r), a2l)));
      const auto tq2 = tr_now();
      const ja::host::Coeffs uni = ja::host::from_evals_and_hint(claim, {e0, e2});
      const ja::host::Coeffs cpr = ja::host::compress(uni);
      JA_REQUIRE(cpr.size() <= 2, "ja_psshout_prove_address: round polynomial of degree > 2");
      t.append_message("UniPoly_begin");
      for (auto& x : cpr) t.append_scalar(x);
      t.append_message("UniPoly_end");
      uint64_t ch[4];
      t.challenge_scalar_optimized(ch);
      const FrH rj = ja::host::from_limbs(ch);
      claim = ja::host::evaluate(uni, rj);
      out_ncoeffs[j] = (uint32_t)cpr.size();
      for (size_t k = 0; k < 2; k++) memcpy(out_coeffs + 4 * (2 * j + k), k < cpr.size() ? cpr[k].l : ja::host::FR_ZERO.l, 32);
      memcpy(out_challenges + 4 * j, ch, 32);
      const auto tq3 = tr_now();
      // ingest_challenge (mod.rs:491-560)
      // H2L binds: clamp lookups leave most entries of most rows zero in the sign-extension phases - a zero difference costs no product
      for (int q = 0; q < 5; q++)
        for (size_t b = 0; b < half; b++) {
          const FrH &lo_ = Q[q][b], &hi_ = Q[q][b + half];
          if ((lo_.l[0] | lo_.l[1] | lo_.l[2] | lo_.l[3] | hi_.l[0] | hi_.l[1] | hi_.l[2] | hi_.l[3]) == 0) continue;   // both zero: nothing to bind
          const FrH d = sub(hi_, lo_);
          if (!d.is_zero()) Q[q][b] = add(lo_, ja::host::mul_chal(d, rj));
        }
      bid = add(bid, ja::host::mul_chal(kappa, rj));
      v_next.resize(v_cur.size() * 2);                                             // ExpandingTable::update, HighToLow (expanding_table.rs:76-86)
      for (size_t i = 0; i < v_cur.size(); i++) { const FrH e1 = ja::host::mul_chal(v_cur[i], rj); v_next[2 * i] = sub(v_cur[i], e1); v_next[2 * i + 1] = e1; }
      v_cur.swap(v_next);
      if (j & 1) {
        PsCp prev[4] = {cp[0], cp[1], cp[2], cp[3]};
        for (int k = 0; k < 4; k++) cp[k] = ps_update_checkpoint(k, prev, r_prev, rj, j, bound_index, pow2, xlen);
      }
      r_prev = rj;
      if (tr_on) { const auto tq4 = tr_now(); tr_acc[0] += tr_us(tq0, tq1); tr_acc[1] += tr_us(tq1, tq2); tr_acc[2] += tr_us(tq2, tq3); tr_acc[3] += tr_us(tq3, tq4); }
    }
    if (tr_on) fprintf(stderr, "[ps_trace] phase %u host rounds %.1f us (sums %.1f, scalars+parts %.1f, interp+transcript %.1f, ingest %.1f)\n", phase,
                       tr_us(t_rounds, tr_now()), tr_acc[0], tr_acc[1], tr_acc[2], tr_acc[3]);
    cp_id = bid;                                                                   // PrefixRegistry::update_checkpoints
    memcpy((void*)(p->h_v.data() + (size_t)phase * m), v_cur.data(), m * sizeof(FrH));
    c_cls[0] = mul(c_cls[0], v_cur[0]); c_cls[1] = mul(c_cls[1], v_cur[m - 1]);
  }
  // val = combine(prefix checkpoints, suffixes of the empty suffix) (mod.rs:527-552): suffixes [1, 0, 0, 1]
  {
    const FrH haz = cp[P_HAZ].has ? cp[P_HAZ].v : ja::host::FR_ONE, hao = cp[P_HAO].has ? cp[P_HAO].v : ja::host::FR_ONE;
    const FrH lw = cp[P_LW].has ? cp[P_LW].v : ja::host::FR_ZERO, msb = cp[P_MSB].has ? cp[P_MSB].v : ja::host::FR_ZERO;
    FrH val = sub(const_upper, mul(msb, lower_coeff));
    val = add(val, mul(haz, sub(lw, const_upper)));
    val = add(val, mul(hao, lw));
    const FrH raf_val = mul(gamma, cp_id);                                         // UnaryRafPS::raf_val (unary.rs:82-86)
    if (out_val) memcpy(out_val, val.l, 32);
    if (out_raf_val) memcpy(out_raf_val, raf_val.l, 32);
  }
  if (out_claim) memcpy(out_claim, claim.l, 32);
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
  return JA_OK;
}
// IdentityRCProver (joltworks/src/subprotocols/identity_range_check.rs:140-325): the LOG_K address rounds of the remainder range
// check - the unsigned identity prefix-suffix decomposition alone (poly/identity_poly.rs:113-166: suffixes [Shift, Identity],
// prefix polynomial  checkpoint * 2^chunk_len + i), no lookup table.  Same split as above: phase passes on the device, the
// m-entry rounds on the host; the prefix polynomial is affine in b, so a round is two plain sums and one index-weighted sum.
int32_t ja_psshout_prove_identity_rc(ja_ctx* c, ja_psshout* p, const uint64_t* claim_in, uint8_t state[32], uint32_t* n_rounds, uint64_t* out_coeffs,
                                     uint32_t* out_ncoeffs, uint64_t* out_challenges, uint64_t* out_input_claim, uint64_t* out_raf_val,
                                     uint64_t* out_claim) {
  JA_REQUIRE(c && p && state && n_rounds && out_coeffs && out_ncoeffs && out_challenges, "ja_psshout_prove_identity_rc: null argument");
  JA_REQUIRE(p->next_phase == 0, "ja_psshout_prove_identity_rc: the address rounds start from a fresh state (phase 0)");
  const unsigned log_m = p->log_m, phases = p->phases;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  const size_t m = size_t(1) << log_m;
  FrH pow2[66];
  pow2[0] = ja::host::FR_ONE;
  for (int i = 1; i < 66; i++) pow2[i] = dbl(pow2[i - 1]);
  const uint32_t kinds[2] = {JA_SUF_SHIFT, JA_SUF_IDENTITY};
  ja::host::Blake2bTranscript t(state, *n_rounds);
  FrH cp = ja::host::FR_ZERO;                                      // PrefixRegistry checkpoint of Prefix::Identity (None = 0)
  FrH claim = claim_in ? ja::host::from_limbs(claim_in) : ja::host::FR_ZERO;
  std::vector<FrH> Qbuf(2 * m), v_cur, v_next;
  p->h_v.assign((size_t)phases * m, ja::host::FR_ZERO);
  for (unsigned phase = 0; phase < phases; phase++) {
    int32_t st = ja_psshout_init_phase(c, p, phase, phase ? reinterpret_cast<const uint64_t*>(p->h_v.data() + (size_t)(phase - 1) * m) : nullptr,
                                       kinds, 2, 0, reinterpret_cast<uint64_t*>(Qbuf.data()));
    if (st) return st;
    FrH* Q0 = Qbuf.data();
    FrH* Q1 = Qbuf.data() + m;
    FrH bid = mul(cp, pow2[log_m]);                                  // identity_poly.rs:143-147: bound_value * 2^chunk_len + i
    v_cur.assign(1, ja::host::FR_ONE);
    for (unsigned tt = 0; tt < log_m; tt++) {
      const unsigned j = phase * log_m + tt, b_len = log_m - 1 - tt;
      const size_t half = size_t(1) << b_len;
      const FrH kappa = pow2[b_len];
      FrH s0[2], w0[2], s1[2];
      for (int sd = 0; sd < 2; sd++) {
        const size_t off = sd ? half : 0;
        ps_sum_wsum(Q0 + off, half, &s0[sd], &w0[sd]); s1[sd] = ps_sum(Q1 + off, half);
      }
      auto part = [&](uint32_t cc, int sd) {
        const FrH base = cc == 0 ? bid : (cc == 1 ? add(bid, kappa) : add(bid, dbl(kappa)));
        return add(add(mul(base, s0[sd]), w0[sd]), s1[sd]);
      };
      if (j == 0) {
        const FrH derived = add(part(0, 0), part(1, 1));
        if (out_input_claim) memcpy(out_input_claim, derived.l, 32);
        if (!claim_in) claim = derived;
      }
      const FrH e0 = part(0, 0);
      const FrH e2 = sub(dbl(part(2, 1)), part(2, 0));
      const ja::host::Coeffs uni = ja::host::from_evals_and_hint(claim, {e0, e2});
      const ja::host::Coeffs cpr = ja::host::compress(uni);
      JA_REQUIRE(cpr.size() <= 2, "ja_psshout_prove_identity_rc: round polynomial of degree > 2");
      t.append_message("UniPoly_begin");
      for (auto& x : cpr) t.append_scalar(x);
      t.append_message("UniPoly_end");
      uint64_t ch[4];
      t.challenge_scalar_optimized(ch);
      const FrH rj = ja::host::from_limbs(ch);
      claim = ja::host::evaluate(uni, rj);
      out_ncoeffs[j] = (uint32_t)cpr.size();
      for (size_t k = 0; k < 2; k++) memcpy(out_coeffs + 4 * (2 * j + k), k < cpr.size() ? cpr[k].l : ja::host::FR_ZERO.l, 32);
      memcpy(out_challenges + 4 * j, ch, 32);
      for (size_t b = 0; b < half; b++) {
        const FrH d0 = sub(Q0[b + half], Q0[b]), d1 = sub(Q1[b + half], Q1[b]);
        if (!d0.is_zero()) Q0[b] = add(Q0[b], ja::host::mul_chal(d0, rj));
        if (!d1.is_zero()) Q1[b] = add(Q1[b], ja::host::mul_chal(d1, rj));
      }
      bid = add(bid, ja::host::mul_chal(kappa, rj));
      v_next.resize(v_cur.size() * 2);
      for (size_t i = 0; i < v_cur.size(); i++) { const FrH e1 = ja::host::mul_chal(v_cur[i], rj); v_next[2 * i] = sub(v_cur[i], e1); v_next[2 * i + 1] = e1; }
      v_cur.swap(v_next);
    }
    cp = bid;
    memcpy((void*)(p->h_v.data() + (size_t)phase * m), v_cur.data(), m * sizeof(FrH));
  }
  if (out_raf_val) memcpy(out_raf_val, cp.l, 32);                   // identity_range_check.rs:316-319
  if (out_claim) memcpy(out_claim, claim.l, 32);
  memcpy(state, t.state, 32); *n_rounds = t.n_rounds;
  return JA_OK;
}
// the expanding tables of the last ja_psshout_prove_address (phases x m Fr)
int32_t ja_psshout_tables(ja_ctx* c, ja_psshout* p, uint64_t* out_v) {
  JA_REQUIRE(c && p && out_v, "ja_psshout_tables: null argument");
  JA_REQUIRE(p->h_v.size() == ((size_t)p->phases << p->log_m), "ja_psshout_tables: run ja_psshout_prove_address first");
  memcpy(out_v, p->h_v.data(), p->h_v.size() * sizeof(ja::host::FrH));
  return JA_OK;
}

void ja_psshout_free(ja_ctx* c, ja_psshout* p) {
  if (!c || !p) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  dev_free(c, p->d_idx); dev_free(c, p->d_u); dev_free(c, p->d_u2); dev_free(c, p->d_cols);
  delete p;
}

}  // extern "C"
