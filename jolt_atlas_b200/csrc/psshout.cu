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
// suffix.  Tile partials are added by a second small kernel.  Deterministic: field addition is exact in any order.
#include "common.hpp"
#include "poly_kernels.cuh"

namespace {

constexpr int kPsTile = 1024;        // entries per block
constexpr int kPsMaxSuffixes = 8;

// suffix MLEs of the clamp-table family (joltworks/src/lookup_tables/suffixes/{higher_all_zero,hzero_mul_lword,hone_mul_lword,one}.rs)
// and of the identity polynomial (poly/identity_poly.rs: the suffix value itself), in closed form: `bits` = the low `len` bits of
// the XLEN-bit index; the "higher" bits are those of significance >= 2^bound.
JA_DEV unsigned long long suffix_mle(uint32_t kind, unsigned long long bits, uint32_t len, uint32_t bound) {
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
    default: return bits;   // JA_SUF_IDENTITY
  }
}

struct PsPhaseArgs {
  const unsigned long long* idx; Fr* u; const Fr* v_prev;
  unsigned long long T;
  uint32_t prev_shift;             // k_bound of the previous phase = (k >> prev_shift) & m_mask
  uint32_t suffix_len, m_mask, bound, n_suf;
  uint32_t kinds[kPsMaxSuffixes];
  Fr* partial;                     // [tiles][n_suf][m]
};

__global__ void __launch_bounds__(256) k_ps_phase(const PsPhaseArgs a) {
  __shared__ Fr s_u[kPsTile];
  __shared__ unsigned long long s_suf[kPsTile];
  __shared__ unsigned short s_key[kPsTile];
  const size_t t0 = (size_t)blockIdx.x * kPsTile;
  const uint32_t m = a.m_mask + 1;
  for (uint32_t i = threadIdx.x; i < (uint32_t)kPsTile; i += blockDim.x) {
    const size_t j = t0 + i;
    if (j < a.T) {
      const unsigned long long k = a.idx[j];
      Fr u = fp_load(a.u + j);
      if (a.v_prev) {                                                   // init_phase: u_evals[j] *= v[phase - 1][k_bound]
        u = fp_mul<FrParams>(u, fp_load(a.v_prev + ((k >> a.prev_shift) & a.m_mask)));
        fp_store(a.u + j, u);
      }
      s_u[i] = u;
      s_suf[i] = a.suffix_len >= 64 ? k : (k & ((1ull << a.suffix_len) - 1));
      s_key[i] = (unsigned short)((a.suffix_len >= 64 ? 0ull : (k >> a.suffix_len)) & a.m_mask);
    } else {
      s_key[i] = 0xffff;
    }
  }
  __syncthreads();
  const uint32_t bin = threadIdx.x;
  Acc320 acc[kPsMaxSuffixes];
  Fr wide[kPsMaxSuffixes];            // suffix values beyond 32 bits (Identity): plain field accumulation
#pragma unroll
  for (int s = 0; s < kPsMaxSuffixes; s++) { acc[s] = acc320_zero(); wide[s] = fp_zero<FrParams>(); }
  if (bin < m) {
    for (int i = 0; i < kPsTile; i++) {
      if (s_key[i] != bin) continue;
      const Fr u = s_u[i];
      const unsigned long long sb = s_suf[i];
#pragma unroll
      for (int s = 0; s < kPsMaxSuffixes; s++) {
        if ((uint32_t)s >= a.n_suf) break;
        const unsigned long long t = suffix_mle(a.kinds[s], sb, a.suffix_len, a.bound);
        if (t == 0) continue;
        if (t >> 32) wide[s] = fp_add<FrParams>(wide[s], fp_mul_u64<FrParams>(u, t));
        else acc320_mad(acc[s], u, (uint32_t)t);
      }
    }
#pragma unroll
    for (int s = 0; s < kPsMaxSuffixes; s++) {
      if ((uint32_t)s >= a.n_suf) break;
      fp_store(a.partial + ((size_t)blockIdx.x * a.n_suf + s) * m + bin, fp_add<FrParams>(acc320_reduce(acc[s]), wide[s]));
    }
  }
}

__global__ void __launch_bounds__(256) k_ps_phase_final(const Fr* __restrict__ partial, uint32_t tiles, uint32_t n_out /* n_suf * m */, Fr* __restrict__ out) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_out) return;
  Fr tot = fp_zero<FrParams>();
  for (uint32_t b = 0; b < tiles; b++) tot = fp_add<FrParams>(tot, fp_load(partial + (size_t)b * n_out + id));
  fp_store(out + id, tot);
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
  size_t T = 0;
  uint32_t log_k = 0, phases = 0, log_m = 0;
  uint32_t next_phase = 0;
};

extern "C" {

int32_t ja_psshout_new(ja_ctx* c, const uint64_t* lookup_indices, size_t T, const uint64_t* r_cycle, size_t log_t, uint32_t log_k,
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
  JA_CUDA(cudaMemcpyAsync(p->d_idx, lookup_indices, T * 8, cudaMemcpyHostToDevice, c->stream));
  // u_evals = EqPolynomial::evals(r_node_output) (mod.rs:236)
  ja_poly* eq = nullptr;
  if ((st = ja_eq_evals(c, r_cycle, log_t, nullptr, &eq))) { dev_free(c, p->d_idx); delete p; return st; }
  p->d_u = eq->buf[eq->cur];
  eq->buf[eq->cur] = nullptr;
  ja_poly_free(c, eq);
  JA_CUDA(cudaStreamSynchronize(c->stream));            // the index array is borrowed for the duration of the call
  *out = p;
  return JA_OK;
}

int32_t ja_psshout_init_phase(ja_ctx* c, ja_psshout* p, uint32_t phase, const uint64_t* v_prev, const uint32_t* suffix_kinds, size_t n_suffixes,
                              uint32_t bound, uint64_t* out_Q) {
  JA_REQUIRE(c && p && suffix_kinds && out_Q && n_suffixes >= 1 && n_suffixes <= (size_t)kPsMaxSuffixes, "ja_psshout_init_phase: bad argument");
  JA_REQUIRE(phase == p->next_phase && phase < p->phases, "ja_psshout_init_phase: phases run in order 0 .. NUM_PHASES - 1");
  JA_REQUIRE((phase == 0) == (v_prev == nullptr), "ja_psshout_init_phase: the expanding table of the previous phase is required from phase 1 on");
  for (size_t s = 0; s < n_suffixes; s++) JA_REQUIRE(suffix_kinds[s] <= JA_SUF_IDENTITY, "ja_psshout_init_phase: unknown suffix kind");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  const uint32_t m = 1u << p->log_m;
  const uint32_t tiles = (uint32_t)((p->T + kPsTile - 1) / kPsTile);
  const size_t n_out = n_suffixes * m;
  Fr *d_v = nullptr, *d_part = nullptr, *d_out = nullptr;
  int32_t st;
  if ((st = dev_alloc(c, (size_t)tiles * n_out * sizeof(Fr), (void**)&d_part))) return st;
  if ((st = dev_alloc(c, n_out * sizeof(Fr), (void**)&d_out))) { dev_free(c, d_part); return st; }
  if (v_prev) {
    if ((st = dev_alloc(c, m * sizeof(Fr), (void**)&d_v))) { dev_free(c, d_part); dev_free(c, d_out); return st; }
    if ((st = stage_h2d(c, d_v, v_prev, m * sizeof(Fr)))) { dev_free(c, d_part); dev_free(c, d_out); dev_free(c, d_v); return st; }
  }
  PsPhaseArgs a;
  memset(&a, 0, sizeof(a));
  a.idx = p->d_idx; a.u = p->d_u; a.v_prev = d_v; a.T = p->T;
  a.prev_shift = (p->phases - phase) * p->log_m;                       // mod.rs:279: k.split((phases - phase) * log_m)
  a.suffix_len = (p->phases - 1 - phase) * p->log_m;                   // mod.rs:309
  a.m_mask = m - 1; a.bound = bound; a.n_suf = (uint32_t)n_suffixes;
  for (size_t s = 0; s < n_suffixes; s++) a.kinds[s] = suffix_kinds[s];
  a.partial = d_part;
  JA_LAUNCH(c, KC_SCATTER, k_ps_phase<<<tiles, 256, 0, c->stream>>>(a));
  JA_LAUNCH(c, KC_SCATTER, k_ps_phase_final<<<(unsigned)((n_out + 255) / 256), 256, 0, c->stream>>>(d_part, tiles, (uint32_t)n_out, d_out));
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_Q, d_out, n_out * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dev_free(c, d_part); dev_free(c, d_out); dev_free(c, d_v);
  if (e != cudaSuccess) return fail(JA_ERR_CUDA, std::string("ja_psshout_init_phase: ") + cudaGetErrorString(e));
  p->next_phase = phase + 1;
  return JA_OK;
}

int32_t ja_psshout_materialize_ra(ja_ctx* c, ja_psshout* p, const uint64_t* v, ja_poly** out_ra) {
  JA_REQUIRE(c && p && v && out_ra, "ja_psshout_materialize_ra: null argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  const uint32_t m = 1u << p->log_m;
  const size_t tab_bytes = (size_t)p->phases * m * sizeof(Fr);
  Fr* d_v = nullptr;
  int32_t st = dev_alloc(c, tab_bytes, (void**)&d_v);
  if (st) return st;
  if ((st = stage_h2d(c, d_v, v, tab_bytes))) { dev_free(c, d_v); return st; }
  if ((st = ja_poly_alloc(c, p->T, out_ra))) { dev_free(c, d_v); return st; }
  JA_LAUNCH(c, KC_CONVERT, k_ps_ra<<<grid_for(p->T), kBlock, 0, c->stream>>>(p->d_idx, p->T, d_v, p->phases, p->log_m, (*out_ra)->buf[0]));
  JA_CUDA(cudaGetLastError());
  dev_free(c, d_v);
  return JA_OK;
}

void ja_psshout_free(ja_ctx* c, ja_psshout* p) {
  if (!c || !p) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  dev_free(c, p->d_idx); dev_free(c, p->d_u);
  delete p;
}

}  // extern "C"
