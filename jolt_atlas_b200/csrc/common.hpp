// Shared internals of the C-ABI translation units: handle structs, error mapping, launch geometry.
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/jolt_atlas_b200.h"
#include "fr_host.hpp"
#include "fp.cuh"
#include "ec.cuh"

#ifndef JA_COMMON_CONSTS
#define JA_COMMON_CONSTS
namespace ja {
constexpr int kSMs = 148;             // B200
constexpr int kBlock = 256;
}
#endif
using namespace ja;
using ja::host::FrH;

std::string& ja_err_slot();   // thread-local last-error text (capi.cu)

static inline int32_t fail(int32_t code, const std::string& msg) {
  ja_err_slot() = msg;
  return code;
}
#define JA_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(JA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
  } while (0)
#define JA_REQUIRE(cond, msg) \
  do { if (!(cond)) return fail(JA_ERR_INVALID, msg); } while (0)

// Split-eq prefix tables are read-only once built and depend on (order, point) only: the instances of a node (read-raf cycle rounds,
// RA checks, operator, range check) share r_node_output, so the context keeps the last few table pairs and hands them out by reference
// count (the reference's ProverOpeningAccumulator::eq_cycle_map caches the same thing for the opening reduction).
struct EqTableCacheEntry {
  std::vector<uint64_t> key;         // order, m, then the 4 m limbs of the point
  Fr* out_levels = nullptr;
  Fr* in_levels = nullptr;
  int refs = 0;
  uint64_t stamp = 0;
};
struct ja_ctx {
  int device = 0;
  std::vector<EqTableCacheEntry> eq_cache;
  uint64_t eq_cache_clock = 0;
  cudaStream_t stream = nullptr;
  std::recursive_mutex mu;
  Fr* d_partials = nullptr;        // kMaxGrid * kMaxOut
  unsigned int* d_counter = nullptr;
  Fr* d_out = nullptr;             // kMaxOut
  uint64_t* h_pinned = nullptr;    // staging for small D2H/H2D
  // Ring of pinned staging memory for small host->device uploads (eq points, tables, gammas): the copy is enqueued and
  // the call returns without synchronising.  A slot is only reused after kRingBytes of further uploads, by which time
  // the stream has long consumed it (every sumcheck round and every API call that returns data synchronises).
  char* h_ring = nullptr;
  size_t ring_off = 0;
  // host-mapped result slots of the sumcheck engine (sumcheck.cu): kSlots x kSlotBytes, sums at 0, sequence word at
  // kSlotSeqOffset; h_mapped / d_mapped are the host and device addresses of the same pinned allocation
  void* h_mapped = nullptr;
  void* d_mapped = nullptr;
  unsigned int seq = 0;
  // challenge mailboxes of the pre-launched round kernels (fused_kernels.cuh: MailRef): ring of kMailEntries 16-byte
  // entries in host-mapped memory + its device-memory twin; ahead_p / ahead_dev / ahead_tag are set only while the engine
  // enqueues the NEXT round's kernels
  void* h_mail = nullptr;
  void* d_mail = nullptr;
  void* d_mail_dev = nullptr;
  uint32_t mail_seq = 0;
  uint8_t mail_uses[64] = {0};
  const void* ahead_p = nullptr;
  void* ahead_dev = nullptr;
  uint32_t ahead_tag = 0;
  unsigned int* ahead_ticket = nullptr;   // relay-election word of the entry (behind the twins in d_mail_dev)
  uint32_t ahead_use = 0;
  // challenge channel of the round-resident kernels (persist_kernels.cuh): ring of kRrEntries 16-byte entries in host-mapped
  // memory (a call takes one entry per round, zeroed before its launch) + device-memory twins (block 0 relays into them)
  void* h_rrmail = nullptr;
  void* d_rrmail = nullptr;
  void* d_rrrelay = nullptr;
  uint32_t rr_off = 0;
  // A round-resident kernel occupies the stream until its last round: a call may run ONE of them (its only device-backed
  // instance, or the RaVirtual + Booleanity pair of an RA one-hot check); set per call by the sumcheck driver
  bool cache_openings = false;       // ja_set_cache_openings: the sumcheck drivers append every final claim to the transcript
  bool rr_call_ok = false, rr_call_pair = false;
  // flat host-mapped value array of the batched opening reduction (kMaxRowVals Fr + a sequence word at kRowSeqOffset)
  void* h_rowvals = nullptr;
  void* d_rowvals = nullptr;
  // MSM index-range shard of this context (shard.cu): ja_msm_run restricts every job to its slice when count > 1
  uint32_t msm_shard_index = 0, msm_shard_count = 1;
  // NCCL communicator of this context (comm.cu: ja_comm_init); with it, sharded MSMs / commitments / sumcheck rounds exchange
  // their partial results inside the library
  void* comm = nullptr;
  uint32_t comm_rank = 0, comm_world = 1;
  // sharded sumcheck (ja_set_sumcheck_shard): device polynomials are contiguous hypercube slices, partial round sums
  // are all-gathered through the caller's callback
  uint32_t sc_rank = 0, sc_world = 1;
  ja_allgather_fn sc_allgather = nullptr;
  void* sc_user = nullptr;
  bool slice_on = false;           // ja_round_eval_slice in progress: eq tables are indexed with g + slice_g_offset
  size_t slice_g_offset = 0;
  // Size-class cache of device buffers (dev_alloc / dev_free below).  Every buffer of a context is used on its one
  // stream, so a freed block can be handed out again immediately (stream order serialises the users).  A proof pass
  // allocates and frees thousands of polynomial buffers; in steady state none of that reaches the driver allocator,
  // whose pool growth showed up as sporadic 100-500 ms stalls (profiles/r1_pass_time_distribution.txt).
  std::unordered_map<void*, size_t> alloc_class;
  std::unordered_map<size_t, std::vector<void*>> free_blocks;
  // final-claim collector of the sumcheck engine: (device source, pinned host destination) of single field elements,
  // flushed by ONE small kernel per 32 entries that stores straight into the pinned staging buffer
  std::vector<std::pair<const void*, void*>> collect;
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // per-kernel-class CUDA-event profile (ja_profile_begin / ja_profile_end; bench.py's roofline leg)
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_events;     // pool, two per recorded launch
  std::vector<int> prof_cls;                // class of launch i (events 2i, 2i+1)
};

// Kernel classes of the launch profile.  Keep in sync with kClassNames (capi.cu).
enum {
  KC_BIND = 0, KC_ROUND_EVAL_S, KC_ROUND_EVAL_PROD, KC_ROUND_EVAL_DOT, KC_ROUND_SUM, KC_EQ_TABLE, KC_TENSOR_FOLD,
  KC_CONVERT, KC_MSM_SORT, KC_MSM_ACCUMULATE, KC_MSM_REDUCE, KC_ONEHOT_SUM, KC_HKZG_EVAL, KC_HKZG_LINCOMB, KC_HKZG_WITNESS, KC_SRS,
  KC_SUMCHECK_FUSED, KC_SCATTER, KC_MISC, KC_COUNT
};
void ja_prof_pre(ja_ctx* c, int cls);
void ja_prof_post(ja_ctx* c);
// every kernel launch of the library goes through this: counts it and, when profiling, brackets it with events
#define JA_LAUNCH(c, cls, ...) do { ja_prof_pre((c), (cls)); __VA_ARGS__; ja_prof_post((c)); } while (0)
static constexpr int kMaxGrid = kSMs * 8;
static constexpr int kMaxOut = 32;
static constexpr size_t kPinnedBytes = 1 << 22;
static constexpr size_t kRingBytes = 1 << 20;
static constexpr int kSlots = 128;
static constexpr size_t kSlotBytes = 2048, kSlotSeqOffset = 1024;
static constexpr size_t kMaxRowVals = 8192, kRowSeqOffset = kMaxRowVals * 32;
static constexpr uint32_t kMailEntries = 64;
static constexpr uint32_t kRrEntries = 4096;

struct ja_poly {
  size_t len = 0;
  Fr* buf[2] = {nullptr, nullptr};
  size_t cap[2] = {0, 0};
  int cur = 0;
  Fr* data() const { return buf[cur]; }
};

struct ja_srs {
  G1Aff* points = nullptr;   // g1_powers, affine Montgomery (kzg.rs:108-143 KZGProverKey)
  size_t n = 0;
  // Fixed-base window table (ja_srs_precompute): table[w * n + i] = 2^(c w) * g1_powers[i], w = 0..nwin-1 (table[0..n) is a
  // copy of the SRS).  With it a full-width MSM needs ONE bucket set instead of one per window and no doubling tail.
  // c = table_c is chosen per SRS size: wider windows mean fewer additions per scalar (nwin = ceil(255 / c)) but need
  // jobs long enough to fill 2^(c-1) buckets.
  G1Aff* table = nullptr;
  uint32_t table_c = 16, table_nwin = 16;
  // Large SRS (>= 2^21 points): a second, wider-window table for the long jobs, stored right behind the first one
  // (table[table2_off + w * n + i] = 2^(table2_c w) * g1_powers[i]); 0 = absent.  Measured on B200 (profiles/
  // r1_msm_table_probe_*.json): c = 20 takes a 2^24 MSM from 57.7 to 44.1 ms, but a 2^18 job is faster with c = 16.
  uint32_t table2_off = 0, table2_c = 0, table2_nwin = 0;
};
constexpr uint32_t kMaxFixedWindows = 22;

struct ja_onehot {
  uint64_t* d_indices = nullptr;        // concatenated base indices k*T + t of every list
  std::vector<uint64_t> offsets;       // count + 1
  uint64_t max_index = 0;
};

struct ja_tensor_i32 {
  int* data = nullptr;       // rows x cols, row-major, device
  size_t rows = 0, cols = 0;
};

struct ja_addr {
  uint32_t* d_k = nullptr;   // d lists x T addresses in [0, K), 0xFFFFFFFF = None
  size_t d = 0, T = 0, K = 0;
};

// one MSM of a batch (msm.cu): kind/nbits as in msm_kernels.cuh (0 = Fr Montgomery scalars, 254 bits)
struct MsmJob {
  const void* d_scalars; size_t n; uint32_t kind; uint32_t nbits; size_t base_offset;
  // MSM_FR only: the caller knows the scalars are pseudo-random field elements (quotient / folded polynomials of a HyperKZG
  // opening).  Their digits almost never collide inside a warp, so the digit passes use plain atomics instead of the
  // match_any aggregation (histogram of a 3 x 2^24 batch 9.8 -> 3.8 ms).  Skewed scalars stay correct, only slower.
  uint32_t dense_random = 0;
};
int32_t ja_msm_run(ja_ctx* c, const ja_srs* srs, const std::vector<MsmJob>& jobs, uint64_t* out_xy, int32_t* is_inf);
struct ja_psshout;
extern "C" int32_t ja_psshout_new_dev(ja_ctx* c, const unsigned long long* d_indices, size_t T, const uint64_t* r_cycle, size_t log_t, uint32_t log_k,
                                      uint32_t phases, ja_psshout** out);
// comm.cu: all-gather over the context's communicator (host buffers, rank-major) and the point-wise sum of every rank's partial points
int32_t comm_allgather(ja_ctx* c, const void* send, size_t bytes, void* recv);
int32_t comm_combine_points(ja_ctx* c, uint64_t* xy, int32_t* inf, size_t count);

struct ja_spliteq {
  int order = 0;
  int m = 0;
  int current_index = 0;
  FrH current_scalar;
  std::vector<FrH> w;
  // prefix tables: level k (2^k entries) at offset 2^k - 1
  Fr* out_levels = nullptr;
  Fr* in_levels = nullptr;
  int cache_slot = -1;               // >= 0: the tables belong to ja_ctx::eq_cache[cache_slot]
  int out_len = 1;   // E_out_vec.len()  (current table = level out_len-1)
  int in_len = 1;    // E_in_vec.len()
  const Fr* e_out() const { return out_levels + ((size_t(1) << (out_len - 1)) - 1); }
  const Fr* e_in() const { return in_levels + ((size_t(1) << (in_len - 1)) - 1); }
};

// ja_round_eval split at its synchronisation point (capi.cu); the sumcheck driver overlaps host math with the kernel
struct RoundEvalPending { size_t n_dev, n_out, n_polys; bool fam_sum; const uint64_t* aux_fr; int prod_lanes, prod_d; };
int32_t ja_round_eval_launch(ja_ctx* c, int32_t kernel_id, const ja_poly* const* polys, size_t n_polys,
                             const ja_spliteq* eq, const uint64_t* aux_fr, size_t n_aux, uint32_t aux_u32, size_t n_out,
                             RoundEvalPending* pend);
int32_t ja_round_eval_collect(ja_ctx* c, const RoundEvalPending& pend, uint64_t* out_evals);

// Host side of the tagged publication protocol (store_tagged, poly_kernels.cuh): element k of a host-mapped buffer is three
// self-validating 16-byte vectors [l0 l1 l2 tag] [l3 l4 l5 tag] [l6 l7 chk tag]; wait until all 3 n of them carry `tag` (and the
// element's xor checksum holds), then unpack n field elements into `out` (4 n limbs).  Bounded: a failed or finished stream and a 20 s timeout end the spin.
static inline int32_t wait_tagged(ja_ctx* c, const void* host_base, uint32_t tag, size_t n, uint64_t* out, const char* what) {
  const volatile uint32_t* h = reinterpret_cast<const volatile uint32_t*>(host_base);
  uint64_t spins = 0;
  const auto t0 = std::chrono::steady_clock::now();
  uint32_t* o = reinterpret_cast<uint32_t*>(out);
  for (size_t k = 0; k < n; k++) {
    const volatile uint32_t* q = h + 12 * k;
    uint32_t* dst = o + 8 * k;
    for (;;) {
      bool ok = q[3] == tag && q[7] == tag && q[11] == tag;
      if (ok) {
        __atomic_thread_fence(__ATOMIC_ACQUIRE);                   // payload words are read after the tags
        dst[0] = q[0]; dst[1] = q[1]; dst[2] = q[2]; dst[3] = q[4]; dst[4] = q[5]; dst[5] = q[6]; dst[6] = q[8]; dst[7] = q[9];
        ok = (dst[0] ^ dst[1] ^ dst[2] ^ dst[3] ^ dst[4] ^ dst[5] ^ dst[6] ^ dst[7] ^ tag) == q[10];
      }
      if (ok) break;
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      if ((++spins & 0xffff) == 0) {
        cudaError_t e = cudaStreamQuery(c->stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) return fail(JA_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
        if (e == cudaSuccess && !(q[3] == tag && q[7] == tag && q[11] == tag))
          return fail(JA_ERR_CUDA, std::string(what) + " finished without publishing its results");
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 20.0)
          return fail(JA_ERR_CUDA, std::string(what) + ": timed out waiting for the published results");
      }
    }
  }
  return JA_OK;
}
static inline uint32_t next_tag(ja_ctx* c) { if (++c->seq == 0) ++c->seq; return c->seq; }

static inline bool is_pow2(size_t n) { return n && !(n & (n - 1)); }
static inline int log2z(size_t n) { int k = 0; while ((size_t(1) << k) < n) k++; return k; }
static inline Challenge to_challenge(const uint64_t r[4]) {
  Challenge c;
  c.c[0] = (uint32_t)r[2]; c.c[1] = (uint32_t)(r[2] >> 32);
  c.c[2] = (uint32_t)r[3]; c.c[3] = (uint32_t)(r[3] >> 32);
  return c;
}
static inline Fr to_dev(const FrH& h) { Fr r; memcpy(r.l, h.l, 32); return r; }
static inline unsigned grid_for(size_t work) {
  size_t b = (work + kBlock - 1) / kBlock;
  if (b < 1) b = 1;
  if (b > (size_t)kSMs * 8) b = (size_t)kSMs * 8;
  return (unsigned)b;
}

// enqueue an H2D copy of a small host buffer through the pinned ring; no synchronisation
static inline int32_t stage_h2d(ja_ctx* c, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return JA_OK;
  JA_REQUIRE(bytes <= kRingBytes / 4, "stage_h2d: buffer too large for the staging ring");
  const size_t need = (bytes + 255) & ~size_t(255);
  if (c->ring_off + need > kRingBytes) c->ring_off = 0;
  char* slot = c->h_ring + c->ring_off;
  c->ring_off += need;
  memcpy(slot, src, bytes);
  JA_CUDA(cudaMemcpyAsync(dst, slot, bytes, cudaMemcpyHostToDevice, c->stream));
  return JA_OK;
}

static inline size_t dev_size_class(size_t bytes) {
  size_t s = 256;
  while (s < bytes) s <<= 1;
  if (s > (size_t(1) << 24)) {                 // above 16 MiB: 1/8-octave steps instead of powers of two (HBM is not free)
    const size_t step = s >> 4;
    s = (bytes + step - 1) / step * step;
  }
  return s;
}
static inline void dev_cache_release(ja_ctx* c) {
  for (auto& kv : c->free_blocks)
    for (void* p : kv.second) { c->alloc_class.erase(p); cudaFree(p); }
  c->free_blocks.clear();
}
static inline int32_t dev_alloc(ja_ctx* c, size_t bytes, void** out) {
  const size_t cls = dev_size_class(bytes ? bytes : 32);
  auto it = c->free_blocks.find(cls);
  if (it != c->free_blocks.end() && !it->second.empty()) {
    *out = it->second.back();
    it->second.pop_back();
    return JA_OK;
  }
  cudaError_t e = cudaMalloc(out, cls);
  if (e != cudaSuccess) {                      // out of memory: give the cached blocks back and retry once
    cudaGetLastError();
    cudaStreamSynchronize(c->stream);
    dev_cache_release(c);
    e = cudaMalloc(out, cls);
  }
  if (e != cudaSuccess) return fail(JA_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(cls) + "): " + cudaGetErrorString(e));
  c->alloc_class[*out] = cls;
  return JA_OK;
}
static inline void dev_free(ja_ctx* c, void* p) {
  if (!p) return;
  auto it = c->alloc_class.find(p);
  if (it == c->alloc_class.end()) { cudaFreeAsync(p, c->stream); return; }
  c->free_blocks[it->second].push_back(p);
}
