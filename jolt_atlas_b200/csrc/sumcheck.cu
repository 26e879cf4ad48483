// The sumcheck engine: Sumcheck::prove (joltworks/src/subprotocols/sumcheck.rs:565-599) and BatchedSumcheck::prove
// (:30-184) with the per-round work on the device and the Fiat-Shamir transcript (blake2b.rs) on the host.
//
// Round structure (one host<->device exchange per round for ALL instances of a batch):
//   launch     every active instance enqueues ONE kernel: [bind the previous challenge +] evaluate this round
//              (fused_kernels.cuh); the finishing block publishes the reduced sums into host-mapped memory
//   overlap    while the kernels run the host does the round's field inversion (gruen_poly_deg_2/3 divide by
//              eq(1), finish_mles_product_sum by eq(0): both depend on the eq state only)
//   collect    the host spins on the slot's sequence word, assembles each instance's univariate (O(degree^2) glue with
//              cached interpolation matrices), batches them, compresses, appends to the transcript (unipoly.rs:550-558),
//              draws r_j (challenge_scalar_optimized) and evaluates the claims
//   ingest     host-side eq state update; the device bind is DEFERRED into the next round's kernel
// Instance kinds: the JA_EVAL_* round bodies on device polynomials, Booleanity (booleanity.rs; address rounds on the
// K-entry tables stay on the host, cycle rounds on the device) and the Hamming-weight instance over the G tables
// (hamming_weight.rs; K entries, host).  A Rust caller that owns the transcript can drive ja_round_eval / ja_bind_many
// itself; this driver is the same protocol with the library's transcript.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <memory>
#include <thread>

#include "common.hpp"
#include "fused_kernels.cuh"
#include "persist_kernels.cuh"
#include "tma_round.cuh"
#include "sumcheck_host.hpp"
#include "transcript_host.hpp"

using host::Coeffs;

extern "C" char** environ;      // tool-injection scan (ahead_allowed)

namespace {

// worker threads for the host-side glue of large batches: JA_HOST_THREADS, else min(8, hardware threads)
int host_threads() {
  static const int n = [] {
    if (const char* e = getenv("JA_HOST_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 256) return v; }
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(8u, hc ? hc : 1u));
  }();
  return n;
}

// Pre-launching a kernel that waits for the host deadlocks under anything that makes launches synchronous (ncu and
// compute-sanitizer serialise kernels, CUDA_LAUNCH_BLOCKING=1): detect tool injection / blocking launches once and
// fall back to plain launches.  JA_NO_AHEAD=1 forces the fallback, JA_AHEAD_TRACE=1 says on stderr which mode is on.
bool ahead_allowed() {
  static const bool ok = [] {
    bool allow = true;
    if (const char* e = getenv("CUDA_LAUNCH_BLOCKING")) if (atoi(e) != 0) allow = false;
    for (char** e = environ; allow && e && *e; e++) {
      static const char* const kTool[] = {"CUDA_INJECTION", "NV_NSIGHT_INJECTION", "NV_COMPUTE_PROFILER", "NV_TPS_LAUNCH", "NV_SANITIZER",
                                         "NVTX_INJECTION", "CUPTI_", "NSIGHT_"};
      for (const char* t : kTool) if (strncmp(*e, t, strlen(t)) == 0) { allow = false; break; }
    }
    if (getenv("JA_AHEAD_TRACE")) fprintf(stderr, "[jolt_atlas_b200] pre-launched round kernels: %s\n", allow ? "on" : "off (tool injection / blocking launches detected)");
    return allow;
  }();
  return ok;
}

// mailbox reference of the round being pre-launched (null reference outside a pre-launch)
static inline MailRef mail_ref(const ja_ctx* c) {
  MailRef m;
  m.p = reinterpret_cast<const uint4*>(c->ahead_p); m.dev = reinterpret_cast<uint4*>(c->ahead_dev); m.tag = c->ahead_tag;
  m.ticket = c->ahead_ticket; m.use = c->ahead_use;
  return m;
}

// JA_SC_TRACE=1: per-phase host wall-clock of the round loop on stderr (tuning aid)
struct ScTrace {
  bool on = getenv("JA_SC_TRACE") != nullptr;
  double t[6] = {0, 0, 0, 0, 0, 0};
  std::chrono::steady_clock::time_point last;
  void start() { if (on) last = std::chrono::steady_clock::now(); }
  void lap(int k) { if (!on) return; auto n = std::chrono::steady_clock::now(); t[k] += std::chrono::duration<double, std::micro>(n - last).count(); last = n; }
};
ScTrace g_trace;

// ---- TMA-staged streaming round kernels (tma_round.cuh) -------------------------------------------------------------
// cuTensorMapEncodeTiled comes from the driver through the runtime's entry-point query: nothing links libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// a polynomial of n Fr as [n/4 rows][32 x u32], 32-row boxes (one warp's slab), SWIZZLE_128B
int32_t make_row_map(CUtensorMap* tm, const Fr* base, size_t n) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(JA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {32, (cuuint64_t)(n / 4)};
  const cuuint64_t gstride[1] = {128};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<Fr*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(JA_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return JA_OK;
}
constexpr size_t kTmaMinPairs = size_t(1) << 16;     // below this the slabs live in L2 and the register-staged kernel is as fast
bool tma_eligible(int kind, bool fused, size_t G) {
  const bool off = getenv("JA_NO_TMA") != nullptr;      // read per call: tests flip it to compare the two kernels
  return !off && fused && (kind == JA_EVAL_ADD || kind == JA_EVAL_SUB || kind == JA_EVAL_IDENT) && G >= kTmaMinPairs && G % kTmaRows == 0;
}
// in[q]: current arrays of 4G Fr, out[q]: bound arrays of 2G Fr
int32_t launch_round_s_tma(ja_ctx* c, int kind, const FusedPolys& P, const Challenge& ch, const Fr* e_out, const Fr* e_in, int bits_in,
                           size_t G, Fr* part, unsigned int* ctr, const Publish& pub, size_t g_off) {
  static bool attr_set = false;
  if (!attr_set) {
    JA_CUDA(cudaFuncSetAttribute(k_round_s_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmemBytes));
    JA_CUDA(cudaFuncSetAttribute(k_round_s_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmemBytes));
    JA_CUDA(cudaFuncSetAttribute(k_round_s_tma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmemBytes));
    attr_set = true;
  }
  const int np = kind == JA_EVAL_IDENT ? 1 : 2;
  CUtensorMap tm[2];
  int32_t st;
  for (int q = 0; q < 2; q++)
    if ((st = make_row_map(&tm[q], P.in[q < np ? q : 0], 4 * G))) return st;
  const size_t slabs = G / kTmaRows;
  size_t grid = slabs < (size_t)kSMs * 2 ? slabs : (size_t)kSMs * 2;
  const size_t spb = (slabs + grid - 1) / grid;
  grid = (slabs + spb - 1) / spb;
  cudaStream_t s = c->stream;
#define JA_TMA_K(KID) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_s_tma<KID><<<(unsigned)grid, kTmaRows, kTmaSmemBytes, s>>>( \
    tm[0], tm[1], P.out[0], P.out[np - 1], ch, e_out, e_in, bits_in, G, spb, part, ctr, pub, g_off))
  if (kind == JA_EVAL_ADD) JA_TMA_K(0);
  else if (kind == JA_EVAL_SUB) JA_TMA_K(1);
  else JA_TMA_K(6);
#undef JA_TMA_K
  return JA_OK;
}

// final-claim collector (ja_ctx::collect): one launch per 32 (source, pinned destination) pairs
int32_t flush_collect(ja_ctx* c) {
  for (size_t base = 0; base < c->collect.size(); base += 32) {
    CollectArgs a;
    const int n = (int)std::min<size_t>(32, c->collect.size() - base);
    for (int i = 0; i < 32; i++) {
      a.src[i] = i < n ? reinterpret_cast<const Fr*>(c->collect[base + i].first) : nullptr;
      a.dst[i] = i < n ? reinterpret_cast<Fr*>(c->collect[base + i].second) : nullptr;
    }
    JA_LAUNCH(c, KC_BIND, k_collect_finals<<<1, 64, 0, c->stream>>>(a, n));
  }
  c->collect.clear();
  JA_CUDA(cudaGetLastError());
  return JA_OK;
}

// ---- host-mapped result slots ------------------------------------------------------------------------------------
struct Slot {
  const uint64_t* host_vals = nullptr;
  volatile unsigned int* host_seq = nullptr;
  Publish pub;
  bool armed = false;
};
Slot arm_slot(ja_ctx* c, int s) {
  Slot r;
  char* h = reinterpret_cast<char*>(c->h_mapped) + (size_t)s * kSlotBytes;
  char* d = reinterpret_cast<char*>(c->d_mapped) + (size_t)s * kSlotBytes;
  r.host_vals = reinterpret_cast<const uint64_t*>(h);
  r.host_seq = reinterpret_cast<volatile unsigned int*>(h + kSlotSeqOffset);
  r.pub.vals = reinterpret_cast<Fr*>(d);
  r.pub.seq = reinterpret_cast<volatile unsigned int*>(d + kSlotSeqOffset);
  if (++c->seq == 0) ++c->seq;                                     // 0 = "no tag" on the device side
  r.pub.value = c->seq;
  r.armed = true;
  return r;
}
int32_t wait_slot(ja_ctx* c, const Slot& s) {
  uint64_t spins = 0;
  auto t0 = std::chrono::steady_clock::now();
  while (*s.host_seq != s.pub.value) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
    if ((++spins & 0xffff) == 0) {
      // a failed launch / faulted kernel never publishes: surface the CUDA error instead of spinning forever
      cudaError_t e = cudaStreamQuery(c->stream);
      if (e != cudaSuccess && e != cudaErrorNotReady) return fail(JA_ERR_CUDA, std::string("sumcheck round kernel: ") + cudaGetErrorString(e));
      if (e == cudaSuccess && *s.host_seq != s.pub.value) return fail(JA_ERR_CUDA, "sumcheck round kernel finished without publishing its sums");
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 20.0)
        return fail(JA_ERR_CUDA, "sumcheck round kernel: timed out waiting for the published sums");
    }
  }
  return JA_OK;
}

// Tagged protocol (store_tagged, poly_kernels.cuh): element k of the slot is three self-validating 16-byte vectors; wait
// until all 3 n of them carry the round's tag, then unpack n field elements into `out` (4 n limbs).
int32_t wait_slot_tagged(ja_ctx* c, const Slot& s, size_t n, uint64_t* out) {
  return wait_tagged(c, s.host_vals, s.pub.value, n, out, "sumcheck round kernel");
}

// ---- round-resident kernels (persist_kernels.cuh): host side of the per-call challenge channel ---------------------------------
// JA_NO_PERSIST=1 keeps every round on the per-round kernels (tests compare the two paths); tool injection / blocking launches
// do the same (a kernel that waits for the host cannot run under a tool that serialises or replays launches).
bool persist_allowed() { return ahead_allowed() && getenv("JA_NO_PERSIST") == nullptr; }
constexpr size_t kRrMaxLenS = size_t(1) << 16;      // family S: longer arrays start on the per-round kernels (TMA-staged from 2^17) and hand over
constexpr size_t kRrMaxLenDot = size_t(1) << 12;    // one block
constexpr size_t kRrMaxLenPair = size_t(1) << 15;   // RaVirtual + Booleanity pair
struct PersistGroup {
  volatile uint32_t* h_entries = nullptr;           // host view of this call's mailbox entries (one per round)
  const uint4* d_entries = nullptr;
  uint4* d_relay = nullptr;
  int rounds = 0, posted = 0;
  unsigned int tag0 = 0;                            // round i publishes with tag0 + i
  int32_t reserve(ja_ctx* c, int n_rounds) {
    JA_REQUIRE(n_rounds >= 1 && n_rounds <= kRrMaxRounds, "sumcheck: round-resident kernel: bad round count");
    rounds = n_rounds;
    if (c->rr_off + (uint32_t)rounds > kRrEntries) c->rr_off = 0;
    const uint32_t off = c->rr_off;
    c->rr_off += (uint32_t)rounds;
    volatile uint32_t* e = reinterpret_cast<volatile uint32_t*>(reinterpret_cast<char*>(c->h_rrmail) + 16 * (size_t)off);
    for (int i = 0; i < 4 * rounds; i++) e[i] = 0;
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    d_entries = reinterpret_cast<const uint4*>(c->d_rrmail) + off;
    d_relay = reinterpret_cast<uint4*>(c->d_rrrelay) + off;
    JA_CUDA(cudaMemsetAsync(d_relay, 0, 16 * (size_t)rounds, c->stream));
    if (c->seq > 0xffffff00u) c->seq = 0;            // a call's tags are consecutive and never 0
    tag0 = c->seq + 1;
    c->seq += (unsigned int)rounds + 2;               // one tag per round and one for the final claims (per instance of a pair)
    h_entries = e;
    return JA_OK;
  }
  void post(size_t i, const uint64_t ch[4]) {
    if (!h_entries || i >= (size_t)rounds || (int)i < posted) return;
    const Challenge cc = to_challenge(ch);
    volatile uint32_t* e = h_entries + 4 * i;
    e[0] = cc.c[0]; e[1] = cc.c[1]; e[2] = cc.c[2];
    __atomic_thread_fence(__ATOMIC_RELEASE);
    e[3] = (cc.c[3] & 0x1fffffffu) | 0x20000000u;    // bit 29: valid
    posted = (int)i + 1;
  }
  ~PersistGroup() {                                  // error paths: the kernel must never be left waiting
    if (!h_entries) return;
    for (int i = posted; i < rounds; i++) { __atomic_thread_fence(__ATOMIC_RELEASE); h_entries[4 * i + 3] = 0x80000000u; }
  }
};
static inline RrEq rr_eq_state(const ja_spliteq* e) {
  RrEq q;
  q.out_levels = e->out_levels; q.in_levels = e->in_levels;
  q.out_len = e->out_len; q.in_len = e->in_len; q.ci = e->current_index; q.m = e->m;
  return q;
}
template <class ARGS>
static inline int32_t rr_launch(ja_ctx* c, const void* kernel, unsigned int grid, ARGS& args) {
  void* params[] = {&args};
  cudaError_t e = cudaSuccess;
  JA_LAUNCH(c, KC_SUMCHECK_FUSED, e = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kBlock), params, 0, c->stream));
  if (e != cudaSuccess) return fail(JA_ERR_CUDA, std::string("round-resident kernel launch: ") + cudaGetErrorString(e));
  return JA_OK;
}

// ---- GruenSplitEqPolynomial on the host for K-entry address rounds (split_eq_poly.rs:86-145,331-372; LowToHigh) ----
struct HostEq {
  std::vector<FrH> w;
  size_t ci = 0;
  FrH scalar = host::FR_ONE;
  void init(const uint64_t* limbs, size_t m) { w.resize(m); for (size_t i = 0; i < m; i++) w[i] = host::from_limbs(limbs + 4 * i); ci = m; }
  FrH current_w() const { return w[ci - 1]; }
  void bind(const FrH& r) {
    const FrH wv = current_w(), prod = host::mul(wv, r);
    FrH f = host::sub(host::sub(host::FR_ONE, wv), r);
    f = host::add(host::add(f, prod), prod);
    scalar = host::mul(scalar, f);
    ci--;
  }
  // E_out (x) E_in over the variables still to be bound after the current one: eq(w[0 .. ci-1), .), w[0] = MSB
  std::vector<FrH> table() const {
    std::vector<FrH> t{host::FR_ONE};
    for (size_t j = 0; j + 1 < ci; j++) {
      std::vector<FrH> n(t.size() * 2);
      for (size_t i = 0; i < t.size(); i++) { n[2 * i + 1] = host::mul(t[i], w[j]); n[2 * i] = host::sub(t[i], n[2 * i + 1]); }
      t.swap(n);
    }
    return t;
  }
};

Coeffs scaled(const Coeffs& c, const FrH& s) {        // &UniPoly * F (unipoly.rs:455-461): from_coeff trims
  Coeffs o(c.size());
  for (size_t i = 0; i < c.size(); i++) o[i] = host::mul(c[i], s);
  return host::trim(o);
}
void add_assign(Coeffs& a, const Coeffs& b) {         // unipoly.rs:401-412
  for (size_t i = 0; i < a.size() && i < b.size(); i++) a[i] = host::add(a[i], b[i]);
  if (a.size() < b.size()) a.insert(a.end(), b.begin() + a.size(), b.end());
}
FrH mul_pow_2(FrH x, size_t k) { for (size_t i = 0; i < k; i++) x = host::dbl(x); return x; }   // field/mod.rs:274-284

// ---- instances -------------------------------------------------------------------------------------------------------
struct Inst {
  size_t rounds = 0;
  FrH claim = host::FR_ZERO;
  uint64_t* out_final = nullptr;
  virtual struct DevInst* pair_candidate(size_t /*local round*/) { return nullptr; }
  int slot_id = -1;                 // host-mapped result slot (assigned by run() to the instances that publish from a kernel)
  virtual bool needs_slot() const { return false; }
  virtual ~Inst() {}
  virtual int32_t launch(ja_ctx* c, size_t round) = 0;
  // host work that depends on the instance state only (the round's field inversion): runs after EVERY instance of the
  // batch has enqueued its kernel, i.e. overlapped with the kernels
  virtual int32_t prework(ja_ctx*, size_t) { return JA_OK; }
  virtual int32_t message(ja_ctx* c, size_t round, const FrH& prev, Coeffs* uni) = 0;
  virtual int32_t ingest(ja_ctx* c, const uint64_t ch[4], size_t round) = 0;
  // flush deferred work and enqueue the D2H of the final claims into `staging` (pinned); *count = number of Fr written
  virtual int32_t finalize(ja_ctx* c, uint64_t* staging, size_t* count) = 0;
  virtual void release(ja_ctx*) {}
  // Pre-launch protocol (prove_loop): the kernels of local round `next` are enqueued while round next-1 is still in flight
  // and receive its challenge through the context's mailbox.  can_ahead: this instance's round `next` may be enqueued
  // that way (host-only instances: always).  ahead_begin / ahead_end bracket the early launch (state as if the missing
  // challenge had been ingested, then back); ahead_commit runs at the top of the next iteration instead of launch().
  // the round's challenge, right after the transcript squeeze (round-resident kernels take it from their mailbox)
  virtual void post_challenge(const uint64_t* /*ch*/, size_t /*local round*/) {}
  virtual void abort_persist() {}          // error paths: release a round-resident kernel that is still waiting for challenges
  virtual bool can_ahead(size_t /*next local round*/) { return false; }
  virtual int32_t ahead_begin(ja_ctx*) { return JA_OK; }
  virtual void ahead_end(ja_ctx*) {}
  virtual void ahead_commit() {}
};

// device-resident polynomials, one JA_EVAL_* body (or booleanity phase 2 = body 7)
struct DevInst : Inst {
  int32_t kind = 0;
  std::vector<ja_poly*> polys;
  ja_spliteq* eq = nullptr;
  bool own_eq = false;
  std::vector<FrH> gammas;          // SUM1 (host combination) / body 7 (device copy in d_gammas)
  Fr* d_gammas = nullptr;
  uint32_t pow_d = 0;
  int order = JA_LOW_TO_HIGH;
  size_t n_out = 0;
  bool fusable = false, pending = false;
  uint64_t pend_ch[4] = {0, 0, 0, 0};
  FrH scale = host::FR_ONE;         // body 7: eq_r_r (booleanity.rs:295-300)
  bool has_scale = false;
  FrH scale_inv = host::FR_ONE;
  // per-round state between launch and message
  Slot slot;
  // pre-launch: slot armed for the NEXT round's kernel, eq level saved across the early launch, "the pending bind of the
  // next ingest has already been consumed"
  Slot slot_next, slot_keep;
  bool ahead_consumed = false;
  struct EqSave { int current_index = 0; size_t in_len = 0, out_len = 0; FrH scalar; } eq_save;
  RoundEvalPending pend;
  bool legacy = false;              // round went through ja_round_eval_launch (pinned staging + stream sync)
  FrH cs, cw, div;
  // Division without a per-round inversion: w_inv[i] = 1 / w_i (split-eq bodies) or 1 / (1 - w_i) (product bodies) for every
  // coordinate of the eq point, ONE batch inversion at the instance's first round; the split-eq bodies also carry
  // nclaim = claim / current_scalar = q(r) of the eq-free polynomial q of the previous round (gruen_poly_deg_*_q1).
  std::vector<FrH> w_inv;
  FrH nclaim = host::FR_ZERO;
  bool nclaim_valid = false;
  FrH qc0 = host::FR_ZERO, qc1 = host::FR_ZERO, qc2 = host::FR_ZERO;   // q(X) = qc0 + qc1 X + qc2 X^2 of the last message
  int prod_lanes = 0;
  // multi-GPU: the polynomials are this rank's contiguous hypercube slices (ja_set_sumcheck_shard)
  bool sharded = false, round_sharded = false;
  uint32_t sc_rank = 0, sc_world = 1;
  // round-resident kernel (persist_kernels.cuh) running the remaining rounds of this instance
  std::shared_ptr<PersistGroup> pg;
  size_t pg_round0 = ~size_t(0);    // local round of the kernel's first round
  bool pg_flip = false;             // the final claims end up in the OTHER ping-pong buffer
  unsigned int pg_tag_off = 0;      // this instance's tag range inside the group's (pairs: product first, booleanity second)
  bool c_rr_ok = false;             // the call's shape allows a round-resident kernel (ja_ctx::rr_call_ok at set-up) for its one device instance
  bool c_rr_pair = false;           // ... for the RaVirtual + Booleanity pair

  int32_t setup(ja_ctx* c) {
    const size_t len = polys[0]->len;
    for (ja_poly* p : polys) JA_REQUIRE(p && p->len == len, "sumcheck: polynomials of one instance must have equal length");
    JA_REQUIRE(len >= 2 && is_pow2(len), "sumcheck: polynomial length must be a power of two >= 2");
    rounds = (size_t)log2z(len);
    c_rr_ok = c->rr_call_ok; c_rr_pair = c->rr_call_pair;
    switch (kind) {
      case JA_EVAL_ADD: case JA_EVAL_SUB: n_out = 1; JA_REQUIRE(polys.size() == 2, "sumcheck: ADD/SUB take two polynomials"); fusable = true; break;
      case JA_EVAL_IDENT: n_out = 1; JA_REQUIRE(polys.size() == 1, "sumcheck: IDENT takes one polynomial"); fusable = true; break;
      case JA_EVAL_MUL: n_out = 2; JA_REQUIRE(polys.size() == 2, "sumcheck: MUL takes two polynomials"); fusable = true; break;
      case JA_EVAL_SQUARE: n_out = 2; JA_REQUIRE(polys.size() == 1, "sumcheck: SQUARE takes one polynomial"); fusable = true; break;
      case 7: n_out = 2; fusable = true; break;
      case JA_EVAL_IFF: n_out = 2; JA_REQUIRE(polys.size() == 3, "sumcheck: IFF takes three polynomials (mask, a, b)"); fusable = true; break;
      case JA_EVAL_DIV: n_out = 2; JA_REQUIRE(polys.size() == 4, "sumcheck: DIV takes four polynomials (l, r, q, R)"); fusable = true; break;
      case JA_EVAL_RSQRT: n_out = 2; JA_REQUIRE(polys.size() == 5 && gammas.size() == 2, "sumcheck: RSQRT takes five polynomials and aux = {gamma, S^3}"); fusable = true; break;
      case JA_EVAL_LIN3: n_out = 1; JA_REQUIRE(polys.size() == 3 && gammas.size() == 1, "sumcheck: LIN3 takes three polynomials and aux = {tau}"); fusable = true; break;
      case JA_EVAL_PROD: n_out = polys.size(); fusable = polys.size() <= 16; break;
      case JA_EVAL_POW: n_out = pow_d; JA_REQUIRE(polys.size() == 1, "sumcheck: POW takes one polynomial"); fusable = pow_d <= 16; break;
      case JA_EVAL_DOT2: n_out = 2; order = JA_HIGH_TO_LOW; JA_REQUIRE(polys.size() == 2, "sumcheck: DOT2 takes two polynomials"); fusable = true; break;
      case JA_EVAL_DOT3: n_out = 3; order = JA_HIGH_TO_LOW; JA_REQUIRE(polys.size() == 3, "sumcheck: DOT3 takes three polynomials"); fusable = true; break;
      case JA_EVAL_OPEN: n_out = 1; order = JA_HIGH_TO_LOW; JA_REQUIRE(polys.size() == 1, "sumcheck: OPEN takes one polynomial"); fusable = true; break;
      case JA_EVAL_SUM1: n_out = 1; break;
      case JA_EVAL_SUMHI: n_out = 1; order = JA_HIGH_TO_LOW; JA_REQUIRE(polys.size() == 1, "sumcheck: SUMHI takes one polynomial"); break;
      default: return fail(JA_ERR_UNSUPPORTED, "sumcheck: kind not implemented");
    }
    JA_REQUIRE(n_out >= 1 && n_out <= (size_t)kMaxOut && polys.size() <= (size_t)kMaxProdPolys, "sumcheck: too many polynomials / outputs");
    if (c->sc_world > 1 && c->sc_allgather && kind != 7) {
      JA_REQUIRE(fusable && order == JA_LOW_TO_HIGH, "sumcheck: only the LowToHigh split-eq / product bodies run on hypercube slices");
      sharded = true; sc_rank = c->sc_rank; sc_world = c->sc_world;
      rounds += (size_t)log2z(sc_world);
    }
    if (kind == 7 || kind == JA_EVAL_RSQRT || kind == JA_EVAL_LIN3) {
      JA_REQUIRE(kind != 7 || gammas.size() == polys.size(), "sumcheck: booleanity takes one gamma per polynomial");
      int32_t st = dev_alloc(c, gammas.size() * sizeof(Fr), (void**)&d_gammas);
      if (st) return st;
      if ((st = stage_h2d(c, d_gammas, gammas.data(), gammas.size() * sizeof(Fr)))) return st;
    }
    return JA_OK;
  }
  bool uses_eq() const { return kind <= JA_EVAL_LIN3 || kind == JA_EVAL_OPEN; }
  bool needs_slot() const override { return fusable; }

  // everything a fused round launch needs, computed before the launch (shared by the single and the paired launch)
  struct Prep {
    FusedPolys P;
    bool fz = false;
    size_t len_eval = 0, G = 0;
    Challenge ch;
    int bits_in = 0;
    const Fr* e_out = nullptr;
    const Fr* e_in = nullptr;
    size_t g_off = 0;
  };
  int32_t prepare(ja_ctx* c, Prep* pr) {
    const bool fz = pending;
    const size_t len_in = polys[0]->len;
    const size_t len_eval = fz ? len_in / 2 : len_in;      // length of the arrays this round evaluates
    const size_t G = len_eval / 2;
    FusedPolys& P = pr->P;
    for (size_t q = 0; q < polys.size(); q++) {
      ja_poly* p = polys[q];
      P.in[q] = p->data();
      P.out[q] = p->data();
      if (fz && order == JA_LOW_TO_HIGH) {
        const int nxt = 1 - p->cur;
        if (p->cap[nxt] < len_eval) {
          dev_free(c, p->buf[nxt]);
          p->buf[nxt] = nullptr; p->cap[nxt] = 0;
          int32_t st = dev_alloc(c, len_eval * sizeof(Fr), (void**)&p->buf[nxt]);
          if (st) return st;
          p->cap[nxt] = len_eval;
        }
        P.out[q] = p->buf[nxt];
      }
    }
    pr->ch = to_challenge(pend_ch);
    if (uses_eq()) {
      JA_REQUIRE(eq && eq->order == order, "sumcheck: split-eq binding order does not match the round body");
      const size_t cover = size_t(1) << ((eq->out_len - 1) + (eq->in_len - 1));
      JA_REQUIRE(cover == G * (sharded ? sc_world : 1), "sumcheck: split-eq tables do not cover len/2 (eq and polys out of lockstep)");
      pr->bits_in = eq->in_len - 1; pr->e_out = eq->e_out(); pr->e_in = eq->e_in();
    }
    pr->g_off = sharded ? (size_t)sc_rank * G : 0;
    pr->fz = fz; pr->len_eval = len_eval; pr->G = G;
    return JA_OK;
  }
  void commit_round(const Prep& pr) {
    if (pr.fz) {
      for (ja_poly* p : polys) { if (order == JA_LOW_TO_HIGH) p->cur = 1 - p->cur; p->len = pr.len_eval; }
      pending = false;
    }
  }
  bool pairable() const {
    return !pg && !sharded && fusable && polys.size() >= 2 && polys.size() <= 16 && (kind == JA_EVAL_PROD || kind == 7);
  }
  // this instance's half of a paired launch (k_round_prod_bool): slot armed, buffers ready, state advanced
  int32_t prepare_pair(ja_ctx* c, PairArgs* a, Prep* pr, int scratch_half) {
    int32_t st = prepare(c, pr);
    if (st) return st;
    slot = arm_slot(c, slot_id);
    legacy = false;
    const int d = (int)polys.size();
    int L = 2; while (L < d) L <<= 1;
    // small slabs: 128-thread blocks, one pass per block (the product of 9..16 factors on 64 threads per pair)
    const bool no_small = getenv("JA_NO_WIDE") != nullptr;                     // JA_NO_WIDE=1: tests compare the variants
    // JA_BIGWIDE=1 (experiment, off): the 64-threads-per-pair form on LARGE slabs too.  Measured on B200 it loses - 600 vs 394 us at
    // 2^16 pairs: both forms sit at ~0.72 of the field-mul peak there, and the wide one does 1.6x the products.
    const bool big_wide = !no_small && L == 16 && pr->G >= kBigWideMinPairs && getenv("JA_BIGWIDE") != nullptr;
    small_round = big_wide || (!no_small && pr->G <= kSmallMaxPairs);
    wide_round = big_wide || (small_round && L == 16 && pr->G <= kWideMaxPairs);
    size_t ppb;
    if (big_wide) {
      ppb = kind == JA_EVAL_PROD ? big_wide_ppb_prod(pr->G) : big_wide_ppb_bool(pr->G);
    } else if (small_round) {
      ppb = (kind == JA_EVAL_PROD && wide_round) ? (size_t)kWideBlock / 64 : (size_t)kWideBlock / L;
    } else {
      const size_t gpb = (size_t)kBlock / L;
      ppb = (pr->G + (size_t)kSMs * 2 - 1) / ((size_t)kSMs * 2);
      ppb = (ppb + gpb - 1) / gpb * gpb;
    }
    a->P = pr->P; a->d = d; a->e_out = pr->e_out; a->e_in = pr->e_in; a->bits_in = pr->bits_in;
    a->nb = (unsigned int)((pr->G + ppb - 1) / ppb); a->G = pr->G; a->ppb = ppb;
    a->gammas = d_gammas;
    a->partials = c->d_partials + (size_t)scratch_half * (kMaxGrid * kMaxOut / 2);
    a->counter = c->d_counter + scratch_half;
    a->pub = slot.pub;
    if (kind == JA_EVAL_PROD) prod_lanes = L;
    commit_round(*pr);
    return JA_OK;
  }

  // Every slice is down to one coefficient: gather the sc_world remaining coefficients of each MLE onto every rank and
  // run the last log2(world) rounds replicated ("the last log2 g rounds run after gathering g elements per poly").
  int32_t unshard(ja_ctx* c) {
    int32_t st;
    if (pending) {
      if ((st = ja_bind_many(c, polys.data(), polys.size(), pend_ch, order))) return st;
      pending = false;
    }
    const size_t np = polys.size();
    std::vector<uint64_t> mine(4 * np), all(4 * np * sc_world);
    for (size_t q = 0; q < np; q++) {
      JA_REQUIRE(polys[q]->len == 1, "sumcheck: slice not reduced to one coefficient at the hand-over");
      JA_CUDA(cudaMemcpyAsync(c->h_pinned + 4 * q, polys[q]->data(), sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
    }
    JA_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(mine.data(), c->h_pinned, 32 * np);
    if (c->sc_allgather(c->sc_user, mine.data(), 32 * np, all.data()) != 0) return fail(JA_ERR_INVALID, "sumcheck: allgather callback failed");
    std::vector<uint64_t> col(4 * sc_world);
    for (size_t q = 0; q < np; q++) {
      ja_poly* p = polys[q];
      for (uint32_t r = 0; r < sc_world; r++) memcpy(col.data() + 4 * r, all.data() + 4 * (r * np + q), 32);
      if (p->cap[p->cur] < sc_world) {
        dev_free(c, p->buf[p->cur]);
        p->buf[p->cur] = nullptr; p->cap[p->cur] = 0;
        if ((st = dev_alloc(c, sc_world * sizeof(Fr), (void**)&p->buf[p->cur]))) return st;
        p->cap[p->cur] = sc_world;
      }
      if ((st = stage_h2d(c, p->buf[p->cur], col.data(), 32 * sc_world))) return st;
      p->len = sc_world;
    }
    sharded = false;
    return JA_OK;
  }

  bool can_ahead(size_t next) override {
    if (!fusable || sharded || next < 1 || next >= rounds) return false;
    {
      const size_t next_len = (pending ? polys[0]->len / 2 : polys[0]->len) / 2;
      if (pg || persist_ok(next_len)) return false;                    // that round runs in (or starts) a round-resident kernel
      if ((kind == JA_EVAL_PROD || kind == 7) && c_rr_pair && !sharded && polys.size() <= 16 && next_len >= 2 && next_len <= kRrMaxLenPair &&
          persist_allowed()) return false;                             // ... the pair's
    }
    switch (kind) {
      case JA_EVAL_ADD: case JA_EVAL_SUB: case JA_EVAL_MUL: case JA_EVAL_SQUARE: case JA_EVAL_IDENT: case 7:
      case JA_EVAL_IFF: case JA_EVAL_DIV: case JA_EVAL_RSQRT: case JA_EVAL_LIN3:
      case JA_EVAL_PROD: case JA_EVAL_POW: case JA_EVAL_DOT2: case JA_EVAL_DOT3: break;
      default: return false;
    }
    // large rounds gain nothing (and would put hundreds of blocks on the mailbox): pairs of the next round <= 2^15
    const size_t len_now = pending ? polys[0]->len / 2 : polys[0]->len;        // length this round evaluates
    return len_now >= 4 && len_now / 4 <= (size_t(1) << 15);
  }
  int32_t ahead_begin(ja_ctx* c) override {
    if (eq) {
      eq_save.current_index = (int)eq->current_index; eq_save.in_len = eq->in_len; eq_save.out_len = eq->out_len; eq_save.scalar = eq->current_scalar;
      const uint64_t zero[4] = {0, 0, 0, 0};
      int32_t st = ja_spliteq_bind(c, eq, zero);                      // advances the table level; the scalar is restored below
      if (st) return st;
    }
    pending = true;
    memset(pend_ch, 0, 32);                                          // the kernel takes the challenge from the mailbox
    slot_keep = slot;
    return JA_OK;
  }
  void ahead_end(ja_ctx*) override {
    slot_next = slot; slot = slot_keep;
    if (eq) { eq->current_index = eq_save.current_index; eq->in_len = eq_save.in_len; eq->out_len = eq_save.out_len; eq->current_scalar = eq_save.scalar; }
    ahead_consumed = true;
  }
  void ahead_commit() override { slot = slot_next; }

  // ---- round-resident kernel: the remaining rounds of this instance in ONE launch (persist_kernels.cuh) ------------------------------
  // len_now = length of the arrays the next round evaluates
  bool persist_ok(size_t len_now) const {
    if (!fusable || sharded || len_now < 2 || !c_rr_ok || !persist_allowed()) return false;
    switch (kind) {
      case JA_EVAL_ADD: case JA_EVAL_SUB: case JA_EVAL_MUL: case JA_EVAL_SQUARE: case JA_EVAL_IDENT:
      case JA_EVAL_IFF: case JA_EVAL_DIV: case JA_EVAL_RSQRT: case JA_EVAL_LIN3: return eq != nullptr && len_now <= kRrMaxLenS;
      case JA_EVAL_DOT2: case JA_EVAL_DOT3: return len_now <= kRrMaxLenDot;
      default: return false;
    }
  }
  void post_challenge(const uint64_t* ch, size_t rnd) override { if (pg && rnd >= pg_round0) pg->post(rnd - pg_round0, ch); }
  void abort_persist() override { pg.reset(); }
  void arm_pg_slot(ja_ctx* c, size_t rnd) {
    if (pg_round0 == ~size_t(0)) pg_round0 = rnd;
    char* h = reinterpret_cast<char*>(c->h_mapped) + (size_t)slot_id * kSlotBytes;
    slot = Slot();
    slot.host_vals = reinterpret_cast<const uint64_t*>(h);
    slot.pub.value = pg->tag0 + pg_tag_off + (unsigned int)(rnd - pg_round0);
    slot.armed = true;
    legacy = false; round_sharded = false;
  }
  // shared by the single-instance and the pair start: the other ping-pong buffer of every polynomial, common arguments
  int32_t persist_prepare(ja_ctx* c, RrCommon* cm, std::shared_ptr<PersistGroup> g, Fr* (*bufs)[2]) {
    const size_t len_in = polys[0]->len;
    const size_t len_now = pending ? len_in / 2 : len_in;
    const int nr = log2z(len_now);
    cm->mail_host = g->d_entries; cm->mail_relay = g->d_relay; cm->rounds = nr;
    cm->first_fused = pending ? 1 : 0; cm->r0 = to_challenge(pend_ch); cm->n_first = len_in;
    for (size_t q = 0; q < polys.size(); q++) {
      ja_poly* p = polys[q];
      const int nxt = 1 - p->cur;
      if (order == JA_LOW_TO_HIGH && p->cap[nxt] < len_in / 2) {
        dev_free(c, p->buf[nxt]);
        p->buf[nxt] = nullptr; p->cap[nxt] = 0;
        int32_t st = dev_alloc(c, (len_in / 2) * sizeof(Fr), (void**)&p->buf[nxt]);
        if (st) return st;
        p->cap[nxt] = len_in / 2;
      }
      bufs[q][0] = p->buf[p->cur]; bufs[q][1] = order == JA_LOW_TO_HIGH ? p->buf[nxt] : p->buf[p->cur];
    }
    // LowToHigh: the array moves to the other buffer with every fused round, and once more with the final bind
    const int fused_rounds = nr - (pending ? 0 : 1);
    pg_flip = order == JA_LOW_TO_HIGH && ((fused_rounds + 1) & 1);
    pg = g;
    pg_round0 = ~size_t(0);
    pending = false;
    return JA_OK;
  }
  int32_t start_persist(ja_ctx* c) {
    const size_t len_in = polys[0]->len;
    const size_t len_now = pending ? len_in / 2 : len_in;
    const int nr = log2z(len_now);
    auto g = std::make_shared<PersistGroup>();
    int32_t st = g->reserve(c, nr);
    if (st) return st;
    Fr* slot_vals = reinterpret_cast<Fr*>(reinterpret_cast<char*>(c->d_mapped) + (size_t)slot_id * kSlotBytes);
    const size_t G0 = len_now / 2;
    if (kind == JA_EVAL_DOT2 || kind == JA_EVAL_DOT3) {
      RrDotArgs a;
      memset(&a, 0, sizeof(a));
      Fr* bufs[3][2];
      if ((st = persist_prepare(c, &a.c, g, bufs))) return st;
      for (size_t q = 0; q < polys.size(); q++) a.buf[q] = bufs[q][0];
      a.partials = c->d_partials; a.counter = c->d_counter; a.slot_vals = slot_vals; a.tag0 = g->tag0;
      return rr_launch(c, kind == JA_EVAL_DOT2 ? (const void*)k_rr_dot<2> : (const void*)k_rr_dot<3>, 1, a);
    }
    JA_REQUIRE(eq && eq->order == order, "sumcheck: split-eq binding order does not match the round body");
    JA_REQUIRE((size_t(1) << ((eq->out_len - 1) + (eq->in_len - 1))) == G0, "sumcheck: split-eq tables do not cover len/2 (eq and polys out of lockstep)");
    RrSArgs a;
    memset(&a, 0, sizeof(a));
    Fr* bufs[kRrMaxSPolys][2];
    if ((st = persist_prepare(c, &a.c, g, bufs))) return st;
    a.np = (int)polys.size();
    for (int q = 0; q < kRrMaxSPolys; q++) { const int qq = q < a.np ? q : 0; a.buf[q][0] = bufs[qq][0]; a.buf[q][1] = bufs[qq][1]; }
    a.gammas = d_gammas;
    a.eq = rr_eq_state(eq);
    a.partials = c->d_partials; a.counter = c->d_counter; a.slot_vals = slot_vals; a.tag0 = g->tag0;
    a.split = RrSplit{(unsigned int)kBlock, 128u, 512u};   // one pair per thread and pass; 512 pairs or fewer on one block
    size_t w0 = G0 <= a.split.single_max ? 1 : std::min<size_t>(a.split.wmax, (G0 + a.split.ppp - 1) / a.split.ppp);
    const void* k = nullptr;
    switch (kind) {
      case JA_EVAL_ADD: k = (const void*)k_rr_s<0>; break;
      case JA_EVAL_SUB: k = (const void*)k_rr_s<1>; break;
      case JA_EVAL_MUL: k = (const void*)k_rr_s<2>; break;
      case JA_EVAL_SQUARE: k = (const void*)k_rr_s<3>; break;
      case JA_EVAL_IFF: k = (const void*)k_rr_s<8>; break;
      case JA_EVAL_DIV: k = (const void*)k_rr_s<9>; break;
      case JA_EVAL_RSQRT: k = (const void*)k_rr_s<10>; break;
      case JA_EVAL_LIN3: k = (const void*)k_rr_s<11>; break;
      default: k = (const void*)k_rr_s<6>; break;
    }
    return rr_launch(c, k, (unsigned int)w0, a);
  }

  // fused round kernel (bind pend_ch first when `pending`)
  int32_t launch_fused(ja_ctx* c) {
    Prep pr;
    int32_t pst = prepare(c, &pr);
    if (pst) return pst;
    const bool fz = pr.fz;
    const size_t G = pr.G;
    const FusedPolys& P = pr.P;
    const Challenge ch = pr.ch;
    const int bits_in = pr.bits_in;
    const Fr *e_out = pr.e_out, *e_in = pr.e_in;
    cudaStream_t s = c->stream;
    Fr* part = c->d_partials; unsigned int* ctr = c->d_counter;
    const Publish pub = slot.pub;
    if (kind == JA_EVAL_PROD || kind == JA_EVAL_POW) {
      const int d = kind == JA_EVAL_POW ? (int)pow_d : (int)polys.size();
      const bool same = kind == JA_EVAL_POW;
      int L = 2; while (L < d) L <<= 1;
      const bool big_wide = G >= kBigWideMinPairs && getenv("JA_BIGWIDE") != nullptr;
      if (!same && L == 16 && (G <= kWideMaxPairs || big_wide) && getenv("JA_NO_WIDE") == nullptr) {
        size_t wppb = 2;
        if (big_wide) { wppb = (G + (size_t)kSMs * 4 - 1) / ((size_t)kSMs * 4); wppb = (wppb + 1) & ~size_t(1); }
        const unsigned grid = (unsigned)((G + wppb - 1) / wppb);
        const MailRef mref = mail_ref(c);
        if (fz) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_prod16_wide<true><<<grid, kWideBlock, 0, s>>>(P, d, ch, e_out, e_in, bits_in, G, wppb, part, ctr, pub, pr.g_off, mref));
        else JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_prod16_wide<false><<<grid, kWideBlock, 0, s>>>(P, d, ch, e_out, e_in, bits_in, G, wppb, part, ctr, pub, pr.g_off, mref));
        prod_lanes = L;
        JA_CUDA(cudaGetLastError());
        commit_round(pr);
        return JA_OK;
      }
      const size_t gpb = (size_t)kBlock / L;
      size_t ppb = (G + (size_t)kSMs * 4 - 1) / ((size_t)kSMs * 4);
      ppb = (ppb + gpb - 1) / gpb * gpb;
      const unsigned grid = (unsigned)((G + ppb - 1) / ppb);
#define JA_PROD_F(LL, SM, FZ) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_prod<LL, SM, FZ><<<grid, kBlock, 0, s>>>(P, d, ch, e_out, e_in, bits_in, G, ppb, part, ctr, pub, pr.g_off, mail_ref(c)))
#define JA_PROD_L(LL) do { if (same) { if (fz) JA_PROD_F(LL, true, true); else JA_PROD_F(LL, true, false); } \
                           else { if (fz) JA_PROD_F(LL, false, true); else JA_PROD_F(LL, false, false); } } while (0)
      switch (L) { case 2: JA_PROD_L(2); break; case 4: JA_PROD_L(4); break; case 8: JA_PROD_L(8); break; default: JA_PROD_L(16); break; }
#undef JA_PROD_L
#undef JA_PROD_F
      prod_lanes = L;
    } else if (kind == JA_EVAL_OPEN) {
      unsigned grid = grid_for(G);
      if (grid > (unsigned)kSMs * 4) grid = kSMs * 4;
      const int bits_out = eq->out_len - 1;
      if (fz) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_open<true><<<dim3(grid, 1), kBlock, 0, s>>>(P, ch, e_out, e_in, bits_out, G, part, ctr, pub));
      else JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_open<false><<<dim3(grid, 1), kBlock, 0, s>>>(P, ch, e_out, e_in, bits_out, G, part, ctr, pub));
    } else if (kind == JA_EVAL_DOT2 || kind == JA_EVAL_DOT3) {
      unsigned grid = grid_for(G);
      if (grid > (unsigned)kSMs * 4) grid = kSMs * 4;
#define JA_DOT_F(NP, FZ) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_dot<NP, FZ><<<grid, kBlock, 0, s>>>(P, ch, G, part, ctr, pub, mail_ref(c)))
      if (kind == JA_EVAL_DOT2) { if (fz) JA_DOT_F(2, true); else JA_DOT_F(2, false); }
      else { if (fz) JA_DOT_F(3, true); else JA_DOT_F(3, false); }
#undef JA_DOT_F
    } else if (kind == 7 && polys.size() > 1 && polys.size() <= 16) {
      const int d = (int)polys.size();
      int L = 2; while (L < d) L <<= 1;
      const size_t gpb = (size_t)kBlock / L;
      size_t ppb = (G + (size_t)kSMs * 4 - 1) / ((size_t)kSMs * 4);
      ppb = (ppb + gpb - 1) / gpb * gpb;
      const unsigned grid = (unsigned)((G + ppb - 1) / ppb);
#define JA_BOOL_F(LL, FZ) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_bool<LL, FZ><<<grid, kBlock, 0, s>>>(P, d, ch, e_out, e_in, bits_in, G, ppb, d_gammas, part, ctr, pub, mail_ref(c)))
#define JA_BOOL_L(LL) do { if (fz) JA_BOOL_F(LL, true); else JA_BOOL_F(LL, false); } while (0)
      switch (L) { case 2: JA_BOOL_L(2); break; case 4: JA_BOOL_L(4); break; case 8: JA_BOOL_L(8); break; default: JA_BOOL_L(16); break; }
#undef JA_BOOL_L
#undef JA_BOOL_F
    } else if (!c->ahead_p && tma_eligible(kind, fz, G)) {
      int32_t tst = launch_round_s_tma(c, kind, P, ch, e_out, e_in, bits_in, G, part, ctr, pub, pr.g_off);
      if (tst) return tst;
    } else {
      size_t tiles = (G + kBlock - 1) / kBlock;
      size_t grid = tiles < (size_t)kSMs * 4 ? tiles : (size_t)kSMs * 4;
      const size_t tpb = (tiles + grid - 1) / grid;
      grid = (tiles + tpb - 1) / tpb;
      const int np = (int)polys.size();
#define JA_S_F(KID, FZ) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_s<KID, FZ><<<(unsigned)grid, kBlock, 0, s>>>(P, np, ch, e_out, e_in, bits_in, G, tpb, d_gammas, part, ctr, pub, pr.g_off, mail_ref(c)))
#define JA_S_K(KID) do { if (fz) JA_S_F(KID, true); else JA_S_F(KID, false); } while (0)
      switch (kind) {
        case JA_EVAL_ADD: JA_S_K(0); break;
        case JA_EVAL_SUB: JA_S_K(1); break;
        case JA_EVAL_MUL: JA_S_K(2); break;
        case JA_EVAL_SQUARE: JA_S_K(3); break;
        case JA_EVAL_IDENT: JA_S_K(6); break;
        case JA_EVAL_IFF: JA_S_K(8); break;
        case JA_EVAL_DIV: JA_S_K(9); break;
        case JA_EVAL_RSQRT: JA_S_K(10); break;
        case JA_EVAL_LIN3: JA_S_K(11); break;
        default: JA_S_K(7); break;
      }
#undef JA_S_K
#undef JA_S_F
    }
    JA_CUDA(cudaGetLastError());
    commit_round(pr);
    return JA_OK;
  }

  bool small_round = false, wide_round = false;   // prepare_pair chose the small-slab variants of the paired kernel
  bool paired_this_round = false;    // the driver already launched this round's kernel together with a partner
  DevInst* pair_candidate(size_t) override { return pairable() ? this : nullptr; }
  int32_t launch(ja_ctx* c, size_t rnd) override {
    int32_t st;
    if (paired_this_round) { paired_this_round = false; return JA_OK; }
    if (!pg && persist_ok(pending ? polys[0]->len / 2 : polys[0]->len) && (st = start_persist(c))) return st;
    if (pg) { arm_pg_slot(c, rnd); return JA_OK; }
    if (sharded && (pending ? polys[0]->len / 2 : polys[0]->len) < 2 && (st = unshard(c))) return st;
    legacy = !fusable;
    round_sharded = sharded;
    if (fusable) {
      slot = arm_slot(c, slot_id);
      if ((st = launch_fused(c))) return st;
    } else {
      const uint64_t* aux = gammas.empty() ? nullptr : reinterpret_cast<const uint64_t*>(gammas.data());
      if ((st = ja_round_eval_launch(c, kind, polys.data(), polys.size(), eq, aux, gammas.size(), pow_d, n_out, &pend))) return st;
    }
    return JA_OK;
  }
  // while the kernels run: the round's one field division (depends on the eq state only)
  int32_t prework(ja_ctx*, size_t) override {
    int32_t st;
    cs = host::FR_ONE; cw = host::FR_ZERO; div = host::FR_ZERO;
    if (eq) {
      uint64_t t[4];
      ja_spliteq_current_scalar(eq, t); cs = host::from_limbs(t);
      if ((st = ja_spliteq_current_w(eq, t))) return st;
      cw = host::from_limbs(t);
      if (w_inv.empty()) {
        const bool prod = kind == JA_EVAL_PROD || kind == JA_EVAL_POW;
        w_inv.resize(eq->w.size());
        for (size_t i = 0; i < w_inv.size(); i++) w_inv[i] = prod ? host::sub(host::FR_ONE, eq->w[i]) : eq->w[i];
        for (const FrH& x : w_inv)
          if (x.is_zero()) return fail(JA_ERR_INVALID, "sumcheck: eq point coordinate is 0 (split-eq bodies) or 1 (product bodies): the round polynomial division is undefined");
        host::batch_inv(w_inv.data(), w_inv.size());
      }
      div = w_inv[order == JA_LOW_TO_HIGH ? eq->current_index - 1 : eq->current_index];
    }
    return JA_OK;
  }
  // q(1) of the eq-free polynomial from the running normalised claim (split-eq bodies)
  FrH gruen_q1(const FrH& prev, const FrH& q0) {
    if (!nclaim_valid) { nclaim = cs == host::FR_ONE ? prev : host::mul(prev, host::inv(cs)); nclaim_valid = true; }
    return host::mul(host::sub(nclaim, host::mul(host::sub(host::FR_ONE, cw), q0)), div);
  }

  int32_t message(ja_ctx* c, size_t, const FrH& prev_in, Coeffs* uni) override {
    uint64_t ev[kMaxOut * 4];
    int32_t st;
    if (legacy) {
      // this path shares the pinned staging buffer: only one legacy instance may be in flight (the drivers collect
      // instances in launch order, and ja_round_eval_collect synchronises the stream)
      if ((st = ja_round_eval_collect(c, pend, ev))) return st;
    } else {
      const bool lanes = prod_lanes && (kind == JA_EVAL_PROD || kind == JA_EVAL_POW);
      const size_t n_raw = lanes ? (size_t)prod_lanes : n_out;
      uint64_t raw[kMaxOut * 4];
      if (kind == JA_EVAL_OPEN) {                                    // k_round_open: fence + flag protocol
        if ((st = wait_slot(c, slot))) return st;
        memcpy(raw, slot.host_vals, n_raw * 32);
      } else if ((st = wait_slot_tagged(c, slot, n_raw, raw))) {
        return st;
      }
      if (round_sharded) {
        // partial sums of this rank's slice: all-gather (<= 17 Fr per rank) and add as field elements
        std::vector<uint64_t> all(4 * n_raw * sc_world);
        if (c->sc_allgather(c->sc_user, raw, 32 * n_raw, all.data()) != 0) return fail(JA_ERR_INVALID, "sumcheck: allgather callback failed");
        for (size_t k = 0; k < n_raw; k++) {
          FrH acc = host::FR_ZERO;
          for (uint32_t r = 0; r < sc_world; r++) acc = host::add(acc, host::from_limbs(all.data() + 4 * (r * n_raw + k)));
          memcpy(raw + 4 * k, acc.l, 32);
        }
      }
      if (lanes) {
        const size_t d = n_out;
        memcpy(ev, raw, (d - 1) * 32);
        memcpy(ev + 4 * (d - 1), raw + 4 * (prod_lanes - 1), 32);
      } else {
        memcpy(ev, raw, n_out * 32);
      }
    }
    std::vector<FrH> e(n_out);
    for (size_t k = 0; k < n_out; k++) e[k] = host::from_limbs(ev + 4 * k);
    const FrH prev = has_scale ? host::mul(prev_in, scale_inv) : prev_in;
    switch (kind) {
      case JA_EVAL_ADD: case JA_EVAL_SUB: case JA_EVAL_IDENT: case JA_EVAL_OPEN: case JA_EVAL_LIN3: {
        const FrH q1 = gruen_q1(prev, e[0]);
        qc0 = e[0]; qc1 = host::sub(q1, e[0]); qc2 = host::FR_ZERO;
        *uni = host::gruen_poly_deg_2_q1(cs, cw, e[0], prev, q1); break;             // ops/add.rs:297-304, opening_reduction.rs:402
      }
      case JA_EVAL_MUL: case JA_EVAL_SQUARE: case 7: case JA_EVAL_IFF: case JA_EVAL_DIV: case JA_EVAL_RSQRT: {
        const FrH q1 = gruen_q1(prev, e[0]);
        qc0 = e[0]; qc1 = host::sub(host::sub(q1, e[0]), e[1]); qc2 = e[1];
        *uni = host::gruen_poly_deg_3_q1(cs, cw, e[0], e[1], prev, q1); break;       // ops/mul.rs:177, booleanity.rs:295-300
      }
      case JA_EVAL_PROD: case JA_EVAL_POW:
        for (auto& x : e) x = host::mul(x, cs);                                      // mles_product_sum.rs:120-128
        *uni = host::finish_mles_product_sum_from_evals(e, prev, cw, div); break;
      default:
        *uni = host::from_evals_and_hint(prev, e); break;                            // einsum/dot.rs:304,349; hamming_weight.rs:137
    }
    if (has_scale) *uni = scaled(*uni, scale);
    return JA_OK;
  }

  int32_t ingest(ja_ctx* c, const uint64_t ch[4], size_t) override {
    int32_t st;
    if (eq && (st = ja_spliteq_bind(c, eq, ch))) return st;
    if (nclaim_valid) {                                            // n' = q(r)
      const FrH r = host::from_limbs(ch);
      nclaim = host::add(qc0, host::mul(r, host::add(qc1, host::mul(r, qc2))));
    }
    if (pg) return JA_OK;                                           // the round-resident kernel binds it
    if (ahead_consumed) { ahead_consumed = false; return JA_OK; }   // the pre-launched kernel of the next round binds this challenge
    if (fusable) { memcpy(pend_ch, ch, 32); pending = true; return JA_OK; }
    return ja_bind_many(c, polys.data(), polys.size(), ch, order);
  }

  int32_t finalize(ja_ctx* c, uint64_t* staging, size_t* count) override {
    int32_t st;
    if (pg) {
      // the kernel's final bind left the claims in element 0 of every polynomial and published them into the instance's slot
      JA_REQUIRE(pg->posted == pg->rounds, "sumcheck: round-resident kernel did not receive every challenge");
      const char* h = reinterpret_cast<const char*>(c->h_mapped) + (size_t)slot_id * kSlotBytes;
      if ((st = wait_tagged(c, h, pg->tag0 + pg_tag_off + (unsigned int)pg->rounds, polys.size(), staging, "round-resident sumcheck kernel (final claims)"))) return st;
      for (ja_poly* p : polys) { if (pg_flip) p->cur = 1 - p->cur; p->len = 1; }
      pg.reset();
      *count = polys.size();
      return JA_OK;
    }
    if (pending) {
      if ((st = ja_bind_many(c, polys.data(), polys.size(), pend_ch, order))) return st;
      pending = false;
    }
    for (size_t i = 0; i < polys.size(); i++) {
      JA_REQUIRE(polys[i]->len == 1, "sumcheck: polynomial not fully bound at the end of the protocol");
      c->collect.push_back({polys[i]->data(), staging + 4 * i});
    }
    *count = polys.size();
    return JA_OK;
  }
  void release(ja_ctx* c) override {
    if (own_eq && eq) ja_spliteq_free(c, eq);
    eq = nullptr;
    if (d_gammas) { dev_free(c, d_gammas); d_gammas = nullptr; }
  }
};

// RaVirtual (product of d) + Booleanity phase 2 of one RA one-hot check: every remaining round of both in ONE launch (k_rr_pair)
bool persist_pair_ok(const DevInst* pa, const DevInst* pb) {
  if (!pa->c_rr_pair || !pb->c_rr_pair || !persist_allowed() || pa->pg || pb->pg || pa->sharded || pb->sharded || !pa->eq || !pb->eq) return false;
  const size_t len_now = pa->pending ? pa->polys[0]->len / 2 : pa->polys[0]->len;
  return len_now >= 2 && len_now <= kRrMaxLenPair && pa->polys.size() <= 16;
}
int32_t start_persist_pair(ja_ctx* c, DevInst* pa, DevInst* pb) {
  const size_t len_in = pa->polys[0]->len;
  const size_t len_now = pa->pending ? len_in / 2 : len_in;
  const int nr = log2z(len_now);
  const size_t G0 = len_now / 2;
  const int d = (int)pa->polys.size();
  int L = 2; while (L < d) L <<= 1;
  JA_REQUIRE((size_t(1) << ((pa->eq->out_len - 1) + (pa->eq->in_len - 1))) == G0 && (size_t(1) << ((pb->eq->out_len - 1) + (pb->eq->in_len - 1))) == G0,
             "sumcheck: split-eq tables do not cover len/2 (eq and polys out of lockstep)");
  auto g = std::make_shared<PersistGroup>();
  int32_t st = g->reserve(c, 2 * nr);                 // two tag ranges: [tag0, tag0 + nr) product, [tag0 + nr, tag0 + 2 nr) booleanity
  if (st) return st;
  g->rounds = nr;                                     // ... but one mailbox entry per round
  RrPairArgs a;
  memset(&a, 0, sizeof(a));
  RrCommon cmB;
  Fr* bufA[16][2];
  Fr* bufB[16][2];
  if ((st = pa->persist_prepare(c, &a.c, g, bufA))) return st;
  if ((st = pb->persist_prepare(c, &cmB, g, bufB))) return st;
  // sub-grids: the product of 16 costs ~3x the booleanity body per pair, below that they are about even
  const unsigned int ppp = (unsigned int)(kBlock / L);
  const unsigned int wmax_a = L == 16 ? 111u : 74u, wmax_b = (unsigned int)kSMs - wmax_a;
  a.split_a = RrSplit{ppp, wmax_a, ppp};
  a.split_wide = RrSplit{(unsigned int)(kBlock / 64), wmax_a, (unsigned int)(kBlock / 64)};
  a.split_b = RrSplit{ppp, wmax_b, ppp};
  a.wide_max_pairs = (L == 16 && getenv("JA_NO_WIDE") == nullptr) ? (unsigned int)kWideMaxPairs : 0u;
  a.d = d;
  for (int q = 0; q < d; q++) { a.bufA[q][0] = bufA[q][0]; a.bufA[q][1] = bufA[q][1]; a.bufB[q][0] = bufB[q][0]; a.bufB[q][1] = bufB[q][1]; }
  a.eqA = rr_eq_state(pa->eq); a.eqB = rr_eq_state(pb->eq);
  a.gammas = pb->d_gammas;
  a.partialsA = c->d_partials; a.counterA = c->d_counter;
  a.partialsB = c->d_partials + (size_t)kMaxGrid * kMaxOut / 2; a.counterB = c->d_counter + 1;
  a.slotA = reinterpret_cast<Fr*>(reinterpret_cast<char*>(c->d_mapped) + (size_t)pa->slot_id * kSlotBytes);
  a.slotB = reinterpret_cast<Fr*>(reinterpret_cast<char*>(c->d_mapped) + (size_t)pb->slot_id * kSlotBytes);
  a.tagA0 = g->tag0; a.tagB0 = g->tag0 + (unsigned int)nr + 1;
  pa->pg_tag_off = 0; pb->pg_tag_off = (unsigned int)nr + 1;
  pa->prod_lanes = L;
  const bool wide0 = a.wide_max_pairs && G0 <= a.wide_max_pairs;
  const RrSplit& sa = wide0 ? a.split_wide : a.split_a;
  const size_t wa0 = G0 <= sa.single_max ? 1 : std::min<size_t>(sa.wmax, (G0 + sa.ppp - 1) / sa.ppp);
  const size_t wb0 = G0 <= a.split_b.single_max ? 1 : std::min<size_t>(a.split_b.wmax, (G0 + a.split_b.ppp - 1) / a.split_b.ppp);
  a.off_b = (unsigned int)wa0;
  const size_t w0 = wa0 + wb0;
  const void* k = L == 2 ? (const void*)k_rr_pair<2> : L == 4 ? (const void*)k_rr_pair<4> : L == 8 ? (const void*)k_rr_pair<8> : (const void*)k_rr_pair<16>;
  return rr_launch(c, k, (unsigned int)w0, a);
}

// HammingWeightSumcheckProver over the K-entry G tables (hamming_weight.rs:60-160): log K rounds, degree 1, host only
struct HammingHostInst : Inst {
  std::vector<std::vector<FrH>> ra;
  std::vector<FrH> gammas;
  int32_t launch(ja_ctx*, size_t) override { return JA_OK; }
  bool can_ahead(size_t) override { return true; }                 // no kernels
  int32_t message(ja_ctx*, size_t, const FrH& prev, Coeffs* uni) override {
    FrH acc = host::FR_ZERO;
    for (size_t i = 0; i < ra.size(); i++) {
      FrH s = host::FR_ZERO;
      for (size_t j = 0; j < ra[i].size() / 2; j++) s = host::add(s, ra[i][2 * j]);
      acc = host::add(acc, host::mul(gammas[i], s));
    }
    *uni = host::from_evals_and_hint(prev, {acc});
    return JA_OK;
  }
  int32_t ingest(ja_ctx*, const uint64_t ch[4], size_t) override {
    const FrH r = host::from_limbs(ch);
    for (auto& p : ra) {
      const size_t n = p.size() / 2;
      for (size_t i = 0; i < n; i++) p[i] = host::add(p[2 * i], host::mul(r, host::sub(p[2 * i + 1], p[2 * i])));
      p.resize(n);
    }
    return JA_OK;
  }
  int32_t finalize(ja_ctx*, uint64_t* staging, size_t* count) override {
    for (size_t i = 0; i < ra.size(); i++) memcpy(staging + 4 * i, ra[i][0].l, 32);
    *count = ra.size();
    return JA_OK;
  }
};

// BooleanitySumcheckProver (booleanity.rs:153-372): log K address rounds on the host (G tables, expanding table F,
// split-eq B over r_address), then log T cycle rounds on the device over H_i[t] = F[k_i[t]] (body 7, split-eq D).
struct BooleanityInst : Inst {
  size_t d = 0, log_k = 0, log_t = 0;
  std::vector<FrH> gammas;
  std::vector<std::vector<FrH>> G;
  std::vector<FrH> F;
  HostEq B;
  const ja_addr* addr = nullptr;
  std::vector<uint64_t> r_cycle;
  std::unique_ptr<DevInst> p2;
  std::vector<ja_poly*> H;

  bool needs_slot() const override { return true; }
  // pre-launch only between two cycle rounds (p2 exists and has launched the round before `next`)
  bool can_ahead(size_t next) override { return next > log_k && p2 && p2->can_ahead(next - log_k); }
  int32_t ahead_begin(ja_ctx* c) override { return p2->ahead_begin(c); }
  void ahead_end(ja_ctx* c) override { p2->ahead_end(c); }
  void ahead_commit() override { p2->ahead_commit(); }
  DevInst* pair_candidate(size_t round) override { return round >= log_k && p2 && p2->pairable() ? p2.get() : nullptr; }
  void post_challenge(const uint64_t* ch, size_t round) override { if (round >= log_k && p2) p2->post_challenge(ch, round - log_k); }
  void abort_persist() override { if (p2) p2->abort_persist(); }
  int32_t launch(ja_ctx* c, size_t round) override { return round < log_k ? (int32_t)JA_OK : p2->launch(c, round - log_k); }
  int32_t prework(ja_ctx* c, size_t round) override { return round < log_k ? (int32_t)JA_OK : p2->prework(c, round - log_k); }
  int32_t message(ja_ctx* c, size_t round, const FrH& prev, Coeffs* uni) override {
    if (round >= log_k) return p2->message(c, round - log_k, prev, uni);
    const size_t m = round + 1;                                       // booleanity.rs:193-252
    const std::vector<FrH> E = B.table();
    FrH q0 = host::FR_ZERO, q1 = host::FR_ZERO;
    for (size_t kp = 0; kp < E.size(); kp++) {
      FrH c0 = host::FR_ZERO, c1 = host::FR_ZERO;
      for (size_t i = 0; i < d; i++) {
        FrH s0 = host::FR_ZERO, s1 = host::FR_ZERO;
        for (size_t k = 0; k < (size_t(1) << m); k++) {
          const FrH Gk = G[i][(kp << m) + k];
          const FrH Fk = F[k % (size_t(1) << (m - 1))];
          const FrH GF = host::mul(Gk, Fk);
          const FrH e_inf = host::mul(GF, Fk);
          if ((k >> (m - 1)) == 0) s0 = host::add(s0, host::sub(e_inf, GF));
          s1 = host::add(s1, e_inf);
        }
        c0 = host::add(c0, host::mul(gammas[i], s0));
        c1 = host::add(c1, host::mul(gammas[i], s1));
      }
      q0 = host::add(q0, host::mul(E[kp], c0));
      q1 = host::add(q1, host::mul(E[kp], c1));
    }
    const FrH cw = B.current_w();
    *uni = host::gruen_poly_deg_3(B.scalar, cw, q0, q1, prev, host::inv(host::gruen_eq1(B.scalar, cw)));
    return JA_OK;
  }
  int32_t ingest(ja_ctx* c, const uint64_t ch[4], size_t round) override {
    if (round >= log_k) return p2->ingest(c, ch, round - log_k);
    const FrH r = host::from_limbs(ch);
    B.bind(r);
    const size_t len = F.size();                                      // ExpandingTable::update, LowToHigh (expanding_table.rs:62-75)
    F.resize(2 * len);
    for (size_t i = 0; i < len; i++) { F[len + i] = host::mul(F[i], r); F[i] = host::sub(F[i], F[len + i]); }
    if (round + 1 < log_k) return JA_OK;
    // transition (booleanity.rs:330-347): eq_r_r = B.current_scalar; H_i = RaPolynomial(indices, F)
    std::vector<uint64_t> tabs(d * addr->K * 4);
    for (size_t i = 0; i < d; i++) memcpy(tabs.data() + i * addr->K * 4, F.data(), addr->K * 32);
    H.assign(d, nullptr);
    int32_t st = ja_addr_gather(c, addr, tabs.data(), H.data());
    if (st) return st;
    p2.reset(new DevInst());
    p2->kind = 7; p2->polys = H; p2->gammas = gammas;
    if ((st = ja_spliteq_new(c, r_cycle.data(), log_t, JA_LOW_TO_HIGH, nullptr, &p2->eq))) return st;
    p2->own_eq = true;
    p2->scale = B.scalar; p2->has_scale = true; p2->scale_inv = host::inv(B.scalar);
    p2->slot_id = slot_id;
    G.clear();
    return p2->setup(c);
  }
  int32_t finalize(ja_ctx* c, uint64_t* staging, size_t* count) override {
    JA_REQUIRE(p2 != nullptr, "sumcheck: booleanity never reached its cycle rounds");
    return p2->finalize(c, staging, count);
  }
  void release(ja_ctx* c) override {
    if (p2) p2->release(c);
    for (ja_poly* h : H) ja_poly_free(c, h);
    H.clear();
  }
};

// OneHotPolynomialProverOpening (opening_reduction.rs:503-723) for the d one-hot polynomials of one address batch that are
// opened at the same point (they share EqAddressState / EqCycleState in the reference, :207-258).  Every polynomial is
// its own sumcheck instance (own claim, own batching coefficient); the group does the shared work once per round:
// log K address rounds on the host (B = eq(r_address, .) bound HighToLow, expanding table F HighToLow, G_i), then log T
// cycle rounds on the device: ONE launch evaluates all d polynomials H_i[j] = F[k_i[j]] (k_round_open, d rows).
struct OpenGroup {
  size_t d = 0, log_k = 0, log_t = 0;
  const ja_addr* addr = nullptr;
  std::vector<FrH> B, F;
  std::vector<std::vector<FrH>> G;
  std::vector<ja_poly*> H;
  ja_spliteq* D = nullptr;
  std::vector<uint64_t> r_cycle;
  FrH ea = host::FR_ONE, ea_inv = host::FR_ONE;       // eq(r_address, r') and its inverse (:667-672)
  bool pending = false;
  uint64_t pend_ch[4] = {0, 0, 0, 0};
  size_t ingested_for = ~size_t(0);
  bool finalized = false;
  FrH cs, cw, div;
  std::vector<FrH> vals;

  struct OpenBatch* batch = nullptr;
  size_t row_base = 0;                                  // first row of this group in the batch's value array
  int32_t ingest(ja_ctx* c, const uint64_t ch[4], size_t round) {
    if (ingested_for == round) return JA_OK;
    ingested_for = round;
    int32_t st;
    if (round >= log_k) {
      if ((st = ja_spliteq_bind(c, D, ch))) return st;
      memcpy(pend_ch, ch, 32); pending = true;
      return JA_OK;
    }
    const FrH r = host::from_limbs(ch);
    const size_t half = B.size() / 2;                               // B.bind_parallel(r, HighToLow)
    for (size_t i = 0; i < half; i++) B[i] = host::add(B[i], host::mul(r, host::sub(B[i + half], B[i])));
    B.resize(half);
    std::vector<FrH> nf(F.size() * 2);                              // ExpandingTable::update, HighToLow (expanding_table.rs:76-86)
    for (size_t i = 0; i < F.size(); i++) { const FrH e1 = host::mul(r, F[i]); nf[2 * i] = host::sub(F[i], e1); nf[2 * i + 1] = e1; }
    F.swap(nf);
    if (round + 1 < log_k) return JA_OK;
    ea = B[0]; ea_inv = host::inv(ea);                              // B.final_claim()
    std::vector<uint64_t> tabs(d * addr->K * 4);
    for (size_t i = 0; i < d; i++) memcpy(tabs.data() + i * addr->K * 4, F.data(), addr->K * 32);
    H.assign(d, nullptr);
    if ((st = ja_addr_gather(c, addr, tabs.data(), H.data()))) return st;
    G.clear();
    return ja_spliteq_new(c, r_cycle.data(), log_t, JA_HIGH_TO_LOW, nullptr, &D);
  }
  int32_t finalize(ja_ctx* c) {
    if (finalized) return JA_OK;
    finalized = true;
    JA_REQUIRE(D != nullptr, "sumcheck: one-hot opening never reached its cycle rounds");
    if (pending) {
      int32_t st = ja_bind_many(c, H.data(), H.size(), pend_ch, JA_HIGH_TO_LOW);
      if (st) return st;
      pending = false;
    }
    return JA_OK;
  }
  void release(ja_ctx* c) {
    if (D) ja_spliteq_free(c, D);
    D = nullptr;
    for (ja_poly* h : H) ja_poly_free(c, h);
    H.clear();
  }
};

// All one-hot opening groups of a batched sumcheck: ONE launch per round for every group in its cycle phase
// (k_round_open_rows), ONE shared inversion per round (Montgomery's trick over the groups' eq(1) values).
struct OpenBatch {
  std::vector<OpenGroup*> groups;
  size_t total_rows = 0;
  OpenRow* d_rows = nullptr;
  Fr* d_partials = nullptr;
  unsigned int* d_counters = nullptr;
  Fr* d_vals = nullptr;                                 // row sums of the round on the device (shipped to the host in one burst)
  static constexpr unsigned long long kBigRowPairs = 1ull << 13;
  std::vector<OpenRow> rows;
  std::vector<OpenGroup*> active;
  size_t launched_round = ~size_t(0), pre_round = ~size_t(0), waited_round = ~size_t(0);
  unsigned int seq = 0;

  int32_t init(ja_ctx* c) {
    total_rows = 0;
    for (OpenGroup* g : groups) { g->row_base = total_rows; total_rows += g->d; }
    JA_REQUIRE(total_rows <= (size_t)kMaxRowVals, "sumcheck: too many one-hot opening instances in one batch");
    int32_t st;
    if ((st = dev_alloc(c, total_rows * sizeof(OpenRow), (void**)&d_rows))) return st;
    // partial sums of the two launches of a round (large rows, small rows): each has at most kSMs * 8 + rows entries
    if ((st = dev_alloc(c, 2 * ((size_t)kSMs * 8 + total_rows) * sizeof(Fr), (void**)&d_partials))) return st;
    if ((st = dev_alloc(c, (total_rows + 1) * sizeof(unsigned int), (void**)&d_counters))) return st;
    if ((st = dev_alloc(c, total_rows * sizeof(Fr), (void**)&d_vals))) return st;
    JA_CUDA(cudaMemsetAsync(d_vals, 0, total_rows * sizeof(Fr), c->stream));
    JA_CUDA(cudaMemsetAsync(d_counters, 0, (total_rows + 1) * sizeof(unsigned int), c->stream));
    return JA_OK;
  }
  // local round of group g at batch round `round`, or -1 when it has not started / is in its address phase
  static long cycle_round(const OpenGroup* g, size_t round, size_t max_rounds) {
    const size_t nr = g->log_k + g->log_t;
    if (max_rounds - round > nr) return -1;
    const size_t local = round - (max_rounds - nr);
    return local >= g->log_k ? (long)local : -1;
  }
  int32_t launch(ja_ctx* c, size_t round, size_t max_rounds) {
    if (launched_round == round) return JA_OK;
    launched_round = round;
    rows.clear(); active.clear();
    uint64_t ch_limbs[4] = {0, 0, 0, 0};
    for (OpenGroup* g : groups) {
      if (cycle_round(g, round, max_rounds) < 0) continue;
      active.push_back(g);
      const bool fz = g->pending;
      const size_t len_in = g->H[0]->len, len_eval = fz ? len_in / 2 : len_in, half = len_eval / 2;
      const size_t cover = size_t(1) << ((g->D->out_len - 1) + (g->D->in_len - 1));
      JA_REQUIRE(cover == half, "sumcheck: opening split-eq tables do not cover len/2");
      if (fz) memcpy(ch_limbs, g->pend_ch, 32);
      for (size_t i = 0; i < g->d; i++) {
        OpenRow r;
        r.z = g->H[i]->data(); r.e_out = g->D->e_out(); r.e_in = g->D->e_in();
        r.half = half; r.bits_out = (unsigned int)(g->D->out_len - 1); r.fused = fz ? 1u : 0u;
        r.out_index = (unsigned int)(g->row_base + i); r.pad = 0;
        rows.push_back(r);
      }
      if (fz) { for (ja_poly* p : g->H) p->len = len_eval; g->pending = false; }
    }
    if (rows.empty()) return JA_OK;
    // Rows of very different lengths share a batch (GPT-2: 2^12 ... 2^20 per polynomial): the long rows go out first with
    // many blocks each, the short ones in a second launch with few, so that neither idles the GPU nor floods it with
    // blocks that have nothing to do.
    std::stable_partition(rows.begin(), rows.end(), [](const OpenRow& r) { return r.half >= kBigRowPairs; });
    size_t n_big = 0;
    while (n_big < rows.size() && rows[n_big].half >= kBigRowPairs) n_big++;
    // the counters wrap back to zero by themselves (atomicInc), except the "rows done" word whose modulus changes with
    // the number of active rows: it sits right after the active rows' counters, freshly zeroed territory each time
    JA_REQUIRE(rows.size() * sizeof(OpenRow) <= kPinnedBytes, "sumcheck: opening row table too large for the staging buffer");
    memcpy(c->h_pinned, rows.data(), rows.size() * sizeof(OpenRow));
    JA_CUDA(cudaMemcpyAsync(d_rows, c->h_pinned, rows.size() * sizeof(OpenRow), cudaMemcpyHostToDevice, c->stream));
    seq = ++c->seq;
    Fr* hv = reinterpret_cast<Fr*>(c->d_rowvals);
    volatile unsigned int* hs = reinterpret_cast<volatile unsigned int*>(reinterpret_cast<char*>(c->d_rowvals) + kRowSeqOffset);
    const unsigned int n_all = (unsigned int)rows.size();
    auto blocks_per_row = [](size_t n_rows, size_t max_half) {
      size_t g = ((size_t)kSMs * 8 + n_rows - 1) / n_rows;
      const size_t need = (max_half + kBlock - 1) / kBlock;
      g = std::min(g, need);
      return (unsigned int)std::max<size_t>(1, std::min<size_t>(g, 65535));
    };
    if (n_big) {
      size_t max_half = 0;
      for (size_t i = 0; i < n_big; i++) max_half = std::max(max_half, (size_t)rows[i].half);
      const unsigned int gxa = blocks_per_row(n_big, max_half);
      JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_open_rows<<<dim3(gxa, (unsigned int)n_big), kBlock, 0, c->stream>>>(
                    d_rows, n_all, 0u, to_challenge(ch_limbs), d_partials, d_counters, d_vals, (unsigned int)total_rows, hv, hs, seq));
    }
    if (n_big < rows.size()) {
      const size_t n_small = rows.size() - n_big;
      const unsigned int gxb = blocks_per_row(n_small, (size_t)kBigRowPairs);
      JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_open_rows<<<dim3(gxb, (unsigned int)n_small), kBlock, 0, c->stream>>>(
                    d_rows, n_all, (unsigned int)n_big, to_challenge(ch_limbs), d_partials + ((size_t)kSMs * 8 + total_rows), d_counters, d_vals,
                    (unsigned int)total_rows, hv, hs, seq));
    }
    JA_CUDA(cudaGetLastError());
    return JA_OK;
  }
  int32_t prework(ja_ctx*, size_t round) {
    if (pre_round == round || active.empty()) return JA_OK;
    pre_round = round;
    // one inversion for all groups: inv(eq1_g) = inv(prod) * prod of the others
    std::vector<FrH> eq1(active.size()), pre(active.size());
    FrH run = host::FR_ONE;
    for (size_t k = 0; k < active.size(); k++) {
      OpenGroup* g = active[k];
      uint64_t t[4];
      ja_spliteq_current_scalar(g->D, t); g->cs = host::from_limbs(t);
      int32_t st = ja_spliteq_current_w(g->D, t);
      if (st) return st;
      g->cw = host::from_limbs(t);
      eq1[k] = host::gruen_eq1(g->cs, g->cw);
      // a zero divisor (a 0 coordinate of the opening point, or an eq scalar that hit 0) would zero the SHARED inverse and
      // silently corrupt every group of the batch; the reference divides per instance and fails loudly: so do we
      if (eq1[k].is_zero()) return fail(JA_ERR_INVALID, "sumcheck: opening reduction: eq(1) = current_scalar * w is zero (zero coordinate in an opening point)");
      pre[k] = run;
      run = host::mul(run, eq1[k]);
    }
    FrH inv = host::inv(run);
    for (size_t k = active.size(); k-- > 0;) {
      active[k]->div = host::mul(inv, pre[k]);
      inv = host::mul(inv, eq1[k]);
    }
    return JA_OK;
  }
  int32_t wait(ja_ctx* c, size_t round) {
    if (waited_round == round) return JA_OK;
    waited_round = round;
    if (rows.empty()) return JA_OK;
    Slot s;
    s.host_vals = reinterpret_cast<const uint64_t*>(c->h_rowvals);
    s.host_seq = reinterpret_cast<volatile unsigned int*>(reinterpret_cast<char*>(c->h_rowvals) + kRowSeqOffset);
    s.pub.value = seq;
    int32_t st = wait_slot(c, s);
    if (st) return st;
    for (OpenGroup* g : active) {
      g->vals.resize(g->d);
      for (size_t i = 0; i < g->d; i++) g->vals[i] = host::from_limbs(s.host_vals + 4 * (g->row_base + i));
    }
    return JA_OK;
  }
  void release(ja_ctx* c) {
    dev_free(c, d_rows); dev_free(c, d_partials); dev_free(c, d_counters); dev_free(c, d_vals);
    d_rows = nullptr; d_partials = nullptr; d_counters = nullptr; d_vals = nullptr;
  }
};

struct OpenMember : Inst {
  std::shared_ptr<OpenGroup> g;
  size_t i = 0;
  size_t batch_round = 0;                               // set by the driver before launch / message
  int32_t launch(ja_ctx*, size_t) override { return JA_OK; }     // the batch launches once for every group (prove_loop)
  int32_t message(ja_ctx* c, size_t round, const FrH& prev, Coeffs* uni) override {
    if (round < g->log_k) {                                         // opening_reduction.rs:579-629
      const size_t nu = g->log_k - round, half = g->B.size() / 2;
      const std::vector<FrH>& Gi = g->G[i];
      FrH e0 = host::FR_ZERO, e2 = host::FR_ZERO;
      for (size_t kp = 0; kp < half; kp++) {
        const FrH b0 = g->B[kp], b1 = g->B[kp + half];
        const FrH b2 = host::add(b1, host::sub(b1, b0));
        FrH i0 = host::FR_ZERO, i2 = host::FR_ZERO;
        for (size_t k = kp; k < Gi.size(); k += half) {
          const FrH GF = host::mul(Gi[k], g->F[k >> nu]);
          if (((k >> (nu - 1)) & 1) == 0) { i0 = host::add(i0, GF); i2 = host::sub(i2, GF); }
          else i2 = host::add(i2, host::add(GF, GF));
        }
        e0 = host::add(e0, host::mul(b0, i0));
        e2 = host::add(e2, host::mul(b2, i2));
      }
      *uni = host::from_evals_and_hint(prev, {e0, e2});
      return JA_OK;
    }
    int32_t st = g->batch->wait(c, batch_round);
    if (st) return st;
    *uni = scaled(host::gruen_poly_deg_2(g->cs, g->cw, g->vals[i], host::mul(prev, g->ea_inv), g->div), g->ea);   // :667-672
    return JA_OK;
  }
  int32_t ingest(ja_ctx* c, const uint64_t ch[4], size_t round) override { return g->ingest(c, ch, round); }
  int32_t finalize(ja_ctx* c, uint64_t* staging, size_t* count) override {
    int32_t st = g->finalize(c);
    if (st) return st;
    JA_REQUIRE(g->H[i]->len == 1, "sumcheck: one-hot opening not fully bound");
    c->collect.push_back({g->H[i]->data(), staging});
    *count = 1;
    return JA_OK;
  }
  void release(ja_ctx* c) override { if (i == 0) g->release(c); }
};

int32_t build_instance(ja_ctx* c, const ja_sc_instance& d, std::vector<std::unique_ptr<Inst>>* out, const uint64_t* pre_G = nullptr) {
  int32_t st;
  if (d.kind == JA_INST_BOOLEANITY) {
    JA_REQUIRE(d.addr && d.host_tables && d.eq_w && d.aux_fr, "sumcheck: booleanity needs addr, G tables, r_cycle and gammas|r_address");
    std::unique_ptr<BooleanityInst> b(new BooleanityInst());
    b->d = d.n_polys; b->log_k = d.aux_u32; b->log_t = d.eq_m; b->addr = d.addr;
    JA_REQUIRE(b->d == d.addr->d && d.table_len == d.addr->K && (size_t(1) << b->log_k) == d.addr->K && (size_t(1) << b->log_t) == d.addr->T,
               "sumcheck: booleanity shape mismatch (d, K = 2^log_k, T = 2^log_t)");
    JA_REQUIRE(d.n_aux == b->d + b->log_k && b->log_k >= 1 && b->log_t >= 1, "sumcheck: booleanity aux_fr = gammas (d) then r_address (log_k)");
    b->gammas.resize(b->d);
    for (size_t i = 0; i < b->d; i++) b->gammas[i] = host::from_limbs(d.aux_fr + 4 * i);
    b->B.init(d.aux_fr + 4 * b->d, b->log_k);
    b->G.resize(b->d);
    for (size_t i = 0; i < b->d; i++) {
      b->G[i].resize(d.table_len);
      memcpy(b->G[i].data(), d.host_tables + 4 * d.table_len * i, d.table_len * 32);
    }
    b->F = {host::FR_ONE};
    b->r_cycle.assign(d.eq_w, d.eq_w + 4 * d.eq_m);
    b->rounds = b->log_k + b->log_t;
    b->claim = host::FR_ZERO;                                        // booleanity.rs:70-72
    b->out_final = d.out_final_claims;
    out->emplace_back(b.release());
    return JA_OK;
  }
  if (d.kind == JA_INST_OPENING_ONEHOT) {
    JA_REQUIRE(d.addr && d.eq_w && d.aux_fr && d.host_tables, "sumcheck: one-hot opening needs addr, r_cycle, r_address and the d claims");
    std::shared_ptr<OpenGroup> g(new OpenGroup());
    g->d = d.addr->d; g->log_k = d.n_aux; g->log_t = d.eq_m; g->addr = d.addr;
    JA_REQUIRE(d.n_polys == g->d && (size_t(1) << g->log_k) == d.addr->K && (size_t(1) << g->log_t) == d.addr->T && g->log_k >= 1 && g->log_t >= 1,
               "sumcheck: one-hot opening shape mismatch (d, K = 2^|r_address|, T = 2^|r_cycle|)");
    g->r_cycle.assign(d.eq_w, d.eq_w + 4 * d.eq_m);
    // B = eq(r_address, .) (EqAddressState::new, :207-226), G_i over D.merge() = eq(r_cycle, .) (:541-566)
    g->B = {host::FR_ONE};
    for (size_t j = 0; j < g->log_k; j++) {
      const FrH wj = host::from_limbs(d.aux_fr + 4 * j);
      std::vector<FrH> nb(g->B.size() * 2);
      for (size_t i = 0; i < g->B.size(); i++) { nb[2 * i + 1] = host::mul(g->B[i], wj); nb[2 * i] = host::sub(g->B[i], nb[2 * i + 1]); }
      g->B.swap(nb);
    }
    g->F = {host::FR_ONE};
    std::vector<uint64_t> Gt_own;
    const uint64_t* Gt_data = pre_G;                                // run() computes the G tables of all groups in one batch
    if (!Gt_data) {
      Gt_own.resize(g->d * d.addr->K * 4);
      int32_t st2 = ja_addr_ra_evals(c, d.addr, d.eq_w, d.eq_m, Gt_own.data());
      if (st2) return st2;
      Gt_data = Gt_own.data();
    }
    struct { const uint64_t* p; const uint64_t* data() const { return p; } } Gt{Gt_data};
    g->G.resize(g->d);
    for (size_t i = 0; i < g->d; i++) { g->G[i].resize(d.addr->K); memcpy(g->G[i].data(), Gt.data() + i * d.addr->K * 4, d.addr->K * 32); }
    for (size_t i = 0; i < g->d; i++) {
      std::unique_ptr<OpenMember> m(new OpenMember());
      m->g = g; m->i = i;
      m->rounds = g->log_k + g->log_t;
      m->claim = host::from_limbs(d.host_tables + 4 * i);
      m->out_final = d.out_final_claims ? d.out_final_claims + 4 * i : nullptr;
      out->emplace_back(m.release());
    }
    return JA_OK;
  }
  if (d.kind == JA_INST_HAMMING_TABLES) {
    JA_REQUIRE(d.host_tables && is_pow2(d.table_len) && d.table_len >= 2, "sumcheck: hamming-weight instance needs power-of-two G tables");
    std::unique_ptr<HammingHostInst> h(new HammingHostInst());
    h->ra.resize(d.n_polys);
    for (size_t i = 0; i < d.n_polys; i++) {
      h->ra[i].resize(d.table_len);
      memcpy(h->ra[i].data(), d.host_tables + 4 * d.table_len * i, d.table_len * 32);
    }
    h->gammas.assign(d.n_polys, host::FR_ONE);
    if (d.aux_fr) {
      JA_REQUIRE(d.n_aux == d.n_polys, "sumcheck: hamming-weight instance takes one gamma per table");
      for (size_t i = 0; i < d.n_polys; i++) h->gammas[i] = host::from_limbs(d.aux_fr + 4 * i);
    }
    h->rounds = (size_t)log2z(d.table_len);
    h->claim = host::from_limbs(d.claim);
    h->out_final = d.out_final_claims;
    out->emplace_back(h.release());
    return JA_OK;
  }
  JA_REQUIRE(d.polys && d.n_polys, "sumcheck: instance without polynomials");
  std::unique_ptr<DevInst> v(new DevInst());
  v->kind = d.kind; v->pow_d = d.aux_u32;
  v->polys.assign(d.polys, d.polys + d.n_polys);
  if ((d.kind == JA_EVAL_RSQRT || d.kind == JA_EVAL_LIN3) && d.aux_fr) {
    v->gammas.resize(d.n_aux);
    for (size_t i = 0; i < d.n_aux; i++) v->gammas[i] = host::from_limbs(d.aux_fr + 4 * i);
  }
  if (d.kind == JA_EVAL_SUM1 && d.aux_fr) {
    JA_REQUIRE(d.n_aux == d.n_polys, "sumcheck: SUM1 takes one gamma per polynomial");
    v->gammas.resize(d.n_aux);
    for (size_t i = 0; i < d.n_aux; i++) v->gammas[i] = host::from_limbs(d.aux_fr + 4 * i);
  }
  if ((st = v->setup(c))) return st;
  if (v->uses_eq()) {
    JA_REQUIRE(d.eq_w && d.eq_m == v->rounds, "sumcheck: family S needs one eq point coordinate per round");
    if ((st = ja_spliteq_new(c, d.eq_w, d.eq_m, v->order, nullptr, &v->eq))) return st;
    v->own_eq = true;
  }
  v->claim = host::from_limbs(d.claim);
  v->out_final = d.out_final_claims;
  out->emplace_back(v.release());
  return JA_OK;
}

// the round loop shared by Sumcheck::prove (batched == false, one instance) and BatchedSumcheck::prove
int32_t prove_loop(ja_ctx* c, std::vector<std::unique_ptr<Inst>>& insts, bool batched, host::Blake2bTranscript& t,
                   size_t max_coeffs, uint64_t* out_coeffs, uint32_t* out_ncoeffs, uint64_t* out_challenges, OpenBatch* ob) {
  const size_t n = insts.size();
  JA_REQUIRE(n >= 1, "sumcheck: empty batch");
  size_t max_rounds = 0;
  for (auto& i : insts) max_rounds = std::max(max_rounds, i->rounds);
  std::vector<FrH> coeffs(n, host::FR_ONE), claims(n);
  for (auto& i : insts) t.append_scalar(i->claim);                                         // sumcheck.rs:43-46 / :574
  if (batched) for (size_t k = 0; k < n; k++) coeffs[k] = t.challenge_scalar();           // :48 challenge_vector
  for (size_t k = 0; k < n; k++) claims[k] = batched ? mul_pow_2(insts[k]->claim, max_rounds - insts[k]->rounds) : insts[k]->claim;
  int32_t st;
  // Host threads for the per-instance glue of LARGE batches (the opening reduction has one instance per committed
  // polynomial: 756 for nanoGPT, 2340 for GPT-2; ~20 field products per instance and round).  Only the one-hot opening
  // members run in parallel: their message() reads shared group state and nothing else.
  std::vector<char> is_open(n, 0);
  size_t n_open = 0;
  for (size_t k = 0; k < n; k++) if (dynamic_cast<OpenMember*>(insts[k].get())) { is_open[k] = 1; n_open++; }
  const int nth = n_open >= 128 ? host_threads() : 1;
  g_trace.start();
  // the launch phase of one round: every active instance enqueues its kernel (pairs / row batches share launches)
  auto launch_round = [&](size_t rnd) -> int32_t {
    const size_t rem = max_rounds - rnd;
    int32_t st = JA_OK;
      bool legacy_in_flight = false;
      if (ob) {
        for (size_t k = 0; k < n; k++) if (OpenMember* om = dynamic_cast<OpenMember*>(insts[k].get())) om->batch_round = rnd;
        if ((st = ob->launch(c, rnd, max_rounds))) return st;
      }
      // RA one-hot checks: RaVirtual (product of d) + Booleanity phase 2 of the same shape go out as ONE launch
      if (n <= 8) {
        DevInst *pa = nullptr, *pb = nullptr;
        for (size_t k = 0; k < n; k++) {
          if (rem > insts[k]->rounds) continue;
          DevInst* cand = insts[k]->pair_candidate(rnd - (max_rounds - insts[k]->rounds));
          if (!cand) continue;
          if (cand->kind == JA_EVAL_PROD && !pa) pa = cand;
          else if (cand->kind == 7 && !pb) pb = cand;
        }
        const bool pair_ok = pa && pb && pa->polys.size() == pb->polys.size() && pa->polys[0]->len == pb->polys[0]->len && pa->pending == pb->pending &&
                             (!pa->pending || memcmp(pa->pend_ch, pb->pend_ch, 32) == 0);
        if (pair_ok && persist_pair_ok(pa, pb)) {
          if ((st = start_persist_pair(c, pa, pb))) return st;          // the instances' launch() below only arm their slots
        } else if (pair_ok) {
          PairArgs A, B;
          DevInst::Prep ra, rb;
          if ((st = pa->prepare_pair(c, &A, &ra, 0))) return st;
          if ((st = pb->prepare_pair(c, &B, &rb, 1))) return st;
          const unsigned int gx = std::max(A.nb, B.nb);
          int L = 2; while (L < A.d) L <<= 1;
          const MailRef mref = mail_ref(c);
#define JA_PAIR_B(LL, FZ, BLK, WD) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_prod_bool<LL, FZ, BLK, WD><<<dim3(gx, 2), BLK, 0, c->stream>>>(A, B, ra.ch, mref))
#define JA_PAIR(LL) do { if (pa->small_round) { if (ra.fz) JA_PAIR_B(LL, true, kWideBlock, false); else JA_PAIR_B(LL, false, kWideBlock, false); } \
                           else { if (ra.fz) JA_PAIR_B(LL, true, kBlock, false); else JA_PAIR_B(LL, false, kBlock, false); } } while (0)
          if (pa->wide_round) { if (ra.fz) JA_PAIR_B(16, true, kWideBlock, true); else JA_PAIR_B(16, false, kWideBlock, true); }
          else switch (L) { case 2: JA_PAIR(2); break; case 4: JA_PAIR(4); break; case 8: JA_PAIR(8); break; default: JA_PAIR(16); break; }
#undef JA_PAIR
#undef JA_PAIR_B
          JA_CUDA(cudaGetLastError());
          pa->paired_this_round = true; pb->paired_this_round = true;
        }
      }
      for (size_t k = 0; k < n; k++) {
        if (rem > insts[k]->rounds) continue;
        // instances on the pinned-staging path cannot overlap each other: collect before launching the next one
        DevInst* dv = dynamic_cast<DevInst*>(insts[k].get());
        const bool is_legacy = dv && !dv->fusable;
        if (is_legacy && legacy_in_flight) return fail(JA_ERR_UNSUPPORTED, "sumcheck: at most one non-fused instance per batch");
        if ((st = insts[k]->launch(c, rnd - (max_rounds - insts[k]->rounds)))) return st;
        legacy_in_flight = legacy_in_flight || is_legacy;
      }
    return JA_OK;
  };
  // Pre-launch (fused_kernels.cuh: MailRef): right behind round j's kernels the engine enqueues round j+1's, which wait on
  // the device for r_j; launch overhead and latency leave the Fiat-Shamir critical path.  JA_NO_AHEAD=1 disables it.
  const bool ahead_on = !ob && c->h_mail && getenv("JA_NO_AHEAD") == nullptr && ahead_allowed();
  size_t prelaunched = ~size_t(0);
  struct MailGuard {                       // an enqueued kernel must never be left waiting: errors abort it
    volatile uint32_t* entry = nullptr; uint32_t tag = 0;
    void post(const uint64_t ch[4]) {      // words 0..2 first, the tagged word 3 last (one 16-byte read on the device sees a prefix of these stores)
      const Challenge cc = to_challenge(ch);
      entry[0] = cc.c[0]; entry[1] = cc.c[1]; entry[2] = cc.c[2];
      __atomic_thread_fence(__ATOMIC_RELEASE);
      entry[3] = (cc.c[3] & 0x1fffffffu) | (tag << 29);
      entry = nullptr;
    }
    ~MailGuard() { if (entry) { __atomic_thread_fence(__ATOMIC_RELEASE); entry[3] = (tag | 4u) << 29; } }
  } mail;
  for (size_t round = 0; round < max_rounds; round++) {
    const size_t remaining = max_rounds - round;
    std::vector<Coeffs> unis(n);
    if (prelaunched == round) {
      for (size_t k = 0; k < n; k++) if (remaining <= insts[k]->rounds) insts[k]->ahead_commit();
    } else if ((st = launch_round(round))) {
      return st;
    }
    if (ahead_on && round + 1 < max_rounds) {
      bool ok = true, any = false;
      for (size_t k = 0; k < n && ok; k++) {
        const bool now = remaining <= insts[k]->rounds, next = remaining - 1 <= insts[k]->rounds;
        if (next && !now) ok = false;                                // an instance starts next round: plain launch
        else if (next) { ok = insts[k]->can_ahead(round + 1 - (max_rounds - insts[k]->rounds)); any = true; }
      }
      if (ok && any) {
        const uint32_t idx = (++c->mail_seq) % kMailEntries;
        const uint32_t tag = c->mail_uses[idx] = (uint8_t)(c->mail_uses[idx] % 3 + 1);   // last tag of the entry -> 1, 2, 3, 1, ...: consecutive uses always differ
        c->ahead_tag = tag;
        c->ahead_p = reinterpret_cast<const char*>(c->d_mail) + 16 * idx;
        c->ahead_dev = reinterpret_cast<char*>(c->d_mail_dev) + 16 * idx;
        c->ahead_ticket = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(c->d_mail_dev) + 16 * kMailEntries) + idx;
        c->ahead_use = c->mail_seq ? c->mail_seq : (c->mail_seq = kMailEntries);      // never 0 (tickets start zeroed)
        for (size_t k = 0; k < n && !st; k++) if (remaining <= insts[k]->rounds) st = insts[k]->ahead_begin(c);
        if (!st) st = launch_round(round + 1);
        for (size_t k = 0; k < n; k++) if (remaining <= insts[k]->rounds) insts[k]->ahead_end(c);
        c->ahead_p = nullptr; c->ahead_dev = nullptr; c->ahead_tag = 0; c->ahead_ticket = nullptr; c->ahead_use = 0;
        mail.entry = reinterpret_cast<volatile uint32_t*>(reinterpret_cast<char*>(c->h_mail) + 16 * idx); mail.tag = tag;
        if (st) return st;
        prelaunched = round + 1;
      }
    }
    if (g_trace.on && getenv("JA_SC_TRACE")[0] == '2') {
      auto nw = std::chrono::steady_clock::now();
      fprintf(stderr, "  round %zu launch %.1f us\n", round, std::chrono::duration<double, std::micro>(nw - g_trace.last).count());
    }
    g_trace.lap(4);
    if (ob && (st = ob->prework(c, round))) return st;
    for (size_t k = 0; k < n; k++)
      if (remaining <= insts[k]->rounds && (st = insts[k]->prework(c, round - (max_rounds - insts[k]->rounds)))) return st;
    g_trace.lap(0);
    for (size_t k = 0; k < n; k++) {
      if (nth > 1 && is_open[k]) continue;
      const size_t nr = insts[k]->rounds;
      if (remaining > nr) unis[k] = host::trim({mul_pow_2(insts[k]->claim, remaining - nr - 1)});    // :96-106
      else if ((st = insts[k]->message(c, round - (max_rounds - nr), claims[k], &unis[k]))) return st;
    }
    Coeffs uni;
    if (nth > 1) {
      if (ob && (st = ob->wait(c, round))) return st;          // the members below only read the collected row sums
      int32_t first_err = JA_OK;
      uni = host::trim({});
#pragma omp parallel num_threads(nth)
      {
        Coeffs local = host::trim({});
#pragma omp for schedule(static)
        for (long k = 0; k < (long)n; k++) {
          if (is_open[k]) {
            const size_t nr = insts[k]->rounds;
            int32_t e = JA_OK;
            if (remaining > nr) unis[k] = host::trim({mul_pow_2(insts[k]->claim, remaining - nr - 1)});
            else e = insts[k]->message(c, round - (max_rounds - nr), claims[k], &unis[k]);
            if (e) {
#pragma omp critical(ja_sc_err)
              first_err = e;
              continue;
            }
          }
          if (batched) add_assign(local, scaled(unis[k], coeffs[k]));                      // :113-121
        }
#pragma omp critical(ja_sc_sum)
        add_assign(uni, local);
      }
      if (first_err) return fail(first_err, "sumcheck: a one-hot opening instance failed in a worker thread");
      g_trace.lap(1);
      if (!batched) uni = unis[0];
    } else {
      g_trace.lap(1);
      if (batched) {
        uni = host::trim({});
        for (size_t k = 0; k < n; k++) add_assign(uni, scaled(unis[k], coeffs[k]));        // :113-121
      } else {
        uni = unis[0];
      }
    }
    const Coeffs cp = host::compress(uni);
    if (cp.size() > max_coeffs) return fail(JA_ERR_INVALID, "sumcheck: max_coeffs too small");
    t.append_message("UniPoly_begin");                                                     // unipoly.rs:550-558
    for (auto& x : cp) t.append_scalar(x);
    t.append_message("UniPoly_end");
    uint64_t ch[4];
    t.challenge_scalar_optimized(ch);                                                      // :126 / :586
    if (mail.entry) mail.post(ch);                                                         // releases the pre-launched kernels of the next round
    for (size_t k = 0; k < n; k++)                                                         // ... and the round-resident ones
      if (remaining <= insts[k]->rounds) insts[k]->post_challenge(ch, round - (max_rounds - insts[k]->rounds));
    const FrH r = host::from_limbs(ch);
#pragma omp parallel for schedule(static) num_threads(nth) if (nth > 1)
    for (long k = 0; k < (long)n; k++) claims[k] = host::evaluate(unis[k], r);             // :130-134 / :589
    g_trace.lap(2);
    for (size_t k = 0; k < n; k++)
      if (remaining <= insts[k]->rounds && (st = insts[k]->ingest(c, ch, round - (max_rounds - insts[k]->rounds)))) return st;
    g_trace.lap(3);
    out_ncoeffs[round] = (uint32_t)cp.size();
    for (size_t k = 0; k < cp.size(); k++) memcpy(out_coeffs + 4 * (round * max_coeffs + k), cp[k].l, 32);
    memcpy(out_challenges + 4 * round, ch, 32);
  }
  // final claims: flush the deferred binds, ONE synchronisation for the whole batch
  uint64_t* staging = c->h_pinned;
  const uint64_t* fin_src = nullptr;
  std::vector<uint64_t> got;
  c->collect.clear();
  std::vector<std::pair<size_t, size_t>> span(n);
  size_t used = 0;
  for (size_t k = 0; k < n; k++) {
    size_t cnt = 0;
    JA_REQUIRE((used + kMaxProdPolys) * 32 <= kPinnedBytes, "sumcheck: too many final claims for the staging buffer");   // before finalize() writes
    if ((st = insts[k]->finalize(c, staging + 4 * used, &cnt))) return st;
    span[k] = {used, cnt};
    used += cnt;
  }
  if (used * 48 <= kRowSeqOffset && getenv("JA_NO_TAGGED_FINALS") == nullptr) {
    // tagged publication of the final claims: no stream synchronisation at the end of the call (everything the caller
    // does next with these buffers is stream-ordered behind the kernels of this call)
    const uint32_t tag = next_tag(c);
    got.assign(4 * (used ? used : 1), 0);
    memcpy(got.data(), staging, used * 32);                 // host-only instances wrote their claims into the staging area directly
    std::vector<unsigned int> published;
    published.reserve(c->collect.size());
    for (size_t base = 0; base < c->collect.size(); base += 32) {
      CollectIdxArgs a;
      const int m = (int)std::min<size_t>(32, c->collect.size() - base);
      for (int i = 0; i < 32; i++) {
        a.src[i] = i < m ? reinterpret_cast<const Fr*>(c->collect[base + i].first) : nullptr;
        a.idx[i] = i < m ? (unsigned int)((reinterpret_cast<const uint64_t*>(c->collect[base + i].second) - staging) / 4) : 0u;
        if (i < m) published.push_back(a.idx[i]);
      }
      JA_LAUNCH(c, KC_BIND, k_collect_finals_tagged<<<1, 32, 0, c->stream>>>(a, m, reinterpret_cast<Fr*>(c->d_rowvals), tag));
    }
    c->collect.clear();
    JA_CUDA(cudaGetLastError());
    for (unsigned int idx : published)
      if ((st = wait_tagged(c, reinterpret_cast<const char*>(c->h_rowvals) + (size_t)idx * 48, tag, 1, got.data() + 4 * (size_t)idx, "final-claim collection"))) return st;
    for (size_t k = 0; k < n; k++)
      if (insts[k]->out_final) memcpy(insts[k]->out_final, got.data() + 4 * span[k].first, span[k].second * 32);
    fin_src = got.data();
  } else {
    if ((st = flush_collect(c))) return st;
    JA_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t k = 0; k < n; k++)
      if (insts[k]->out_final) memcpy(insts[k]->out_final, staging + 4 * span[k].first, span[k].second * 32);
    fin_src = staging;
  }
  if (c->cache_openings) {
    // SumcheckInstanceProver::cache_openings at the end of Sumcheck::prove / BatchedSumcheck::prove (sumcheck.rs:167-176, :593-597):
    // every opening claim an instance hands to the accumulator is appended to the transcript (opening_proof.rs:281, :338, :398),
    // instance by instance, in the order of the instance's polynomials
    for (size_t k = 0; k < n; k++)
      for (size_t i = 0; i < span[k].second; i++) t.append_scalar(host::from_limbs(fin_src + 4 * (span[k].first + i)));
  }
  if (g_trace.on)
    fprintf(stderr, "[sc n=%zu rounds=%zu] cumulative us: launch=%.0f inv=%.0f wait+interp=%.0f transcript=%.0f ingest=%.0f\n", n, max_rounds,
            g_trace.t[4], g_trace.t[0], g_trace.t[1], g_trace.t[2], g_trace.t[3]);
  return JA_OK;
}

int32_t run(ja_ctx* c, const ja_sc_instance* descs, size_t n, bool batched, uint8_t transcript_state[32], uint32_t* n_rounds_io,
            size_t max_coeffs, uint64_t* out_coeffs, uint32_t* out_ncoeffs, uint64_t* out_challenges) {
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  std::vector<std::unique_ptr<Inst>> insts;
  int32_t st = JA_OK;
  const auto t_enter = std::chrono::steady_clock::now();
  // G tables (compute_ra_evals) of every one-hot opening group in ONE batch: their points are all known here
  std::vector<std::vector<uint64_t>> pre(n);
  {
    std::vector<const ja_addr*> addrs; std::vector<const uint64_t*> pts; std::vector<size_t> lts; std::vector<uint64_t*> outs;
    for (size_t k = 0; k < n; k++) {
      const ja_sc_instance& d = descs[k];
      if (d.kind != JA_INST_OPENING_ONEHOT || !d.addr || !d.eq_w || (size_t(1) << d.eq_m) != d.addr->T) continue;
      pre[k].resize(d.addr->d * d.addr->K * 4);
      addrs.push_back(d.addr); pts.push_back(d.eq_w); lts.push_back(d.eq_m); outs.push_back(pre[k].data());
    }
    if (addrs.size() >= 2) st = ja_addr_ra_evals_many(c, addrs.data(), pts.data(), lts.data(), addrs.size(), outs.data());
    else for (auto& v : pre) v.clear();
  }
  {
    // shape of the call: one device-backed instance, or [PROD, (host-only ...), BOOLEANITY] (the RA one-hot checks)
    size_t n_dev = 0, n_prod = 0, n_bool = 0, n_open = 0;
    for (size_t k = 0; k < n; k++) {
      const int kd = descs[k].kind;
      if (kd == JA_INST_BOOLEANITY) n_bool++;
      else if (kd == JA_INST_OPENING_ONEHOT) n_open++;
      else if (kd != JA_INST_HAMMING_TABLES) { n_dev++; if (kd == JA_EVAL_PROD) n_prod++; }
    }
    c->rr_call_ok = n_open == 0 && n_dev == 1 && n_bool == 0;
    c->rr_call_pair = n_open == 0 && n_dev == 1 && n_prod == 1 && n_bool == 1;
  }
  for (size_t k = 0; k < n && !st; k++) st = build_instance(c, descs[k], &insts, pre[k].empty() ? nullptr : pre[k].data());
  int next_slot = 0;
  for (auto& i : insts) if (i->needs_slot()) i->slot_id = next_slot++;
  if (!st && next_slot > kSlots) st = fail(JA_ERR_UNSUPPORTED, "sumcheck: more than " + std::to_string(kSlots) + " kernel-backed instances / groups in one batch");
  OpenBatch ob;
  for (auto& i : insts)
    if (OpenMember* om = dynamic_cast<OpenMember*>(i.get()))
      if (om->i == 0) { ob.groups.push_back(om->g.get()); om->g->batch = &ob; }
  if (!st && !ob.groups.empty()) st = ob.init(c);
  host::Blake2bTranscript t(transcript_state, *n_rounds_io);
  const auto t_built = std::chrono::steady_clock::now();
  if (!st) st = prove_loop(c, insts, batched, t, max_coeffs, out_coeffs, out_ncoeffs, out_challenges, ob.groups.empty() ? nullptr : &ob);
  const auto t_proved = std::chrono::steady_clock::now();
  if (st) {
    for (auto& i : insts) if (i) i->abort_persist();
    cudaStreamSynchronize(c->stream);      // nothing of this call may still be in flight when the handles are released
  }
  for (auto& i : insts) if (i) i->release(c);
  ob.release(c);
  if (g_trace.on) {
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    fprintf(stderr, "[sc-call n=%zu] build=%.0f us loop=%.0f us release=%.0f us\n", n, us(t_enter, t_built), us(t_built, t_proved),
            us(t_proved, std::chrono::steady_clock::now()));
  }
  if (st) return st;
  memcpy(transcript_state, t.state, 32);
  *n_rounds_io = t.n_rounds;
  return JA_OK;
}

}  // namespace

extern "C" {

// bench hook: ONE fused round kernel (bind the previous challenge + evaluate) re-run `iters` times on resident synthetic
// operands of 2^log_n Fr per polynomial.  which: 0 ADD (2 polys), 1 MUL (2), 2 IDENT (1), 3 product of 4, 4 product of 16,
// 5 booleanity over 16, 6 opening reduction HighToLow (1 poly, in place), 7 / 8 = ADD / IDENT through the TMA-staged kernel,
// 9 product of 16 on 64 threads per pair (small-slab variant), 10 / 11 = the paired RA-check launch (product of 16 +
// booleanity over 16) in its 256-thread / 64-threads-per-pair form, 12 = the same in 128-thread blocks.  Algorithmic bytes per launch: 48 * 2^log_n per polynomial
// (32 n read + 16 n written).
int32_t ja_bench_fused(ja_ctx* c, int32_t which, int32_t log_n, int32_t iters, float* out_ms) {
  JA_REQUIRE(c && out_ms && iters > 0 && log_n >= 3 && log_n <= 28 && which >= 0 && which <= 12, "ja_bench_fused: bad argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  static const int kNp[13] = {2, 2, 1, 4, 16, 16, 1, 2, 1, 16, 32, 32, 32};
  const int np = kNp[which];
  const size_t n = size_t(1) << log_n, G = n / 4;
  std::vector<ja_poly*> src(np, nullptr);
  std::vector<Fr*> dst(np, nullptr);
  int32_t st;
  for (int i = 0; i < np; i++) {
    if ((st = ja_poly_random(c, n, 31u + i, &src[i]))) return st;
    if ((st = dev_alloc(c, (n / 2) * sizeof(Fr), (void**)&dst[i]))) return st;
  }
  std::vector<uint64_t> w((size_t)log_n * 4, 0);
  for (int i = 0; i < log_n; i++) { w[4 * i + 2] = 0x9e3779b97f4a7c15ull * (i + 1); w[4 * i + 3] = 0x0123456789abcdefull + i; }
  ja_spliteq* eq = nullptr;
  if ((st = ja_spliteq_new(c, w.data(), (size_t)log_n - 1, which == 6 ? JA_HIGH_TO_LOW : JA_LOW_TO_HIGH, nullptr, &eq))) return st;
  Fr* d_gam = nullptr;
  if ((st = dev_alloc(c, 16 * sizeof(Fr), (void**)&d_gam))) return st;
  JA_CUDA(cudaMemcpyAsync(d_gam, src[0]->data(), 16 * sizeof(Fr), cudaMemcpyDeviceToDevice, c->stream));
  const uint64_t rr[4] = {0, 0, 0x0123456789abcdefull, 0x0fedcba987654321ull};
  const Challenge ch = to_challenge(rr);
  FusedPolys P;
  for (int i = 0; i < np; i++) { P.in[i] = src[i]->data(); P.out[i] = which == 6 ? src[i]->data() : dst[i]; }
  const int bits_in = eq->in_len - 1, bits_out = eq->out_len - 1;
  Slot slot = arm_slot(c, 0), slot2 = arm_slot(c, 1);
  int32_t lst = JA_OK;
  auto launch = [&]() {
    cudaStream_t s = c->stream;
    if (which == 9) {
      size_t wppb = 2;
      if (G > kWideMaxPairs) { wppb = (G + (size_t)kSMs * 4 - 1) / ((size_t)kSMs * 4); wppb = (wppb + 1) & ~size_t(1); }
      JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_prod16_wide<true><<<(unsigned)((G + wppb - 1) / wppb), kWideBlock, 0, s>>>(P, np, ch, eq->e_out(), eq->e_in(), bits_in, G, wppb, c->d_partials, c->d_counter, slot.pub));
    } else if (which >= 10) {
      const bool wide = which == 11;
      PairArgs A, B;
      for (int q = 0; q < 16; q++) { A.P.in[q] = P.in[q]; A.P.out[q] = P.out[q]; B.P.in[q] = P.in[16 + q]; B.P.out[q] = P.out[16 + q]; }
      A.d = B.d = 16; A.e_out = B.e_out = eq->e_out(); A.e_in = B.e_in = eq->e_in(); A.bits_in = B.bits_in = bits_in; A.G = B.G = G;
      if (wide && G > kWideMaxPairs) { A.ppb = big_wide_ppb_prod(G); B.ppb = big_wide_ppb_bool(G); }
      else if (wide) { A.ppb = 2; B.ppb = kWideBlock / 16; }
      else if (which == 12) { A.ppb = B.ppb = kWideBlock / 16; }
      else { size_t ppb = (G + (size_t)kSMs * 2 - 1) / ((size_t)kSMs * 2); ppb = (ppb + 15) / 16 * 16; A.ppb = B.ppb = ppb; }
      A.nb = (unsigned)((G + A.ppb - 1) / A.ppb); B.nb = (unsigned)((G + B.ppb - 1) / B.ppb);
      A.gammas = B.gammas = d_gam;
      A.partials = c->d_partials; B.partials = c->d_partials + kMaxGrid * kMaxOut / 2;
      A.counter = c->d_counter; B.counter = c->d_counter + 1;
      A.pub = slot.pub; B.pub = slot2.pub;
      const unsigned gx = std::max(A.nb, B.nb);
      if (wide) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_prod_bool<16, true, kWideBlock, true><<<dim3(gx, 2), kWideBlock, 0, s>>>(A, B, ch));
      else if (which == 12) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_prod_bool<16, true, kWideBlock, false><<<dim3(gx, 2), kWideBlock, 0, s>>>(A, B, ch));
      else JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_prod_bool<16, true, kBlock><<<dim3(gx, 2), kBlock, 0, s>>>(A, B, ch));
    } else if (which >= 7) {
      const int32_t e = launch_round_s_tma(c, which == 7 ? JA_EVAL_ADD : JA_EVAL_IDENT, P, ch, eq->e_out(), eq->e_in(), bits_in, G,
                                           c->d_partials, c->d_counter, slot.pub, 0);
      if (e) lst = e;
    } else if (which <= 2) {
      size_t tiles = (G + kBlock - 1) / kBlock;
      size_t grid = tiles < (size_t)kSMs * 4 ? tiles : (size_t)kSMs * 4;
      const size_t tpb = (tiles + grid - 1) / grid;
      grid = (tiles + tpb - 1) / tpb;
      if (which == 0) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_s<0, true><<<(unsigned)grid, kBlock, 0, s>>>(P, np, ch, eq->e_out(), eq->e_in(), bits_in, G, tpb, d_gam, c->d_partials, c->d_counter, slot.pub));
      else if (which == 1) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_s<2, true><<<(unsigned)grid, kBlock, 0, s>>>(P, np, ch, eq->e_out(), eq->e_in(), bits_in, G, tpb, d_gam, c->d_partials, c->d_counter, slot.pub));
      else JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_s<6, true><<<(unsigned)grid, kBlock, 0, s>>>(P, np, ch, eq->e_out(), eq->e_in(), bits_in, G, tpb, d_gam, c->d_partials, c->d_counter, slot.pub));
    } else if (which <= 5) {
      const int L = np;
      const size_t gpb = (size_t)kBlock / L;
      size_t ppb = (G + (size_t)kSMs * 4 - 1) / ((size_t)kSMs * 4);
      ppb = (ppb + gpb - 1) / gpb * gpb;
      const unsigned grid = (unsigned)((G + ppb - 1) / ppb);
      if (which == 3) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_prod<4, false, true><<<grid, kBlock, 0, s>>>(P, np, ch, eq->e_out(), eq->e_in(), bits_in, G, ppb, c->d_partials, c->d_counter, slot.pub));
      else if (which == 4) JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_prod<16, false, true><<<grid, kBlock, 0, s>>>(P, np, ch, eq->e_out(), eq->e_in(), bits_in, G, ppb, c->d_partials, c->d_counter, slot.pub));
      else JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_bool<16, true><<<grid, kBlock, 0, s>>>(P, np, ch, eq->e_out(), eq->e_in(), bits_in, G, ppb, d_gam, c->d_partials, c->d_counter, slot.pub));
    } else {
      unsigned grid = grid_for(G);
      if (grid > (unsigned)kSMs * 4) grid = kSMs * 4;
      JA_LAUNCH(c, KC_SUMCHECK_FUSED, k_round_open<true><<<dim3(grid, 1), kBlock, 0, s>>>(P, ch, eq->e_out(), eq->e_in(), bits_out, G, c->d_partials, c->d_counter, slot.pub));
    }
  };
  for (int i = 0; i < 3; i++) launch();
  if (lst) return lst;
  JA_CUDA(cudaEventRecord(c->ev0, c->stream));
  for (int i = 0; i < iters; i++) launch();
  JA_CUDA(cudaEventRecord(c->ev1, c->stream));
  JA_CUDA(cudaEventSynchronize(c->ev1));
  JA_CUDA(cudaGetLastError());
  float ms = 0;
  JA_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  *out_ms = ms / iters;
  for (int i = 0; i < np; i++) { ja_poly_free(c, src[i]); dev_free(c, dst[i]); }
  dev_free(c, d_gam);
  ja_spliteq_free(c, eq);
  return JA_OK;
}

static int32_t lib_allgather(void* user, const void* send, size_t bytes, void* recv) {
  return comm_allgather(static_cast<ja_ctx*>(user), send, bytes, recv) == JA_OK ? 0 : -1;
}
int32_t ja_set_sumcheck_shard(ja_ctx* c, uint32_t rank, uint32_t world, ja_allgather_fn allgather, void* user) {
  JA_REQUIRE(c && world >= 1 && rank < world && is_pow2(world), "ja_set_sumcheck_shard: world must be a power of two and rank < world");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  if (!allgather && world > 1 && c->comm && c->comm_world == world && c->comm_rank == rank) { allgather = lib_allgather; user = c; }   // the context's own communicator
  c->sc_rank = rank; c->sc_world = allgather ? world : 1; c->sc_allgather = allgather; c->sc_user = user;
  return JA_OK;
}

int32_t ja_batched_sumcheck_prove(ja_ctx* c, const ja_sc_instance* instances, size_t n_instances, uint8_t transcript_state[32],
                                  uint32_t* n_rounds_io, size_t max_coeffs, uint64_t* out_coeffs, uint32_t* out_ncoeffs,
                                  uint64_t* out_challenges) {
  JA_REQUIRE(c && instances && n_instances && transcript_state && n_rounds_io && out_coeffs && out_ncoeffs && out_challenges,
             "ja_batched_sumcheck_prove: null argument");
  return run(c, instances, n_instances, true, transcript_state, n_rounds_io, max_coeffs, out_coeffs, out_ncoeffs, out_challenges);
}

int32_t ja_sumcheck_prove(ja_ctx* c, int32_t kind, ja_poly* const* polys, size_t n_polys, const uint64_t* eq_w, size_t eq_m,
                          const uint64_t* aux_fr, size_t n_aux, uint32_t aux_u32, const uint64_t claim[4],
                          uint8_t transcript_state[32], uint32_t* n_rounds_io, size_t max_coeffs, uint64_t* out_coeffs,
                          uint32_t* out_ncoeffs, uint64_t* out_challenges, uint64_t* out_final_claims) {
  JA_REQUIRE(c && polys && n_polys && claim && transcript_state && n_rounds_io && out_coeffs && out_ncoeffs && out_challenges,
             "ja_sumcheck_prove: null argument");
  ja_sc_instance d;
  memset(&d, 0, sizeof(d));
  d.kind = kind; d.aux_u32 = aux_u32; d.n_polys = n_polys; d.polys = polys;
  d.eq_w = eq_w; d.eq_m = eq_m; d.aux_fr = aux_fr; d.n_aux = n_aux;
  memcpy(d.claim, claim, 32);
  d.out_final_claims = out_final_claims;
  return run(c, &d, 1, false, transcript_state, n_rounds_io, max_coeffs, out_coeffs, out_ncoeffs, out_challenges);
}

}  // extern "C"
