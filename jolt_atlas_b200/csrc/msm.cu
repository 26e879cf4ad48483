// C ABI for the commitment half of the hot path: SRS residency, batched Pippenger MSM, one-hot point sums.
// Kernels: msm_kernels.cuh.  No CPU fallback.
#include "common.hpp"
#include "fq_host.hpp"
#include "msm_kernels.cuh"

#include <algorithm>
#include <cstdlib>

static inline uint32_t ceil_div_u32(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// window width for n pairs of nbits-bit scalars: minimise nwin * (n + ~3 * 2^(c-1)) (madds + bucket reduction)
static uint32_t pick_window(size_t n, uint32_t nbits) {
  if (const char* e = getenv("JA_MSM_C")) { int v = atoi(e); if (v >= 2 && v <= 22) return (uint32_t)v; }
  uint32_t best_c = 2; double best = 1e300;
  for (uint32_t c = 2; c <= 16; c++) {   // c > 16: scatter + bucket reduction outgrow the saved additions (measured)
    const uint32_t nwin = (nbits + 1 + c - 1) / c;
    const uint32_t top_bits = nbits + 1 - (nwin - 1) * c;                 // content bits of the top window (incl. carry)
    const double top_frac = 1.0 - 1.0 / (double)(1ull << (top_bits < 30 ? top_bits : 30));
    // mixed additions (one per non-zero digit) + bucket reduction (2 full additions per bucket) + fixed per-window work
    const double cost = ((double)(nwin - 1) + top_frac) * (double)n + (double)nwin * (2.8 * (double)(1u << (c - 1)) + 2000.0);
    if (cost < best) { best = cost; best_c = c; }
  }
  return best_c;
}

// Jobs shorter than the table's bucket count keep per-window buckets: their buckets would hold < ~1 entry each and the
// accumulate pass degenerates into bucket flushes (measured: HyperKZG phase-1 batch 1.3 -> 5 ms when every folded
// polynomial went through the table).
static size_t table_min_n(const ja_srs* srs) {
  if (const char* e = getenv("JA_MSM_TABLE_MIN_LOG")) { int v = atoi(e); if (v >= 0 && v <= 40) return size_t(1) << v; }
  return size_t(1) << (srs->table_c - 1);
}
// the wider second table: window bits (JA_MSM_TABLE2_C overrides, 0 disables) and the shortest job that uses it
static uint32_t table2_window_bits(size_t n) {
  if (const char* e = getenv("JA_MSM_TABLE2_C")) { int v = atoi(e); if (v == 0 || (v >= 12 && v <= 22)) return (uint32_t)v; }
  return n >= (size_t(1) << 21) ? 20u : 0u;
}
static size_t table2_min_n(const ja_srs* srs) {
  if (const char* e = getenv("JA_MSM_TABLE2_MIN_LOG")) { int v = atoi(e); if (v >= 0 && v <= 40) return size_t(1) << v; }
  return size_t(1) << (srs->table2_c + 1);
}

static uint32_t run_length() {
  if (const char* e = getenv("JA_MSM_T")) { int v = atoi(e); if (v >= 4 && v <= 4096) return (uint32_t)v; }
  return 64;
}

// Batch of indexed point sums (every job MSM_INDEXED): two launches + one host-side batch normalisation.
static int32_t indexed_engine(ja_ctx* c, const ja_srs* srs, const std::vector<MsmJob>& jobs, MsmResult* out) {
  const uint32_t count = (uint32_t)jobs.size();
  std::vector<IdxJob> ij(count);
  uint64_t blocks = 0;
  for (uint32_t m = 0; m < count; m++) {
    JA_REQUIRE(jobs[m].n < (1ull << 32), "indexed sum: list too long");
    ij[m] = IdxJob{(const unsigned long long*)jobs[m].d_scalars, (uint32_t)jobs[m].n, (uint32_t)blocks};
    blocks += (jobs[m].n + kIdxBlock * kIdxRun - 1) / (kIdxBlock * kIdxRun);
    JA_REQUIRE(blocks < (1ull << 31), "indexed sum: batch too large");
  }
  auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t o_jobs = 0, o_part = align(sizeof(IdxJob) * count), o_out = align(o_part + sizeof(G1X) * (blocks ? blocks : 1));
  const size_t total = align(o_out + sizeof(G1X) * count);
  char* ws = nullptr;
  int32_t st = dev_alloc(c, total, (void**)&ws);
  if (st) return st;
  IdxJob* d_jobs = (IdxJob*)(ws + o_jobs);
  G1X* d_part = (G1X*)(ws + o_part);
  G1X* d_out = (G1X*)(ws + o_out);
  cudaStream_t s = c->stream;
  JA_CUDA(cudaMemcpyAsync(d_jobs, ij.data(), sizeof(IdxJob) * count, cudaMemcpyHostToDevice, s));
  if (blocks) JA_LAUNCH(c, KC_ONEHOT_SUM, k_indexed_partial<<<(unsigned)blocks, kIdxBlock, 0, s>>>(d_jobs, count, srs->points, d_part));
  JA_LAUNCH(c, KC_ONEHOT_SUM, k_indexed_final<<<count, kIdxBlock, 0, s>>>(d_jobs, count, d_part, (uint32_t)blocks, d_out));
  JA_CUDA(cudaGetLastError());
  std::vector<host::G1XH> sums(count);
  JA_CUDA(cudaMemcpyAsync(sums.data(), d_out, sizeof(G1X) * count, cudaMemcpyDeviceToHost, s));
  JA_CUDA(cudaStreamSynchronize(s));
  dev_free(c, ws);
  std::vector<uint64_t> xy((size_t)count * 8);
  std::vector<int32_t> inf(count);
  host::xyzz_batch_to_affine(sums.data(), count, xy.data(), inf.data());
  for (uint32_t m = 0; m < count; m++) {
    memcpy(out[m].x.l, xy.data() + 8 * m, 32);
    memcpy(out[m].y.l, xy.data() + 8 * m + 4, 32);
    out[m].inf = (uint32_t)inf[m];
  }
  return JA_OK;
}

// Runs the whole pipeline for `jobs` on c->stream; results (count x MsmResult) land in host memory `out`.
static int32_t msm_engine(ja_ctx* c, const ja_srs* srs, const std::vector<MsmJob>& jobs, MsmResult* out) {
  const uint32_t count = (uint32_t)jobs.size();
  if (count == 0) return JA_OK;
  bool all_indexed = true;
  for (const MsmJob& j : jobs) all_indexed = all_indexed && j.kind == MSM_INDEXED;
  if (all_indexed) return indexed_engine(c, srs, jobs, out);
  std::vector<MsmDesc> descs(count);
  std::vector<MsmWindow> wins;
  uint64_t total_n = 0, nbt = 0, e_max = 0;
  uint32_t max_segs = 1, max_nwin = 1;
  for (uint32_t m = 0; m < count; m++) {
    const MsmJob& j = jobs[m];
    MsmDesc& d = descs[m];
    d.scalars = j.d_scalars; d.n = (uint32_t)j.n; d.kind = j.kind; d.fixed_stride = 0; d.table_off = 0; d.sub = 1;
    if (j.kind == MSM_INDEXED) { d.c = 1; d.nwin = 1; d.nb = 1; }
    else if (j.kind == MSM_FR && srs->table && srs->table2_c && j.n >= table2_min_n(srs)) {
      d.c = srs->table2_c; d.nwin = srs->table2_nwin; d.nb = 1u << (srs->table2_c - 1);
      d.fixed_stride = (uint32_t)srs->n; d.table_off = srs->table2_off;
    } else if (j.kind == MSM_FR && srs->table && j.n >= table_min_n(srs)) {
      // fixed-base window table: nwin windows of c bits, ONE bucket set
      d.c = srs->table_c; d.nwin = srs->table_nwin; d.nb = 1u << (srs->table_c - 1);
      d.fixed_stride = (uint32_t)srs->n;
    } else if (j.kind == MSM_FR && srs->table && srs->table_c == 16 && getenv("JA_MSM_NO_SUB") == nullptr) {
      // SHORT job on the window table: sub-windows of c = 16 / sub bits, sub bucket sets, (sub - 1) c doublings at the end
      // (a classic windowed MSM ends in one doubling per scalar bit on ONE thread: 1.6 ms for the short folded
      // polynomials of every HyperKZG opening)
      d.c = j.n >= 256 ? 8u : 4u;
      d.sub = 16u / d.c;
      d.nwin = srs->table_nwin * d.sub;
      d.nb = 1u << (d.c - 1);
      d.fixed_stride = (uint32_t)srs->n;
    } else {
      d.c = pick_window(j.n, j.nbits);
      d.nwin = (j.nbits + 1 + d.c - 1) / d.c;
      d.nb = 1u << (d.c - 1);
    }
    d.bucket_base = (uint32_t)nbt; d.win_base = (uint32_t)wins.size();
    d.entry_base = (uint32_t)total_n; d.base_offset = (uint32_t)j.base_offset;
    const uint32_t nsets = d.fixed_stride ? d.sub : d.nwin;    // bucket sets (= window sums) of this job
    for (uint32_t w = 0; w < nsets; w++) wins.push_back(MsmWindow{(uint32_t)(nbt + (uint64_t)w * d.nb), d.nb, d.c, m});
    max_segs = std::max(max_segs, ceil_div_u32(d.nb, kSegBuckets));
    max_nwin = std::max(max_nwin, d.nwin);
    nbt += (uint64_t)nsets * d.nb;
    total_n += j.n;
    e_max += (uint64_t)j.n * d.nwin;
  }
  JA_REQUIRE(total_n < (1ull << 31) && nbt < (1ull << 31) && e_max < (1ull << 32) - 4096, "msm: batch too large for 32-bit indexing");
  if (total_n == 0) {
    for (uint32_t m = 0; m < count; m++) { memset(&out[m], 0, sizeof(MsmResult)); out[m].inf = 1; }
    return JA_OK;
  }
  const uint32_t T = run_length();
  const uint32_t nruns = ceil_div_u32(e_max, T);
  const uint32_t nwins = (uint32_t)wins.size();

  // one workspace allocation, carved up (all sub-buffers 128 B aligned)
  auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t off = 0;
  const size_t o_desc = off; off = align(off + sizeof(MsmDesc) * count);
  const size_t o_wins = off; off = align(off + sizeof(MsmWindow) * nwins);
  const size_t o_offsets = off; off = align(off + sizeof(uint32_t) * (nbt + 1));
  const size_t o_cursor = off; off = align(off + sizeof(uint32_t) * (nbt + 1));
  const size_t ntiles = (nbt + 1 + kScanTile - 1) / kScanTile;
  const size_t o_tiles = off; off = align(off + sizeof(uint32_t) * ntiles);
  const size_t o_entries = off; off = align(off + sizeof(uint32_t) * (e_max + 1));
  const size_t o_big = off; off = align(off + sizeof(uint32_t) * (nruns / kBigSpan + 2));
  const size_t o_bigcount = off; off = align(off + sizeof(uint32_t));
  const size_t o_buckets = off; off = align(off + sizeof(G1X) * nbt);
  const size_t o_head = off; off = align(off + sizeof(G1X) * nruns);
  const size_t o_tail = off; off = align(off + sizeof(G1X) * nruns);
  const size_t o_seg = off; off = align(off + sizeof(G1X) * (size_t)nwins * max_segs);
  const uint32_t max_chunks = ceil_div_u32(max_segs, kSegSpan);
  const size_t o_chunk = off; off = align(off + sizeof(G1X) * (size_t)nwins * max_chunks);
  const size_t o_wsum = off; off = align(off + sizeof(G1X) * nwins);
  const size_t o_res = off; off = align(off + sizeof(MsmResult) * count);
  char* ws = nullptr;
  int32_t st = dev_alloc(c, off, (void**)&ws);
  if (st) return st;
  MsmDesc* d_desc = (MsmDesc*)(ws + o_desc);
  MsmWindow* d_wins = (MsmWindow*)(ws + o_wins);
  uint32_t* d_offsets = (uint32_t*)(ws + o_offsets);
  uint32_t* d_cursor = (uint32_t*)(ws + o_cursor);
  uint32_t* d_tiles = (uint32_t*)(ws + o_tiles);
  uint32_t* d_entries = (uint32_t*)(ws + o_entries);
  uint32_t* d_big = (uint32_t*)(ws + o_big);
  uint32_t* d_bigcount = (uint32_t*)(ws + o_bigcount);
  G1X* d_buckets = (G1X*)(ws + o_buckets);
  G1X* d_head = (G1X*)(ws + o_head);
  G1X* d_tail = (G1X*)(ws + o_tail);
  G1X* d_seg = (G1X*)(ws + o_seg);
  G1X* d_chunk = (G1X*)(ws + o_chunk);
  G1X* d_wsum = (G1X*)(ws + o_wsum);
  MsmResult* d_res = (MsmResult*)(ws + o_res);

  cudaStream_t s = c->stream;
  const G1Aff* bases = srs->table ? srs->table : srs->points;     // table[0 .. n) is the SRS itself
  // JA_MSM_PROFILE=1: per-stage CUDA-event timings on stderr (tuning aid; adds synchronisation)
  const bool prof = getenv("JA_MSM_PROFILE") != nullptr;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  auto stage = [&](const char* name) { if (!prof) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); marks.push_back({name, e}); };
#define STAGE(name) stage(name)
  stage("begin");
  JA_CUDA(cudaMemcpyAsync(d_desc, descs.data(), sizeof(MsmDesc) * count, cudaMemcpyHostToDevice, s));
  JA_CUDA(cudaMemcpyAsync(d_wins, wins.data(), sizeof(MsmWindow) * nwins, cudaMemcpyHostToDevice, s));
  JA_CUDA(cudaMemsetAsync(d_offsets, 0, sizeof(uint32_t) * (nbt + 1), s));
  JA_CUDA(cudaMemsetAsync(d_bigcount, 0, sizeof(uint32_t), s));
  const uint32_t nthreads_n = (uint32_t)total_n;
  // plain atomics in the digit passes when every job is declared pseudo-random (MsmJob::dense_random); JA_MSM_AGG=1 forces
  // the aggregated form, JA_MSM_PLAIN=1 the plain one for any all-MSM_FR batch
  bool plain = getenv("JA_MSM_AGG") == nullptr;
  const bool force_plain = getenv("JA_MSM_PLAIN") != nullptr;
  for (const MsmJob& j : jobs) plain = plain && j.kind == MSM_FR && (j.dense_random || force_plain);
  JA_LAUNCH(c, KC_MSM_SORT, k_msm_digits<false><<<ceil_div_u32(nthreads_n, 256), 256, 0, s>>>(d_desc, count, nthreads_n, max_nwin, d_offsets, nullptr, plain));
  STAGE("hist");
  JA_LAUNCH(c, KC_MSM_SORT, k_scan_tiles<<<(unsigned)ntiles, kScanBlock, 0, s>>>(d_offsets, nbt + 1, d_tiles));
  JA_LAUNCH(c, KC_MSM_SORT, k_scan_top<<<1, kScanBlock, 0, s>>>(d_tiles, ntiles));
  JA_LAUNCH(c, KC_MSM_SORT, k_scan_add<<<(unsigned)ntiles, kScanBlock, 0, s>>>(d_offsets, nbt + 1, d_tiles));
  JA_CUDA(cudaMemcpyAsync(d_cursor, d_offsets, sizeof(uint32_t) * (nbt + 1), cudaMemcpyDeviceToDevice, s));
  STAGE("scan");
  JA_LAUNCH(c, KC_MSM_SORT, k_msm_digits<true><<<ceil_div_u32(nthreads_n, 256), 256, 0, s>>>(d_desc, count, nthreads_n, max_nwin, d_cursor, d_entries, plain));
  STAGE("scatter");
  {
    int occ = 4;
    if (const char* e = getenv("JA_MSM_OCC")) occ = atoi(e);
    const unsigned g = ceil_div_u32(nruns, 128);
    if (occ <= 4) JA_LAUNCH(c, KC_MSM_ACCUMULATE, k_msm_accumulate<4><<<g, 128, 0, s>>>(d_offsets, (uint32_t)nbt, d_entries, bases, T, d_buckets, d_head, d_tail));
    else if (occ == 5) JA_LAUNCH(c, KC_MSM_ACCUMULATE, k_msm_accumulate<5><<<g, 128, 0, s>>>(d_offsets, (uint32_t)nbt, d_entries, bases, T, d_buckets, d_head, d_tail));
    else JA_LAUNCH(c, KC_MSM_ACCUMULATE, k_msm_accumulate<6><<<g, 128, 0, s>>>(d_offsets, (uint32_t)nbt, d_entries, bases, T, d_buckets, d_head, d_tail));
  }
  STAGE("accumulate");
  JA_LAUNCH(c, KC_MSM_REDUCE, k_msm_combine<<<ceil_div_u32(nbt, 128), 128, 0, s>>>(d_offsets, (uint32_t)nbt, T, d_head, d_tail, d_buckets, d_big,
                                                      d_bigcount));
  JA_LAUNCH(c, KC_MSM_REDUCE, k_msm_combine_big<<<kSMs * 2, 128, 0, s>>>(d_offsets, T, d_head, d_tail, d_buckets, d_big, d_bigcount));
  STAGE("combine");
  dim3 g_red(ceil_div_u32(max_segs, 128), nwins);
  JA_LAUNCH(c, KC_MSM_REDUCE, k_msm_bucket_reduce<<<g_red, 128, 0, s>>>(d_wins, d_buckets, max_segs, d_seg));
  JA_LAUNCH(c, KC_MSM_REDUCE, k_msm_part_sum<<<dim3(max_chunks, nwins), 128, 0, s>>>(d_wins, d_seg, max_segs, kSegBuckets, kSegSpan, d_chunk, max_chunks));
  JA_LAUNCH(c, KC_MSM_REDUCE, k_msm_part_sum<<<dim3(1, nwins), 128, 0, s>>>(d_wins, d_chunk, max_chunks, kSegBuckets * kSegSpan, 0xffffffffu, d_wsum, 1));
  STAGE("bucket_reduce");
  JA_LAUNCH(c, KC_MSM_REDUCE, k_msm_final<<<ceil_div_u32(count, 32), 32, 0, s>>>(d_desc, count, d_wsum, d_res));
  STAGE("final");
  JA_CUDA(cudaGetLastError());
  JA_CUDA(cudaMemcpyAsync(out, d_res, sizeof(MsmResult) * count, cudaMemcpyDeviceToHost, s));
  JA_CUDA(cudaStreamSynchronize(s));
  if (prof) {
    fprintf(stderr, "[msm] n=%llu nbt=%llu e_max=%llu T=%u:", (unsigned long long)total_n, (unsigned long long)nbt, (unsigned long long)e_max, T);
    for (size_t i = 1; i < marks.size(); i++) { float ms = 0; cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second); fprintf(stderr, " %s=%.3f", marks[i].first, ms); }
    fprintf(stderr, "\n");
    for (auto& m : marks) cudaEventDestroy(m.second);
  }
#undef STAGE
  dev_free(c, ws);
  return JA_OK;
}

static void store_result(const MsmResult& r, uint64_t* out_xy, int32_t* is_inf) {
  memcpy(out_xy, r.x.l, 32);
  memcpy(out_xy + 4, r.y.l, 32);
  if (is_inf) *is_inf = (int32_t)r.inf;
}

static inline size_t kKindBytesFwd(uint32_t kind) { const size_t b[8] = {32, 1, 2, 4, 8, 4, 8, 8}; return b[kind & 7]; }
// index-range slice [lo, hi) of a job of n pairs for shard i of k (shard.cu): balanced, contiguous
static inline void shard_range(size_t n, uint32_t i, uint32_t k, size_t* lo, size_t* hi) {
  *lo = n * i / k; *hi = n * (i + 1) / k;
}

int32_t ja_msm_run(ja_ctx* c, const ja_srs* srs, const std::vector<MsmJob>& jobs_in, uint64_t* out_xy, int32_t* is_inf) {
  std::vector<MsmJob> jobs = jobs_in;
  if (c->msm_shard_count > 1) {
    // every GPU multiplies its index range only; the results are PARTIAL points the caller all-gathers and adds
    for (MsmJob& j : jobs) {
      size_t lo, hi;
      shard_range(j.n, c->msm_shard_index, c->msm_shard_count, &lo, &hi);
      const size_t bytes = j.kind == 7 ? 8 : kKindBytesFwd(j.kind);
      j.d_scalars = (const char*)j.d_scalars + lo * bytes;
      if (j.kind != 7) j.base_offset += lo;
      j.n = hi - lo;
    }
  }
  std::vector<MsmResult> res(jobs.size());
  int32_t st = msm_engine(c, srs, jobs, res.data());
  if (st) return st;
  for (size_t i = 0; i < jobs.size(); i++) store_result(res[i], out_xy + 8 * i, is_inf ? is_inf + i : nullptr);
  // with a communicator the partial points of the index ranges are exchanged and added here: every rank returns the full results
  if (c->msm_shard_count > 1 && c->comm && c->comm_world == c->msm_shard_count) {
    std::vector<int32_t> flags(jobs.size());
    for (size_t i = 0; i < jobs.size(); i++) flags[i] = (int32_t)res[i].inf;
    if ((st = comm_combine_points(c, out_xy, flags.data(), jobs.size()))) return st;
    if (is_inf) memcpy(is_inf, flags.data(), sizeof(int32_t) * jobs.size());
  }
  return JA_OK;
}

static const uint32_t kKindBits[8] = {254, 8, 16, 32, 64, 32, 64, 1};
static const uint32_t kKindBytes[8] = {32, 1, 2, 4, 8, 4, 8, 8};

extern "C" {

int32_t ja_srs_upload(ja_ctx* c, const uint64_t* g1_affine_xy, size_t n_points, ja_srs** out) {
  JA_REQUIRE(c && g1_affine_xy && out && n_points > 0, "ja_srs_upload: null or empty argument");
  JA_REQUIRE(n_points < (size_t(1) << 31), "ja_srs_upload: SRS too large");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  ja_srs* s = new ja_srs();
  s->n = n_points;
  cudaError_t e = cudaMalloc((void**)&s->points, n_points * sizeof(G1Aff));
  if (e != cudaSuccess) { delete s; return fail(JA_ERR_CUDA, std::string("ja_srs_upload: ") + cudaGetErrorString(e)); }
  JA_CUDA(cudaMemcpyAsync(s->points, g1_affine_xy, n_points * sizeof(G1Aff), cudaMemcpyHostToDevice, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));
  *out = s;
  return JA_OK;
}

int32_t ja_srs_generate(ja_ctx* c, const uint64_t g1_xy[8], const uint64_t beta[4], size_t n_points, ja_srs** out) {
  JA_REQUIRE(c && g1_xy && beta && out && n_points > 0, "ja_srs_generate: null or empty argument");
  JA_REQUIRE(n_points < (size_t(1) << 31), "ja_srs_generate: SRS too large");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  ja_srs* s = new ja_srs();
  s->n = n_points;
  cudaError_t e = cudaMalloc((void**)&s->points, n_points * sizeof(G1Aff));
  if (e != cudaSuccess) { delete s; return fail(JA_ERR_CUDA, std::string("ja_srs_generate: ") + cudaGetErrorString(e)); }
  G1Aff* table = nullptr;
  int32_t st = dev_alloc(c, 256 * sizeof(G1Aff), (void**)&table);
  if (st) { cudaFree(s->points); delete s; return st; }
  G1Aff g; memcpy(g.x.l, g1_xy, 32); memcpy(g.y.l, g1_xy + 4, 32);
  Fr b; memcpy(b.l, beta, 32);
  JA_LAUNCH(c, KC_SRS, k_srs_table<<<1, 1, 0, c->stream>>>(g, table));
  JA_LAUNCH(c, KC_SRS, k_srs_powers<<<ceil_div_u32(n_points, 128), 128, 0, c->stream>>>(table, b, (uint32_t)n_points, s->points));
  JA_CUDA(cudaGetLastError());
  JA_CUDA(cudaStreamSynchronize(c->stream));
  dev_free(c, table);
  *out = s;
  return JA_OK;
}

int32_t ja_srs_to_host(ja_ctx* c, const ja_srs* s, size_t first, size_t count, uint64_t* out_xy) {
  JA_REQUIRE(c && s && out_xy, "ja_srs_to_host: null argument");
  JA_REQUIRE(first + count <= s->n, "ja_srs_to_host: range outside the SRS");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  JA_CUDA(cudaMemcpyAsync(out_xy, s->points + first, count * sizeof(G1Aff), cudaMemcpyDeviceToHost, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));
  return JA_OK;
}

size_t ja_srs_len(const ja_srs* s) { return s ? s->n : 0; }

int32_t ja_srs_precompute(ja_ctx* c, ja_srs* s) {
  JA_REQUIRE(c && s, "ja_srs_precompute: null argument");
  if (s->table) return JA_OK;
  const uint32_t tc = 16, tw = (254 + 1 + tc - 1) / tc;
  const uint32_t tc2 = table2_window_bits(s->n), tw2 = tc2 ? (254 + 1 + tc2 - 1) / tc2 : 0;
  JA_REQUIRE(tw <= kMaxFixedWindows && tw2 <= kMaxFixedWindows && (uint64_t)s->n * (tw + tw2) < (1ull << 31),
             "ja_srs_precompute: SRS too large for 31-bit table indices");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  cudaError_t e = cudaMalloc((void**)&s->table, (size_t)s->n * (tw + tw2) * sizeof(G1Aff));
  if (e != cudaSuccess) { s->table = nullptr; cudaGetLastError(); return fail(JA_ERR_CUDA, std::string("ja_srs_precompute: ") + cudaGetErrorString(e)); }
  s->table_c = tc; s->table_nwin = tw;
  JA_LAUNCH(c, KC_SRS, k_srs_window_table<<<ceil_div_u32(s->n, 128), 128, 0, c->stream>>>(s->points, (uint32_t)s->n, s->table, (int)tc, (int)tw));
  if (tc2) {
    s->table2_c = tc2; s->table2_nwin = tw2; s->table2_off = (uint32_t)(s->n * tw);
    JA_LAUNCH(c, KC_SRS, k_srs_window_table<<<ceil_div_u32(s->n, 128), 128, 0, c->stream>>>(s->points, (uint32_t)s->n, s->table + s->table2_off, (int)tc2, (int)tw2));
  }
  JA_CUDA(cudaGetLastError());
  JA_CUDA(cudaStreamSynchronize(c->stream));
  return JA_OK;
}

void ja_srs_free(ja_ctx* c, ja_srs* s) {
  if (!c || !s) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaFree(s->table);
  cudaFree(s->points);
  delete s;
}

int32_t ja_msm_fr_batch(ja_ctx* c, const ja_srs* srs, const ja_poly* const* polys, size_t count, uint64_t* out_xy,
                        int32_t* is_inf) {
  JA_REQUIRE(c && srs && (polys || count == 0) && out_xy, "ja_msm_fr_batch: null argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  std::vector<MsmJob> jobs(count);
  for (size_t i = 0; i < count; i++) {
    JA_REQUIRE(polys[i], "ja_msm_fr_batch: null polynomial");
    if (polys[i]->len > srs->n)
      return fail(JA_ERR_KEY_LENGTH, "KeyLengthError: SRS has " + std::to_string(srs->n) + " powers, polynomial needs " +
                                         std::to_string(polys[i]->len));
    jobs[i] = MsmJob{polys[i]->data(), polys[i]->len, MSM_FR, 254, 0};
  }
  // through the sharded runner: with ja_set_msm_shard every rank multiplies its index range, and with the library communicator
  // the partial points are exchanged and added before the call returns (every rank holds the full results)
  std::vector<int32_t> flags(count ? count : 1);
  int32_t st = ja_msm_run(c, srs, jobs, out_xy, flags.data());
  if (st) return st;
  if (is_inf) memcpy(is_inf, flags.data(), sizeof(int32_t) * count);
  return JA_OK;
}

int32_t ja_msm_fr(ja_ctx* c, const ja_srs* srs, const ja_poly* scalars, uint64_t out_xy[8], int32_t* is_inf) {
  const ja_poly* arr[1] = {scalars};
  return ja_msm_fr_batch(c, srs, arr, 1, out_xy, is_inf);
}

int32_t ja_msm_host(ja_ctx* c, const ja_srs* srs, size_t base_offset, const void* scalars, int32_t width_tag, size_t n,
                    uint64_t out_xy[8], int32_t* is_inf) {
  JA_REQUIRE(c && srs && (scalars || n == 0) && out_xy, "ja_msm_host: null argument");
  JA_REQUIRE(width_tag >= 0 && width_tag <= 6, "ja_msm_host: bad width tag");
  if (base_offset + n > srs->n)
    return fail(JA_ERR_KEY_LENGTH, "KeyLengthError: SRS has " + std::to_string(srs->n) + " powers, MSM needs " +
                                       std::to_string(base_offset + n));
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  MsmResult res;
  if (n == 0) { memset(&res, 0, sizeof(res)); res.inf = 1; store_result(res, out_xy, is_inf); return JA_OK; }
  void* d = nullptr;
  const size_t bytes = n * kKindBytes[width_tag];
  int32_t st = dev_alloc(c, bytes, &d);
  if (st) return st;
  JA_CUDA(cudaMemcpyAsync(d, scalars, bytes, cudaMemcpyHostToDevice, c->stream));
  std::vector<MsmJob> jobs{MsmJob{d, n, (uint32_t)width_tag, kKindBits[width_tag], base_offset}};
  st = msm_engine(c, srs, jobs, &res);
  dev_free(c, d);
  if (st) return st;
  store_result(res, out_xy, is_inf);
  return JA_OK;
}

int32_t ja_g1_sum_indexed_batch(ja_ctx* c, const ja_srs* srs, const uint64_t* indices, const uint64_t* offsets,
                                size_t count, uint64_t* out_xy, int32_t* is_inf) {
  JA_REQUIRE(c && srs && offsets && out_xy, "ja_g1_sum_indexed_batch: null argument");
  if (count == 0) return JA_OK;
  const size_t total = offsets[count];
  JA_REQUIRE(indices || total == 0, "ja_g1_sum_indexed_batch: null indices");
  for (size_t i = 0; i < total; i++)
    if (indices[i] >= srs->n)
      return fail(JA_ERR_KEY_LENGTH, "KeyLengthError: SRS has " + std::to_string(srs->n) + " powers, index " +
                                         std::to_string(indices[i]) + " requested");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  uint64_t* d = nullptr;
  int32_t st = dev_alloc(c, (total ? total : 1) * 8, (void**)&d);
  if (st) return st;
  if (total) JA_CUDA(cudaMemcpyAsync(d, indices, total * 8, cudaMemcpyHostToDevice, c->stream));
  std::vector<MsmJob> jobs(count);
  for (size_t i = 0; i < count; i++) {
    JA_REQUIRE(offsets[i + 1] >= offsets[i], "ja_g1_sum_indexed_batch: offsets must be non-decreasing");
    jobs[i] = MsmJob{d + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), MSM_INDEXED, 1, 0};
  }
  std::vector<MsmResult> res(count);
  st = msm_engine(c, srs, jobs, res.data());
  dev_free(c, d);
  if (st) return st;
  for (size_t i = 0; i < count; i++) store_result(res[i], out_xy + 8 * i, is_inf ? is_inf + i : nullptr);
  return JA_OK;
}

// Device-resident batch of one-hot index lists (the committed OneHotPolynomials of a proof stay in HBM between the
// commitment and the opening reduction).
int32_t ja_onehot_upload(ja_ctx* c, const uint64_t* indices, const uint64_t* offsets, size_t count, ja_onehot** out) {
  JA_REQUIRE(c && offsets && out && count > 0, "ja_onehot_upload: null or empty argument");
  const size_t total = offsets[count];
  JA_REQUIRE(indices || total == 0, "ja_onehot_upload: null indices");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  ja_onehot* h = new ja_onehot();
  h->offsets.assign(offsets, offsets + count + 1);
  for (size_t i = 0; i < count; i++)
    if (offsets[i + 1] < offsets[i]) { delete h; return fail(JA_ERR_INVALID, "ja_onehot_upload: offsets must be non-decreasing"); }
  for (size_t i = 0; i < total; i++) if (indices[i] > h->max_index) h->max_index = indices[i];
  int32_t st = dev_alloc(c, (total ? total : 1) * 8, (void**)&h->d_indices);
  if (st) { delete h; return st; }
  if (total) JA_CUDA(cudaMemcpyAsync(h->d_indices, indices, total * 8, cudaMemcpyHostToDevice, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));
  *out = h;
  return JA_OK;
}

int32_t ja_onehot_commit(ja_ctx* c, const ja_srs* srs, const ja_onehot* h, uint64_t* out_xy, int32_t* is_inf) {
  JA_REQUIRE(c && srs && h && out_xy, "ja_onehot_commit: null argument");
  const size_t count = h->offsets.size() - 1;
  if (h->offsets[count] && h->max_index >= srs->n)
    return fail(JA_ERR_KEY_LENGTH, "KeyLengthError: SRS has " + std::to_string(srs->n) + " powers, index " +
                                       std::to_string(h->max_index) + " requested");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  std::vector<MsmJob> jobs(count);
  for (size_t i = 0; i < count; i++)
    jobs[i] = MsmJob{h->d_indices + h->offsets[i], (size_t)(h->offsets[i + 1] - h->offsets[i]), MSM_INDEXED, 1, 0};
  return ja_msm_run(c, srs, jobs, out_xy, is_inf);
}

void ja_onehot_free(ja_ctx* c, ja_onehot* h) {
  if (!c || !h) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  dev_free(c, h->d_indices);
  delete h;
}

int32_t ja_g1_sum_indexed(ja_ctx* c, const ja_srs* srs, const uint64_t* indices, size_t n, uint64_t out_xy[8],
                          int32_t* is_inf) {
  const uint64_t offs[2] = {0, n};
  return ja_g1_sum_indexed_batch(c, srs, indices, offs, 1, out_xy, is_inf);
}

}  // extern "C"
