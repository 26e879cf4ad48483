// C ABI implementation (include/jolt_atlas_b200.h): handle management, launch geometry, error mapping.
// Kernels live in poly_kernels.cuh / msm_kernels.cuh.  No CPU fallback anywhere in this file.
#include <memory>
#include "common.hpp"
#include "poly_kernels.cuh"
#include "aux_bodies.cuh"

static thread_local std::string g_last_error;
std::string& ja_err_slot() { return g_last_error; }

static const char* const kClassNames[KC_COUNT] = {
    "bind", "round_eval_split_eq", "round_eval_product", "round_eval_dot", "round_sum", "eq_table", "tensor_fold",
    "convert_gather", "msm_sort", "msm_accumulate", "msm_reduce", "onehot_point_sum", "hkzg_univariate_eval", "hkzg_lincomb", "hkzg_witness",
    "srs_generate", "sumcheck_fused", "scatter_add", "misc"};

void ja_prof_pre(ja_ctx* c, int cls) {
  c->launches++;
  if (!c->prof_on) return;
  const size_t i = c->prof_cls.size();
  while (c->prof_events.size() < 2 * (i + 1)) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { c->prof_on = false; return; }
    c->prof_events.push_back(e);
  }
  c->prof_cls.push_back(cls);
  cudaEventRecord(c->prof_events[2 * i], c->stream);
}
void ja_prof_post(ja_ctx* c) {
  if (!c->prof_on || c->prof_cls.empty()) return;
  cudaEventRecord(c->prof_events[2 * (c->prof_cls.size() - 1) + 1], c->stream);
}

extern "C" {

int32_t ja_profile_begin(ja_ctx* c) {
  JA_REQUIRE(c, "ja_profile_begin: null ctx");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  c->prof_cls.clear();
  c->prof_on = true;
  return JA_OK;
}

int32_t ja_profile_end(ja_ctx* c, uint64_t* out_launches, double* out_ms, size_t n_classes) {
  JA_REQUIRE(c && out_launches && out_ms, "ja_profile_end: null argument");
  JA_REQUIRE(n_classes >= (size_t)KC_COUNT, "ja_profile_end: output arrays shorter than ja_profile_class_count()");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  c->prof_on = false;
  JA_CUDA(cudaStreamSynchronize(c->stream));
  for (size_t k = 0; k < n_classes; k++) { out_launches[k] = 0; out_ms[k] = 0; }
  for (size_t i = 0; i < c->prof_cls.size(); i++) {
    float ms = 0;
    JA_CUDA(cudaEventElapsedTime(&ms, c->prof_events[2 * i], c->prof_events[2 * i + 1]));
    out_launches[c->prof_cls[i]]++;
    out_ms[c->prof_cls[i]] += ms;
  }
  c->prof_cls.clear();
  return JA_OK;
}

int32_t ja_profile_class_count(void) { return KC_COUNT; }
const char* ja_profile_class_name(int32_t k) { return k >= 0 && k < KC_COUNT ? kClassNames[k] : ""; }

void ja_last_error(char* buf, size_t cap) {
  if (!buf || !cap) return;
  snprintf(buf, cap, "%s", g_last_error.c_str());
}

int32_t ja_init(int32_t device, ja_ctx** out) {
  if (!out) return fail(JA_ERR_INVALID, "ja_init: out is null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(JA_ERR_NO_DEVICE, std::string("ja_init: no CUDA device (") + cudaGetErrorString(e) +
                                      "); this library has no CPU fallback");
  if (device < 0 || device >= n) return fail(JA_ERR_INVALID, "ja_init: bad device index");
  JA_CUDA(cudaSetDevice(device));
  ja_ctx* c = new ja_ctx();
  c->device = device;
  JA_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  cudaMemPool_t pool;
  JA_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thresh = UINT64_MAX;
  JA_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  JA_CUDA(cudaMalloc(&c->d_partials, sizeof(Fr) * kMaxGrid * kMaxOut));
  JA_CUDA(cudaMalloc(&c->d_counter, 64 * sizeof(unsigned int)));
  JA_CUDA(cudaMemset(c->d_counter, 0, 64 * sizeof(unsigned int)));
  JA_CUDA(cudaMalloc(&c->d_out, sizeof(Fr) * kMaxOut));
  JA_CUDA(cudaHostAlloc((void**)&c->h_pinned, kPinnedBytes, cudaHostAllocMapped));   // device-addressable: k_collect_finals stores into it
  JA_CUDA(cudaMallocHost((void**)&c->h_ring, kRingBytes));
  JA_CUDA(cudaHostAlloc(&c->h_mapped, kSlots * kSlotBytes, cudaHostAllocMapped));
  memset(c->h_mapped, 0, kSlots * kSlotBytes);
  JA_CUDA(cudaHostGetDevicePointer(&c->d_mapped, c->h_mapped, 0));
  JA_CUDA(cudaHostAlloc(&c->h_mail, kMailEntries * 16, cudaHostAllocMapped));
  memset(c->h_mail, 0, kMailEntries * 16);
  JA_CUDA(cudaHostGetDevicePointer(&c->d_mail, c->h_mail, 0));
  JA_CUDA(cudaMalloc((void**)&c->d_mail_dev, kMailEntries * 20));      // 16-byte twins, then one relay-election word per entry
  JA_CUDA(cudaMemset(c->d_mail_dev, 0, kMailEntries * 20));
  JA_CUDA(cudaHostAlloc(&c->h_rrmail, kRrEntries * 16, cudaHostAllocMapped));
  memset(c->h_rrmail, 0, kRrEntries * 16);
  JA_CUDA(cudaHostGetDevicePointer(&c->d_rrmail, c->h_rrmail, 0));
  JA_CUDA(cudaMalloc((void**)&c->d_rrrelay, kRrEntries * 16));
  JA_CUDA(cudaMemset(c->d_rrrelay, 0, kRrEntries * 16));
  JA_CUDA(cudaHostAlloc(&c->h_rowvals, kRowSeqOffset + 64, cudaHostAllocMapped));
  memset(c->h_rowvals, 0, kRowSeqOffset + 64);
  JA_CUDA(cudaHostGetDevicePointer(&c->d_rowvals, c->h_rowvals, 0));
  JA_CUDA(cudaEventCreate(&c->ev0));
  JA_CUDA(cudaEventCreate(&c->ev1));
  *out = c;
  return JA_OK;
}

void ja_shutdown(ja_ctx* c) {
  if (!c) return;
  ja_comm_free(c);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (EqTableCacheEntry& e : c->eq_cache) { dev_free(c, e.out_levels); dev_free(c, e.in_levels); e.out_levels = e.in_levels = nullptr; }
  dev_cache_release(c);
  cudaFree(c->d_partials); cudaFree(c->d_counter); cudaFree(c->d_out);
  cudaFreeHost(c->h_pinned);
  cudaFreeHost(c->h_ring);
  cudaFreeHost(c->h_mapped);
  cudaFreeHost(c->h_rowvals);
  cudaFreeHost(c->h_mail);
  cudaFree(c->d_mail_dev);
  cudaFreeHost(c->h_rrmail);
  cudaFree(c->d_rrrelay);
  for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
  cudaStreamDestroy(c->stream);
  delete c;
}

int32_t ja_sync(ja_ctx* c) {
  JA_REQUIRE(c, "ja_sync: null ctx");
  JA_CUDA(cudaStreamSynchronize(c->stream));
  return JA_OK;
}

uint64_t ja_launch_count(const ja_ctx* c) { return c ? c->launches : 0; }

// ---- polynomials --------------------------------------------------------------------------------
int32_t ja_poly_alloc(ja_ctx* c, size_t n, ja_poly** out) {
  JA_REQUIRE(c && out, "ja_poly_alloc: null argument");
  JA_REQUIRE(is_pow2(n), "ja_poly_alloc: length must be a power of two");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  ja_poly* p = new ja_poly();
  p->len = n;
  int32_t st = dev_alloc(c, n * sizeof(Fr), (void**)&p->buf[0]);
  if (st) { delete p; return st; }
  p->cap[0] = n;
  *out = p;
  return JA_OK;
}

int32_t ja_poly_from_fr(ja_ctx* c, const uint64_t* z, size_t n, ja_poly** out) {
  JA_REQUIRE(c && z && out, "ja_poly_from_fr: null argument");
  int32_t st = ja_poly_alloc(c, n, out);
  if (st) return st;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaMemcpyAsync((*out)->buf[0], z, n * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));   // host buffer is only borrowed for the call
  return JA_OK;
}

int32_t ja_poly_from_i32(ja_ctx* c, const int32_t* z, size_t n, ja_poly** out) {
  JA_REQUIRE(c && z && out, "ja_poly_from_i32: null argument");
  int32_t st = ja_poly_alloc(c, n, out);
  if (st) return st;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  int* tmp = nullptr;
  st = dev_alloc(c, n * sizeof(int), (void**)&tmp);
  if (st) return st;
  JA_CUDA(cudaMemcpyAsync(tmp, z, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  JA_LAUNCH(c, KC_CONVERT, k_i32_to_fr<<<grid_for(n), kBlock, 0, c->stream>>>(tmp, (*out)->buf[0], n));
  JA_CUDA(cudaGetLastError());
  dev_free(c, tmp);
  JA_CUDA(cudaStreamSynchronize(c->stream));
  return JA_OK;
}

// the same for `count` polynomials of n coefficients each (row-major): one copy, one synchronisation
int32_t ja_poly_from_i32_many(ja_ctx* c, const int32_t* z, size_t count, size_t n, ja_poly** out) {
  JA_REQUIRE(c && z && out && count >= 1, "ja_poly_from_i32_many: null argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  int* tmp = nullptr;
  int32_t st = dev_alloc(c, count * n * sizeof(int), (void**)&tmp);
  if (st) return st;
  for (size_t i = 0; i < count; i++) out[i] = nullptr;
  for (size_t i = 0; i < count && !st; i++) st = ja_poly_alloc(c, n, &out[i]);
  if (st) { for (size_t i = 0; i < count; i++) if (out[i]) { ja_poly_free(c, out[i]); out[i] = nullptr; } dev_free(c, tmp); return st; }
  JA_CUDA(cudaMemcpyAsync(tmp, z, count * n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  for (size_t i = 0; i < count; i++)
    JA_LAUNCH(c, KC_CONVERT, k_i32_to_fr<<<grid_for(n), kBlock, 0, c->stream>>>(tmp + i * n, out[i]->buf[0], n));
  JA_CUDA(cudaGetLastError());
  dev_free(c, tmp);
  JA_CUDA(cudaStreamSynchronize(c->stream));      // the host array is borrowed for the duration of the call
  return JA_OK;
}

int32_t ja_poly_from_lookup(ja_ctx* c, const uint64_t* table, size_t K, const uint32_t* idx, size_t n, ja_poly** out) {
  JA_REQUIRE(c && table && idx && out && K > 0, "ja_poly_from_lookup: null argument");
  for (size_t i = 0; i < n; i++)
    JA_REQUIRE(idx[i] == 0xffffffffu || idx[i] < K, "ja_poly_from_lookup: index outside the table");
  int32_t st = ja_poly_alloc(c, n, out);
  if (st) return st;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  Fr* d_table = nullptr; uint32_t* d_idx = nullptr;
  if ((st = dev_alloc(c, K * sizeof(Fr), (void**)&d_table))) return st;
  if ((st = dev_alloc(c, n * sizeof(uint32_t), (void**)&d_idx))) return st;
  JA_CUDA(cudaMemcpyAsync(d_table, table, K * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  JA_CUDA(cudaMemcpyAsync(d_idx, idx, n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  JA_LAUNCH(c, KC_CONVERT, k_gather_small_table<<<grid_for(n), kBlock, 0, c->stream>>>(d_table, d_idx, n, (*out)->buf[0]));
  JA_CUDA(cudaGetLastError());
  dev_free(c, d_table); dev_free(c, d_idx);
  JA_CUDA(cudaStreamSynchronize(c->stream));
  return JA_OK;
}

int32_t ja_poly_clone(ja_ctx* c, const ja_poly* src, ja_poly** out) {
  JA_REQUIRE(c && src && out, "ja_poly_clone: null argument");
  int32_t st = ja_poly_alloc(c, src->len, out);
  if (st) return st;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaMemcpyAsync((*out)->buf[0], src->data(), src->len * sizeof(Fr), cudaMemcpyDeviceToDevice, c->stream));
  return JA_OK;
}

size_t ja_poly_len(const ja_poly* p) { return p ? p->len : 0; }

int32_t ja_poly_to_host(ja_ctx* c, const ja_poly* p, uint64_t* out, size_t cap) {
  JA_REQUIRE(c && p && out, "ja_poly_to_host: null argument");
  JA_REQUIRE(cap >= p->len, "ja_poly_to_host: output too small");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  JA_CUDA(cudaMemcpyAsync(out, p->data(), p->len * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));
  return JA_OK;
}

void ja_poly_free(ja_ctx* c, ja_poly* p) {
  if (!c || !p) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  dev_free(c, p->buf[0]);
  dev_free(c, p->buf[1]);
  delete p;
}

void ja_poly_free_many(ja_ctx* c, ja_poly* const* polys, size_t n) {
  if (!c || !polys) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  for (size_t i = 0; i < n; i++) {
    ja_poly* p = polys[i];
    if (!p) continue;
    dev_free(c, p->buf[0]);
    dev_free(c, p->buf[1]);
    delete p;
  }
}

int32_t ja_bind_many(ja_ctx* c, ja_poly* const* polys, size_t n_polys, const uint64_t r[4], int32_t order) {
  JA_REQUIRE(c && polys && r, "ja_bind: null argument");
  JA_REQUIRE(order == JA_LOW_TO_HIGH || order == JA_HIGH_TO_LOW, "ja_bind: bad binding order");
  if (n_polys == 0) return JA_OK;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  const Challenge ch = to_challenge(r);
  JA_REQUIRE(r[0] == 0 && r[1] == 0, "ja_bind: challenge limbs 0,1 must be zero (MontU128Challenge)");
  size_t done = 0;
  while (done < n_polys) {
    // group consecutive polys of equal length into one launch
    const size_t len = polys[done]->len;
    JA_REQUIRE(len >= 2, "ja_bind: polynomial is already fully bound");
    const size_t half = len / 2;
    BindArgs args;
    size_t cnt = 0;
    while (done + cnt < n_polys && cnt < (size_t)kMaxBindPolys && polys[done + cnt]->len == len) {
      ja_poly* p = polys[done + cnt];
      JA_REQUIRE(p != nullptr, "ja_bind: null polynomial");
      if (order == JA_HIGH_TO_LOW) {
        args.in[cnt] = p->data(); args.out[cnt] = p->data();   // in place: thread i touches i and i+half only
      } else {
        const int nxt = 1 - p->cur;
        if (p->cap[nxt] < half) {
          dev_free(c, p->buf[nxt]);
          p->buf[nxt] = nullptr; p->cap[nxt] = 0;
          int32_t st = dev_alloc(c, half * sizeof(Fr), (void**)&p->buf[nxt]);
          if (st) return st;
          p->cap[nxt] = half;
        }
        args.in[cnt] = p->data(); args.out[cnt] = p->buf[nxt];
      }
      cnt++;
    }
    dim3 grid(grid_for(half), (unsigned)cnt);
    if (order == JA_LOW_TO_HIGH) JA_LAUNCH(c, KC_BIND, k_bind<true><<<grid, kBlock, 0, c->stream>>>(args, ch, half));
    else                         JA_LAUNCH(c, KC_BIND, k_bind<false><<<grid, kBlock, 0, c->stream>>>(args, ch, half));
    JA_CUDA(cudaGetLastError());
    for (size_t q = 0; q < cnt; q++) {
      ja_poly* p = polys[done + q];
      if (order == JA_LOW_TO_HIGH) p->cur = 1 - p->cur;
      p->len = half;
    }
    done += cnt;
  }
  return JA_OK;
}

int32_t ja_bind(ja_ctx* c, ja_poly* p, const uint64_t r[4], int32_t order) {
  JA_REQUIRE(p, "ja_bind: null polynomial");
  ja_poly* arr[1] = {p};
  return ja_bind_many(c, arr, 1, r, order);
}

int32_t ja_final_claim(ja_ctx* c, const ja_poly* p, uint64_t out[4]) {
  JA_REQUIRE(c && p && out, "ja_final_claim: null argument");
  JA_REQUIRE(p->len == 1, "ja_final_claim: polynomial is not fully bound (len != 1)");
  return ja_poly_to_host(c, p, out, 1);
}

// ---- eq tables -----------------------------------------------------------------------------------
static int32_t eq_evals_device(ja_ctx* c, const FrH* r, size_t m, const FrH& scale, Fr* out /* 2^m */) {
  const size_t mh = m / 2, ml = m - mh;
  Fr* d_r = nullptr;
  Fr* lv[2] = {nullptr, nullptr};
  int32_t st = dev_alloc(c, (size_t(2) << mh) * sizeof(Fr), (void**)&lv[0]);
  if (st) return st;
  st = dev_alloc(c, (size_t(2) << ml) * sizeof(Fr), (void**)&lv[1]);
  if (st) return st;
  if (ml <= (size_t)kEqValMax) {
    EqLevelsValArgs v;
    for (size_t i = 0; i < mh; i++) v.w[0][i] = to_dev(r[i]);
    for (size_t i = 0; i < ml; i++) v.w[1][i] = to_dev(r[mh + i]);
    v.m[0] = (int)mh; v.rev[0] = 0; v.buf[0] = lv[0]; v.scale[0] = to_dev(scale);
    v.m[1] = (int)ml; v.rev[1] = 0; v.buf[1] = lv[1]; v.scale[1] = to_dev(host::FR_ONE);
    JA_LAUNCH(c, KC_EQ_TABLE, k_eq_levels_val<<<2, kBlock, 0, c->stream>>>(v));
  } else {
    st = dev_alloc(c, (m ? m : 1) * sizeof(Fr), (void**)&d_r);
    if (st) return st;
    if (m && (st = stage_h2d(c, d_r, r, m * 32))) return st;
    EqLevelsArgs a;
    a.w[0] = d_r; a.m[0] = (int)mh; a.rev[0] = 0; a.buf[0] = lv[0]; a.scale[0] = to_dev(scale);
    a.w[1] = d_r + mh; a.m[1] = (int)ml; a.rev[1] = 0; a.buf[1] = lv[1]; a.scale[1] = to_dev(host::FR_ONE);
    JA_LAUNCH(c, KC_EQ_TABLE, k_eq_levels<<<2, 1024, 0, c->stream>>>(a));
  }
  JA_CUDA(cudaGetLastError());
  const size_t n = size_t(1) << m;
  JA_LAUNCH(c, KC_EQ_TABLE, k_eq_expand<<<grid_for(n), kBlock, 0, c->stream>>>(lv[0] + ((size_t(1) << mh) - 1), lv[1] + ((size_t(1) << ml) - 1),
                                                    (int)ml, n, out));
  JA_CUDA(cudaGetLastError());
  dev_free(c, d_r); dev_free(c, lv[0]); dev_free(c, lv[1]);     // stream-ordered reuse (one stream per context)
  return JA_OK;
}

}  // extern "C"
// the two half tables of eq(r, .) without the outer product: eq[x] = hi[x >> bits_lo] * lo[x & mask]; lv_* are the level
// buffers to release (dev_free) once the consumer has been enqueued
int32_t eq_halves_device_pub(ja_ctx* c, const uint64_t* r, size_t m, Fr** lv_hi, Fr** lv_lo, const Fr** hi, const Fr** lo, int* bits_lo) {
  const size_t mh = m / 2, ml = m - mh;
  if (ml > (size_t)kEqValMax) return fail(JA_ERR_UNSUPPORTED, "eq tables: too many variables for the by-value kernel");
  const FrH* rr = reinterpret_cast<const FrH*>(r);
  int32_t st = dev_alloc(c, (size_t(2) << mh) * sizeof(Fr), (void**)lv_hi);
  if (st) return st;
  if ((st = dev_alloc(c, (size_t(2) << ml) * sizeof(Fr), (void**)lv_lo))) return st;
  EqLevelsValArgs v;
  for (size_t i = 0; i < mh; i++) v.w[0][i] = to_dev(rr[i]);
  for (size_t i = 0; i < ml; i++) v.w[1][i] = to_dev(rr[mh + i]);
  v.m[0] = (int)mh; v.rev[0] = 0; v.buf[0] = *lv_hi; v.scale[0] = to_dev(host::FR_ONE);
  v.m[1] = (int)ml; v.rev[1] = 0; v.buf[1] = *lv_lo; v.scale[1] = to_dev(host::FR_ONE);
  JA_LAUNCH(c, KC_EQ_TABLE, k_eq_levels_val<<<2, kBlock, 0, c->stream>>>(v));
  JA_CUDA(cudaGetLastError());
  *hi = *lv_hi + ((size_t(1) << mh) - 1);
  *lo = *lv_lo + ((size_t(1) << ml) - 1);
  *bits_lo = (int)ml;
  return JA_OK;
}
int32_t eq_evals_device_pub(ja_ctx* c, const uint64_t* r, size_t m, Fr* out) {
  return eq_evals_device(c, reinterpret_cast<const FrH*>(r), m, host::FR_ONE, out);
}
extern "C" {

int32_t ja_eq_evals(ja_ctx* c, const uint64_t* r, size_t m, const uint64_t* scale_or_null, ja_poly** out) {
  JA_REQUIRE(c && out && (r || m == 0), "ja_eq_evals: null argument");
  JA_REQUIRE(m <= 34, "ja_eq_evals: too many variables");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  int32_t st = ja_poly_alloc(c, size_t(1) << m, out);
  if (st) return st;
  FrH scale = scale_or_null ? host::from_limbs(scale_or_null) : host::FR_ONE;
  return eq_evals_device(c, reinterpret_cast<const FrH*>(r), m, scale, (*out)->buf[0]);
}

// ---- split eq --------------------------------------------------------------------------------------
int32_t ja_spliteq_new(ja_ctx* c, const uint64_t* w, size_t m, int32_t order, const uint64_t* scale_or_null,
                       ja_spliteq** out) {
  JA_REQUIRE(c && out && (w || m == 0), "ja_spliteq_new: null argument");
  JA_REQUIRE(order == JA_LOW_TO_HIGH || order == JA_HIGH_TO_LOW, "ja_spliteq_new: bad binding order");
  JA_REQUIRE(m <= 60, "ja_spliteq_new: too many variables");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  ja_spliteq* s = new ja_spliteq();
  s->order = order; s->m = (int)m;
  s->current_scalar = scale_or_null ? host::from_limbs(scale_or_null) : host::FR_ONE;
  s->w.resize(m);
  for (size_t i = 0; i < m; i++) s->w[i] = host::from_limbs(w + 4 * i);
  // split_eq_poly.rs:95-143
  size_t n_out_vars, n_in_vars, off_out, off_in;
  if (m == 0) { n_out_vars = n_in_vars = 0; off_out = off_in = 0; s->current_index = 0; }
  else if (order == JA_LOW_TO_HIGH) {
    const size_t half = m / 2;            // w = [w_out (half) | w_in (m-1-half) | w_last]
    n_out_vars = half; n_in_vars = m - 1 - half; off_out = 0; off_in = half;
    s->current_index = (int)m;
  } else {
    const size_t half = m / 2;            // w = [w_first | w_in (half) | w_out (m-1-half)]
    n_in_vars = half; n_out_vars = m - 1 - half; off_in = 1; off_out = 1 + half;
    if (n_in_vars > m - 1) { n_in_vars = m - 1; n_out_vars = 0; }
    s->current_index = 0;
  }
  s->out_len = (int)n_out_vars + 1; s->in_len = (int)n_in_vars + 1;
  // table cache: same (order, point) -> the tables already on the device
  static const bool no_eq_cache = getenv("JA_NO_EQ_CACHE") != nullptr;
  constexpr size_t kEqCacheSlots = 6;
  std::vector<uint64_t> key;
  if (!no_eq_cache && m >= 1) {
    key.resize(2 + 4 * m);
    key[0] = (uint64_t)order; key[1] = m;
    memcpy(key.data() + 2, w, 32 * m);
    for (size_t i = 0; i < c->eq_cache.size(); i++) {
      EqTableCacheEntry& e = c->eq_cache[i];
      if (e.out_levels && e.key == key) {
        e.refs++; e.stamp = ++c->eq_cache_clock;
        s->out_levels = e.out_levels; s->in_levels = e.in_levels; s->cache_slot = (int)i;
        *out = s;
        return JA_OK;
      }
    }
  }
  int32_t st = dev_alloc(c, (size_t(2) << n_out_vars) * sizeof(Fr), (void**)&s->out_levels);
  if (st) { delete s; return st; }
  st = dev_alloc(c, (size_t(2) << n_in_vars) * sizeof(Fr), (void**)&s->in_levels);
  if (st) { delete s; return st; }
  if (!key.empty()) {
    // take a free slot, or the least recently used one nobody references (its buffers return to the allocator: stream-ordered reuse)
    int slot = -1;
    if (c->eq_cache.size() < kEqCacheSlots) { c->eq_cache.emplace_back(); slot = (int)c->eq_cache.size() - 1; }
    else {
      for (size_t i = 0; i < c->eq_cache.size(); i++)
        if (c->eq_cache[i].refs == 0 && (slot < 0 || c->eq_cache[i].stamp < c->eq_cache[slot].stamp)) slot = (int)i;
      if (slot >= 0 && c->eq_cache[slot].out_levels) { dev_free(c, c->eq_cache[slot].out_levels); dev_free(c, c->eq_cache[slot].in_levels); }
    }
    if (slot >= 0) {
      EqTableCacheEntry& e = c->eq_cache[slot];
      e.key = std::move(key); e.out_levels = s->out_levels; e.in_levels = s->in_levels; e.refs = 1; e.stamp = ++c->eq_cache_clock;
      s->cache_slot = slot;
    }
  }
  const int rev = order == JA_HIGH_TO_LOW ? 1 : 0;
  if (n_out_vars <= (size_t)kEqValMax && n_in_vars <= (size_t)kEqValMax) {
    // the point travels in the kernel parameters: no staged copy ahead of the launch
    EqLevelsValArgs v;
    for (size_t i = 0; i < n_out_vars; i++) v.w[0][i] = to_dev(s->w[off_out + i]);
    for (size_t i = 0; i < n_in_vars; i++) v.w[1][i] = to_dev(s->w[off_in + i]);
    v.m[0] = (int)n_out_vars; v.rev[0] = rev; v.buf[0] = s->out_levels; v.scale[0] = to_dev(host::FR_ONE);
    v.m[1] = (int)n_in_vars;  v.rev[1] = rev; v.buf[1] = s->in_levels;  v.scale[1] = to_dev(host::FR_ONE);
    JA_LAUNCH(c, KC_EQ_TABLE, k_eq_levels_val<<<2, kBlock, 0, c->stream>>>(v));
    JA_CUDA(cudaGetLastError());
    *out = s;
    return JA_OK;
  }
  Fr* d_w = nullptr;
  auto drop = [&](int32_t e) {        // error exit: the half-built tables must not stay in the cache
    if (s->cache_slot >= 0) { c->eq_cache[s->cache_slot].out_levels = nullptr; c->eq_cache[s->cache_slot].in_levels = nullptr; c->eq_cache[s->cache_slot].refs = 0; }
    dev_free(c, s->out_levels); dev_free(c, s->in_levels); dev_free(c, d_w);
    delete s;
    return e;
  };
  st = dev_alloc(c, (m ? m : 1) * sizeof(Fr), (void**)&d_w);
  if (st) return drop(st);
  if (m && (st = stage_h2d(c, d_w, w, m * 32))) return drop(st);
  EqLevelsArgs a;
  a.w[0] = d_w + off_out; a.m[0] = (int)n_out_vars; a.rev[0] = rev; a.buf[0] = s->out_levels; a.scale[0] = to_dev(host::FR_ONE);
  a.w[1] = d_w + off_in;  a.m[1] = (int)n_in_vars;  a.rev[1] = rev; a.buf[1] = s->in_levels;  a.scale[1] = to_dev(host::FR_ONE);
  JA_LAUNCH(c, KC_EQ_TABLE, k_eq_levels<<<2, 1024, 0, c->stream>>>(a));
  JA_CUDA(cudaGetLastError());
  dev_free(c, d_w);                 // stream-ordered reuse (one stream per context)
  *out = s;
  return JA_OK;
}

int32_t ja_spliteq_bind(ja_ctx* c, ja_spliteq* s, const uint64_t r[4]) {
  JA_REQUIRE(c && s && r, "ja_spliteq_bind: null argument");
  const FrH rr = host::from_limbs(r);
  const int n = s->m;
  // split_eq_poly.rs:331-372
  if (s->order == JA_LOW_TO_HIGH) {
    JA_REQUIRE(s->current_index >= 1, "ja_spliteq_bind: already fully bound");
    const FrH wv = s->w[s->current_index - 1];
    const FrH prod = host::mul(wv, rr);
    FrH f = host::sub(host::sub(host::FR_ONE, wv), rr);
    f = host::add(host::add(f, prod), prod);
    s->current_scalar = host::mul(s->current_scalar, f);
    s->current_index -= 1;
    if (n / 2 < s->current_index && s->in_len > 1) s->in_len--;
    else if (0 < s->current_index && s->out_len > 1) s->out_len--;
  } else {
    JA_REQUIRE(s->current_index < n, "ja_spliteq_bind: already fully bound");
    const FrH wv = s->w[s->current_index];
    const FrH prod = host::mul(wv, rr);
    FrH f = host::sub(host::sub(host::FR_ONE, wv), rr);
    f = host::add(host::add(f, prod), prod);
    s->current_scalar = host::mul(s->current_scalar, f);
    s->current_index += 1;
    if (s->current_index <= n / 2 && s->in_len > 1) s->in_len--;
    else if (s->current_index <= n && s->out_len > 1) s->out_len--;
  }
  return JA_OK;
}

int32_t ja_spliteq_current_scalar(const ja_spliteq* s, uint64_t out[4]) {
  JA_REQUIRE(s && out, "ja_spliteq_current_scalar: null argument");
  memcpy(out, s->current_scalar.l, 32);
  return JA_OK;
}

int32_t ja_spliteq_current_w(const ja_spliteq* s, uint64_t out[4]) {
  JA_REQUIRE(s && out, "ja_spliteq_current_w: null argument");
  if (s->order == JA_LOW_TO_HIGH) {
    JA_REQUIRE(s->current_index >= 1, "ja_spliteq_current_w: fully bound");
    memcpy(out, s->w[s->current_index - 1].l, 32);
  } else {
    JA_REQUIRE(s->current_index < s->m, "ja_spliteq_current_w: fully bound");
    memcpy(out, s->w[s->current_index].l, 32);
  }
  return JA_OK;
}

int32_t ja_spliteq_merge(ja_ctx* c, const ja_spliteq* s, ja_poly** out) {
  JA_REQUIRE(c && s && out, "ja_spliteq_merge: null argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  const FrH* w; size_t m;
  if (s->order == JA_LOW_TO_HIGH) { w = s->w.data(); m = (size_t)s->current_index; }
  else { w = s->w.data() + s->current_index; m = (size_t)(s->m - s->current_index); }
  int32_t st = ja_poly_alloc(c, size_t(1) << m, out);
  if (st) return st;
  return eq_evals_device(c, w, m, s->current_scalar, (*out)->buf[0]);
}

void ja_spliteq_free(ja_ctx* c, ja_spliteq* s) {
  if (!c || !s) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  if (s->cache_slot >= 0 && (size_t)s->cache_slot < c->eq_cache.size() && c->eq_cache[s->cache_slot].out_levels == s->out_levels) {
    if (c->eq_cache[s->cache_slot].refs > 0) c->eq_cache[s->cache_slot].refs--;      // the cache keeps the tables
  } else {
    dev_free(c, s->out_levels); dev_free(c, s->in_levels);
  }
  delete s;
}

}  // extern "C"

// ---- round evaluation ------------------------------------------------------------------------------
template <int KID>
static void launch_s(ja_ctx* c, const EvalPolys& P, const ja_spliteq* eq, size_t G) {
  const int bits_in = eq->in_len - 1;
  size_t tiles = (G + kBlock - 1) / kBlock;
  size_t grid = tiles < (size_t)kSMs * 4 ? tiles : (size_t)kSMs * 4;
  size_t tpb = (tiles + grid - 1) / grid;
  grid = (tiles + tpb - 1) / tpb;
  JA_LAUNCH(c, KC_ROUND_EVAL_S, k_round_eval_s<KID><<<(unsigned)grid, kBlock, 0, c->stream>>>(P, eq->e_out(), eq->e_in(), bits_in, G, tpb,
                                                                c->d_partials, c->d_counter, c->d_out,
                                                                c->slice_on ? c->slice_g_offset : 0));
}

extern "C" {

// One GPU's share of a round evaluation when every MLE is sharded into contiguous hypercube slices (shard.cu):
// `polys` hold the slice, `eq` is the replicated split-eq of the WHOLE instance, g_offset = slice_start / 2.
// Family S and PROD/POW only (LowToHigh).  out_evals are PARTIAL sums: all-gather them and add (ja_fr_sum).
int32_t ja_round_eval_slice(ja_ctx* c, int32_t kernel_id, const ja_poly* const* polys, size_t n_polys,
                            const ja_spliteq* eq, uint32_t aux_u32, size_t g_offset, uint64_t* out_evals, size_t n_out) {
  JA_REQUIRE(c && eq, "ja_round_eval_slice: null argument");
  JA_REQUIRE(kernel_id <= JA_EVAL_IDENT, "ja_round_eval_slice: only the split-eq (LowToHigh) round bodies shard by contiguous slices");
  JA_REQUIRE(kernel_id != JA_EVAL_PROD && kernel_id != JA_EVAL_POW ? true : (kernel_id == JA_EVAL_POW ? aux_u32 : n_polys) <= 16,
             "ja_round_eval_slice: product degree must be <= 16");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  c->slice_on = true; c->slice_g_offset = g_offset;
  const int32_t st = ja_round_eval(c, kernel_id, polys, n_polys, eq, nullptr, 0, aux_u32, out_evals, n_out);
  c->slice_on = false;
  return st;
}

int32_t ja_round_eval(ja_ctx* c, int32_t kernel_id, const ja_poly* const* polys, size_t n_polys,
                      const ja_spliteq* eq, const uint64_t* aux_fr, size_t n_aux, uint32_t aux_u32,
                      uint64_t* out_evals, size_t n_out) {
  JA_REQUIRE(c && out_evals, "ja_round_eval: null argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);      // launch + collect share the pinned staging buffer
  RoundEvalPending pend;
  int32_t st = ja_round_eval_launch(c, kernel_id, polys, n_polys, eq, aux_fr, n_aux, aux_u32, n_out, &pend);
  if (st) return st;
  return ja_round_eval_collect(c, pend, out_evals);
}

}  // extern "C"

// The asynchronous half: validates, launches the kernel and enqueues the D2H copy of the reduced sums into the
// context's pinned staging buffer.  The caller may do host work (e.g. the round's field inversion) before collect.
// table-weighted / eq-scheduled bodies (aux_bodies.cuh)
static int32_t ja_round_eval_w_launch(ja_ctx* c, int32_t kernel_id, const ja_poly* const* polys, size_t n_polys, uint32_t shift, size_t n_out,
                                      RoundEvalPending* pend) {
  const size_t len = polys[0]->len, half = len / 2;
  WArgs a;
  memset(&a, 0, sizeof(a));
  a.shift = shift;
  size_t want_polys = 0, want_out = 0;
  switch (kernel_id) {
    case JA_EVAL_WSUM: want_polys = 2; want_out = 1; break;
    case JA_EVAL_WDOT2: want_polys = 3; want_out = 3; break;
    case JA_EVAL_DOT2_L2H: want_polys = 2; want_out = 2; break;
    case JA_EVAL_SQ_EQHI: want_polys = 2; want_out = 3; break;
    case JA_EVAL_DOT2_EQHI: case JA_EVAL_DOT2_EQLOW: want_polys = 3; want_out = 3; break;
    default: return fail(JA_ERR_UNSUPPORTED, "ja_round_eval: kernel_id not implemented");
  }
  JA_REQUIRE(n_polys == want_polys && n_out == want_out && shift < 60, "ja_round_eval: wrong n_polys / n_out / shift for kernel_id");
  const size_t n_ops = kernel_id == JA_EVAL_DOT2_L2H ? 2 : n_polys - 1;
  for (size_t q = 0; q < n_ops; q++) a.p[q] = polys[q]->data();
  if (kernel_id != JA_EVAL_DOT2_L2H) {
    const ja_poly* t = polys[n_polys - 1];
    a.tab = t->data(); a.tab_len = t->len;
    const size_t top = (half - 1) >> shift;
    if (kernel_id == JA_EVAL_WSUM || kernel_id == JA_EVAL_WDOT2) JA_REQUIRE(top < t->len, "ja_round_eval: table shorter than (len/2) >> shift");
    else if (kernel_id == JA_EVAL_DOT2_EQLOW) JA_REQUIRE((size_t(1) << shift) <= t->len, "ja_round_eval: table shorter than 2^shift");
    else JA_REQUIRE(t->len == 1 || top < t->len / 2, "ja_round_eval: eq polynomial shorter than 2 * ((len/2) >> shift)");
  }
  unsigned grid = grid_for(half);
  if (grid > (unsigned)kSMs * 4) grid = kSMs * 4;
#define JA_W(MODE) JA_LAUNCH(c, KC_ROUND_EVAL_DOT, k_round_eval_w<MODE><<<grid, kBlock, 0, c->stream>>>(a, half, c->d_partials, c->d_counter, c->d_out))
  switch (kernel_id) {
    case JA_EVAL_WSUM: JA_W(W_SUM); break;
    case JA_EVAL_WDOT2: JA_W(W_DOT2); break;
    case JA_EVAL_DOT2_L2H: JA_W(W_DOT2_L2H); break;
    case JA_EVAL_SQ_EQHI: JA_W(W_SQ_EQHI); break;
    case JA_EVAL_DOT2_EQHI: JA_W(W_DOT2_EQHI); break;
    default: JA_W(W_DOT2_EQLOW); break;
  }
#undef JA_W
  JA_CUDA(cudaGetLastError());
  JA_CUDA(cudaMemcpyAsync(c->h_pinned, c->d_out, n_out * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  pend->n_dev = n_out; pend->n_out = n_out; pend->n_polys = n_polys; pend->fam_sum = false; pend->aux_fr = nullptr;
  pend->prod_lanes = 0; pend->prod_d = 0;
  return JA_OK;
}

int32_t ja_round_eval_launch(ja_ctx* c, int32_t kernel_id, const ja_poly* const* polys, size_t n_polys,
                             const ja_spliteq* eq, const uint64_t* aux_fr, size_t n_aux, uint32_t aux_u32, size_t n_out,
                             RoundEvalPending* pend) {
  JA_REQUIRE(c && polys && pend, "ja_round_eval: null argument");
  JA_REQUIRE(n_polys >= 1 && n_polys <= (size_t)kMaxProdPolys, "ja_round_eval: unsupported number of polynomials");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  const size_t len = polys[0]->len;
  JA_REQUIRE(len >= 2, "ja_round_eval: polynomial already fully bound");
  // table-weighted bodies: the last polynomial is the (shorter) table / eq polynomial
  const bool has_table = kernel_id == JA_EVAL_WIDENT || kernel_id == JA_EVAL_WSUM || kernel_id == JA_EVAL_WDOT2 || kernel_id == JA_EVAL_SQ_EQHI ||
                         kernel_id == JA_EVAL_DOT2_EQHI || kernel_id == JA_EVAL_DOT2_EQLOW;
  for (size_t i = 0; i < n_polys; i++)
    JA_REQUIRE(polys[i] && (polys[i]->len == len || (has_table && i + 1 == n_polys)), "ja_round_eval: polynomial length mismatch");
  if (has_table || kernel_id == JA_EVAL_DOT2_L2H) {
    if (kernel_id != JA_EVAL_WIDENT) {
      JA_REQUIRE(eq == nullptr, "ja_round_eval: this kernel_id takes no split-eq handle");
      return ja_round_eval_w_launch(c, kernel_id, polys, n_polys, aux_u32, n_out, pend);
    }
  }
  EvalPolys P;
  for (size_t i = 0; i < 6; i++) P.p[i] = i < n_polys ? polys[i]->data() : nullptr;
  P.aux = nullptr; P.shift = 0;
  const size_t G = len / 2;
  size_t want_out = 0, want_polys = 0;
  enum { FAM_S, FAM_D, FAM_PROD, FAM_SUM } fam = FAM_S;
  switch (kernel_id) {
    case JA_EVAL_ADD: case JA_EVAL_SUB: want_out = 1; want_polys = 2; break;
    case JA_EVAL_MUL: want_out = 2; want_polys = 2; break;
    case JA_EVAL_SQUARE: want_out = 2; want_polys = 1; break;
    case JA_EVAL_IDENT: want_out = 1; want_polys = 1; break;
    case JA_EVAL_WIDENT: want_out = 1; want_polys = 2;
      JA_REQUIRE(aux_u32 < 60 && ((len / 2 - 1) >> aux_u32) < polys[1]->len, "ja_round_eval: WIDENT table shorter than (len/2) >> shift"); break;
    case JA_EVAL_IFF: want_out = 2; want_polys = 3; break;
    case JA_EVAL_DIV: want_out = 2; want_polys = 4; break;
    case JA_EVAL_RSQRT: want_out = 2; want_polys = 5; JA_REQUIRE(aux_fr && n_aux == 2, "ja_round_eval: RSQRT takes aux_fr = {gamma, S^3}"); break;
    case JA_EVAL_LIN3: want_out = 1; want_polys = 3; JA_REQUIRE(aux_fr && n_aux == 1, "ja_round_eval: LIN3 takes aux_fr = {tau}"); break;
    case JA_EVAL_PROD: want_out = n_polys; want_polys = n_polys; fam = FAM_PROD; break;
    case JA_EVAL_POW: want_out = aux_u32; want_polys = 1; fam = FAM_PROD; break;
    case JA_EVAL_DOT2: want_out = 2; want_polys = 2; fam = FAM_D; break;
    case JA_EVAL_DOT3: want_out = 3; want_polys = 3; fam = FAM_D; break;
    case JA_EVAL_SUM1: want_out = 1; want_polys = n_polys; fam = FAM_SUM; break;
    case JA_EVAL_SUMHI: want_out = 1; want_polys = 1; fam = FAM_SUM; break;
    default: return fail(JA_ERR_UNSUPPORTED, "ja_round_eval: kernel_id not implemented");
  }
  JA_REQUIRE(n_out == want_out, "ja_round_eval: wrong n_out for kernel_id");
  JA_REQUIRE(n_polys == want_polys, "ja_round_eval: wrong n_polys for kernel_id");
  size_t n_dev = n_out;      // Fr values to copy back from d_out
  int prod_lanes = 0, prod_d = 0;
  if (fam == FAM_S || fam == FAM_PROD) {
    JA_REQUIRE(eq != nullptr, "ja_round_eval: split-eq handle required for family S");
    JA_REQUIRE(eq->order == JA_LOW_TO_HIGH, "ja_round_eval: family S expects a LowToHigh split-eq");
    const size_t cover = size_t(1) << ((eq->out_len - 1) + (eq->in_len - 1));
    if (c->slice_on) JA_REQUIRE(c->slice_g_offset + G <= cover, "ja_round_eval_slice: slice outside the split-eq tables");
    else JA_REQUIRE(cover == G, "ja_round_eval: split-eq tables do not cover len/2 (eq and polys out of lockstep)");
  } else {
    JA_REQUIRE(eq == nullptr, "ja_round_eval: this kernel_id takes no split-eq handle");
  }
  if (fam == FAM_S) {
    switch (kernel_id) {
      case JA_EVAL_ADD: launch_s<0>(c, P, eq, G); break;
      case JA_EVAL_SUB: launch_s<1>(c, P, eq, G); break;
      case JA_EVAL_MUL: launch_s<2>(c, P, eq, G); break;
      case JA_EVAL_SQUARE: launch_s<3>(c, P, eq, G); break;
      case JA_EVAL_IDENT: launch_s<6>(c, P, eq, G); break;
      case JA_EVAL_WIDENT: P.shift = aux_u32; launch_s<12>(c, P, eq, G); break;
      case JA_EVAL_IFF: launch_s<8>(c, P, eq, G); break;
      case JA_EVAL_DIV: launch_s<9>(c, P, eq, G); break;
      case JA_EVAL_RSQRT: case JA_EVAL_LIN3: {
        Fr* d_aux = nullptr;                            // the body's scalars travel through the staging ring; stream-ordered reuse of the buffer
        int32_t ast = dev_alloc(c, n_aux * sizeof(Fr), (void**)&d_aux);
        if (ast) return ast;
        if ((ast = stage_h2d(c, d_aux, aux_fr, n_aux * 32))) { dev_free(c, d_aux); return ast; }
        P.aux = d_aux;
        if (kernel_id == JA_EVAL_RSQRT) launch_s<10>(c, P, eq, G); else launch_s<11>(c, P, eq, G);
        dev_free(c, d_aux);
        break;
      }
    }
  } else if (fam == FAM_PROD) {
    const int d = kernel_id == JA_EVAL_POW ? (int)aux_u32 : (int)n_polys;
    JA_REQUIRE(d >= 2 && d <= kMaxProdPolys, "ja_round_eval: product degree must be in 2..32");
    ProdPolys PP;
    for (int i = 0; i < kMaxProdPolys; i++) PP.p[i] = i < (int)n_polys ? polys[i]->data() : nullptr;
    const int bits_in = eq->in_len - 1;
    const bool same = kernel_id == JA_EVAL_POW;
    if (d <= 16) {
      // warp-transposed kernel: L lanes per pair, one output point per lane
      int L = 2; while (L < d) L <<= 1;
      const size_t gpb = (size_t)kBlock / L;
      size_t ppb = (G + (size_t)kSMs * 4 - 1) / ((size_t)kSMs * 4);
      ppb = (ppb + gpb - 1) / gpb * gpb;
      const unsigned grid = (unsigned)((G + ppb - 1) / ppb);
#define JA_PROD_T(LL)                                                                                                   \
      if (same) JA_LAUNCH(c, KC_ROUND_EVAL_PROD, k_round_eval_prod_t<LL, true><<<grid, kBlock, 0, c->stream>>>(       \
                    PP, d, eq->e_out(), eq->e_in(), bits_in, G, ppb, c->d_partials, c->d_out, c->d_counter, g_off));    \
      else JA_LAUNCH(c, KC_ROUND_EVAL_PROD, k_round_eval_prod_t<LL, false><<<grid, kBlock, 0, c->stream>>>(            \
               PP, d, eq->e_out(), eq->e_in(), bits_in, G, ppb, c->d_partials, c->d_out, c->d_counter, g_off))
      const size_t g_off = c->slice_on ? c->slice_g_offset : 0;
      switch (L) {
        case 2: JA_PROD_T(2); break;
        case 4: JA_PROD_T(4); break;
        case 8: JA_PROD_T(8); break;
        default: JA_PROD_T(16); break;
      }
#undef JA_PROD_T
      prod_lanes = L; prod_d = d; n_dev = (size_t)L;
    } else {
      const unsigned chunks = (unsigned)((d + kProdChunk - 1) / kProdChunk);
      size_t tiles = (G + kBlock - 1) / kBlock;
      size_t gx = tiles < (size_t)kSMs * 2 ? tiles : (size_t)kSMs * 2;
      size_t tpb = (tiles + gx - 1) / gx;
      gx = (tiles + tpb - 1) / tpb;
      dim3 grid(gx, chunks);
      if (same)
        JA_LAUNCH(c, KC_ROUND_EVAL_PROD, k_round_eval_prod<true><<<grid, kBlock, 0, c->stream>>>(PP, d, eq->e_out(), eq->e_in(), bits_in, G, tpb, c->d_partials, c->d_out, c->d_counter));
      else
        JA_LAUNCH(c, KC_ROUND_EVAL_PROD, k_round_eval_prod<false><<<grid, kBlock, 0, c->stream>>>(PP, d, eq->e_out(), eq->e_in(), bits_in, G, tpb, c->d_partials, c->d_out, c->d_counter));
    }
  } else if (fam == FAM_D) {
    unsigned grid = grid_for(G);
    if (grid > (unsigned)kSMs * 4) grid = kSMs * 4;
    if (kernel_id == JA_EVAL_DOT2)
      JA_LAUNCH(c, KC_ROUND_EVAL_DOT, k_round_eval_dot<2><<<grid, kBlock, 0, c->stream>>>(P, G, c->d_partials, c->d_counter, c->d_out));
    else
      JA_LAUNCH(c, KC_ROUND_EVAL_DOT, k_round_eval_dot<3><<<grid, kBlock, 0, c->stream>>>(P, G, c->d_partials, c->d_counter, c->d_out));
  } else {
    SumPolys SP;
    for (int i = 0; i < kMaxProdPolys; i++) SP.p[i] = i < (int)n_polys ? polys[i]->data() : nullptr;
    unsigned gx = grid_for(G);
    if (gx > (unsigned)kSMs * 2) gx = kSMs * 2;
    dim3 grid(gx, (unsigned)n_polys);
    if (kernel_id == JA_EVAL_SUM1) JA_LAUNCH(c, KC_ROUND_SUM, k_round_sum<2><<<grid, kBlock, 0, c->stream>>>(SP, G, c->d_partials, c->d_out, c->d_counter));
    else                           JA_LAUNCH(c, KC_ROUND_SUM, k_round_sum<1><<<grid, kBlock, 0, c->stream>>>(SP, G, c->d_partials, c->d_out, c->d_counter));
    n_dev = n_polys;
    JA_REQUIRE(aux_fr == nullptr || n_aux == n_polys, "ja_round_eval: SUM1 takes one gamma per polynomial");
  }
  JA_CUDA(cudaGetLastError());
  JA_CUDA(cudaMemcpyAsync(c->h_pinned, c->d_out, n_dev * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  pend->n_dev = n_dev; pend->n_out = n_out; pend->n_polys = n_polys; pend->fam_sum = fam == FAM_SUM; pend->aux_fr = aux_fr;
  pend->prod_lanes = prod_lanes; pend->prod_d = prod_d;
  return JA_OK;
}

int32_t ja_round_eval_collect(ja_ctx* c, const RoundEvalPending& pend, uint64_t* out_evals) {
  JA_CUDA(cudaStreamSynchronize(c->stream));
  if (pend.fam_sum) {
    // hamming_weight.rs:126-133: sum_i gamma_i * (sum_j ra_i[2j]); O(d) scalar glue on the d returned sums
    const FrH* sums = reinterpret_cast<const FrH*>(c->h_pinned);
    FrH acc = host::FR_ZERO;
    for (size_t i = 0; i < pend.n_polys; i++)
      acc = host::add(acc, pend.aux_fr ? host::mul(host::from_limbs(pend.aux_fr + 4 * i), sums[i]) : sums[i]);
    memcpy(out_evals, acc.l, 32);
  } else if (pend.prod_lanes) {
    // lane j of the transposed kernel holds grid point j: X = j + 1 for j < L - 1, X = inf at j = L - 1
    memcpy(out_evals, c->h_pinned, (size_t)(pend.prod_d - 1) * sizeof(Fr));
    memcpy(out_evals + 4 * (pend.prod_d - 1), c->h_pinned + 4 * (pend.prod_lanes - 1), sizeof(Fr));
  } else {
    memcpy(out_evals, c->h_pinned, pend.n_out * sizeof(Fr));
  }
  return JA_OK;
}

extern "C" {

// ---- MLE evaluation ----------------------------------------------------------------------------------
int32_t ja_poly_evaluate(ja_ctx* c, const ja_poly* p, const uint64_t* point, size_t m, uint64_t out[4]) {
  JA_REQUIRE(c && p && out && (point || m == 0), "ja_poly_evaluate: null argument");
  JA_REQUIRE((size_t(1) << m) == p->len, "ja_poly_evaluate: point length does not match polynomial");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  // sum_x eq(point, x) * Z[x]  (dense_mlpoly.rs:265-305 computes the same sum with the two half tables)
  ja_poly* eq = nullptr;
  int32_t st = ja_eq_evals(c, point, m, nullptr, &eq);
  if (st) return st;
  unsigned grid = grid_for(p->len);
  if (grid > (unsigned)kSMs * 4) grid = kSMs * 4;
  JA_LAUNCH(c, KC_ROUND_EVAL_DOT, k_dot_full<<<grid, kBlock, 0, c->stream>>>(p->data(), eq->data(), p->len, c->d_partials, c->d_counter, c->d_out));
  JA_CUDA(cudaGetLastError());
  JA_CUDA(cudaMemcpyAsync(c->h_pinned, c->d_out, sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(out, c->h_pinned, sizeof(Fr));
  ja_poly_free(c, eq);
  return JA_OK;
}

// ---- tensor fold ---------------------------------------------------------------------------------------
// core of the einsum operand fold on a device-resident tensor
static int32_t tensor_fold_dev(ja_ctx* c, const int* dA, size_t rows, size_t cols, const ja_poly* eq, int32_t transpose, ja_poly** out) {
  const size_t out_n = transpose ? rows : cols;
  int32_t st = ja_poly_alloc(c, out_n, out);
  if (st) return st;
  if (transpose) {
    size_t threads = rows * 32;
    JA_LAUNCH(c, KC_TENSOR_FOLD, k_fold_rows<<<(unsigned)((threads + kBlock - 1) / kBlock), kBlock, 0, c->stream>>>(dA, rows, cols, eq->data(), (*out)->buf[0]));
  } else {
    // split rows so that the grid has >= ~4 waves worth of threads
    size_t col_blocks = (cols + kBlock - 1) / kBlock;
    size_t slices = 1;
    while (col_blocks * slices < (size_t)kSMs * 4 && slices * 8 <= rows) slices *= 2;
    size_t rps = (rows + slices - 1) / slices;
    Fr* partial = nullptr;
    st = dev_alloc(c, slices * cols * sizeof(Fr), (void**)&partial);
    if (st) return st;
    dim3 grid((unsigned)col_blocks, (unsigned)slices);
    JA_LAUNCH(c, KC_TENSOR_FOLD, k_fold_cols<<<grid, kBlock, 0, c->stream>>>(dA, rows, cols, eq->data(), rps, partial));
    JA_LAUNCH(c, KC_TENSOR_FOLD, k_fold_cols_finish<<<(unsigned)col_blocks, kBlock, 0, c->stream>>>(partial, slices, cols, (*out)->buf[0]));
    dev_free(c, partial);
  }
  JA_CUDA(cudaGetLastError());
  return JA_OK;
}

int32_t ja_tensor_fold_i32(ja_ctx* c, const int32_t* A, size_t rows, size_t cols, const ja_poly* eq,
                           int32_t transpose, ja_poly** out) {
  JA_REQUIRE(c && A && eq && out, "ja_tensor_fold_i32: null argument");
  JA_REQUIRE(rows && cols, "ja_tensor_fold_i32: empty tensor");
  JA_REQUIRE(eq->len >= (transpose ? cols : rows), "ja_tensor_fold_i32: eq table shorter than the folded axis");
  JA_REQUIRE(is_pow2(transpose ? rows : cols), "ja_tensor_fold_i32: output length must be a power of two");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  int* dA = nullptr;
  int32_t st = dev_alloc(c, rows * cols * sizeof(int), (void**)&dA);
  if (st) return st;
  JA_CUDA(cudaMemcpyAsync(dA, A, rows * cols * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  st = tensor_fold_dev(c, dA, rows, cols, eq, transpose, out);
  dev_free(c, dA);
  if (st) return st;
  JA_CUDA(cudaStreamSynchronize(c->stream));      // the host tensor is borrowed for the duration of the call only
  return JA_OK;
}

// Model weights / activations of a node live on the device for the whole proof (they are inputs of several stages):
// upload once, fold any number of times.
int32_t ja_tensor_i32_upload(ja_ctx* c, const int32_t* A, size_t rows, size_t cols, ja_tensor_i32** out) {
  JA_REQUIRE(c && A && out && rows && cols, "ja_tensor_i32_upload: null or empty argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  std::unique_ptr<ja_tensor_i32> t(new ja_tensor_i32());
  t->rows = rows; t->cols = cols;
  int32_t st = dev_alloc(c, rows * cols * sizeof(int), (void**)&t->data);
  if (st) return st;
  JA_CUDA(cudaMemcpyAsync(t->data, A, rows * cols * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));
  *out = t.release();
  return JA_OK;
}
void ja_tensor_i32_free(ja_ctx* c, ja_tensor_i32* t) {
  if (!c || !t) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  dev_free(c, t->data);
  delete t;
}
int32_t ja_tensor_fold_resident(ja_ctx* c, const ja_tensor_i32* t, const ja_poly* eq, int32_t transpose, ja_poly** out) {
  JA_REQUIRE(c && t && eq && out, "ja_tensor_fold_resident: null argument");
  JA_REQUIRE(eq->len >= (transpose ? t->cols : t->rows), "ja_tensor_fold_resident: eq table shorter than the folded axis");
  JA_REQUIRE(is_pow2(transpose ? t->rows : t->cols), "ja_tensor_fold_resident: output length must be a power of two");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  return tensor_fold_dev(c, t->data, t->rows, t->cols, eq, transpose, out);
}

// ---- measurement hooks ------------------------------------------------------------------------------------
int32_t ja_timer_begin(ja_ctx* c) {
  JA_REQUIRE(c, "ja_timer_begin: null ctx");
  JA_CUDA(cudaEventRecord(c->ev0, c->stream));
  return JA_OK;
}
int32_t ja_timer_end(ja_ctx* c, float* out_ms) {
  JA_REQUIRE(c && out_ms, "ja_timer_end: null argument");
  JA_CUDA(cudaEventRecord(c->ev1, c->stream));
  JA_CUDA(cudaEventSynchronize(c->ev1));
  JA_CUDA(cudaEventElapsedTime(out_ms, c->ev0, c->ev1));
  return JA_OK;
}

__global__ void k_fill_pseudo(Fr* out, size_t n, uint32_t seed) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Fr v;   // canonical: top limb < 0x30000000
    uint32_t x = (uint32_t)i * 2654435761u + seed;
#pragma unroll
    for (int k = 0; k < 8; k++) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; v.l[k] = x; }
    v.l[7] &= 0x1fffffffu;
    fp_store(out + i, v);
  }
}

int32_t ja_poly_random(ja_ctx* c, size_t n, uint32_t seed, ja_poly** out) {
  JA_REQUIRE(c && out, "ja_poly_random: null argument");
  int32_t st = ja_poly_alloc(c, n, out);
  if (st) return st;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_LAUNCH(c, KC_MISC, k_fill_pseudo<<<kSMs * 8, kBlock, 0, c->stream>>>((*out)->buf[0], n, seed));
  JA_CUDA(cudaGetLastError());
  return JA_OK;
}

int32_t ja_bench_kernel(ja_ctx* c, int32_t which, int32_t log_n, int32_t n_polys, int32_t iters, float* out_ms) {
  JA_REQUIRE(c && out_ms && iters > 0 && log_n >= 1 && log_n <= 30, "ja_bench_kernel: bad argument");
  JA_REQUIRE(n_polys >= 1 && n_polys <= kMaxBindPolys, "ja_bench_kernel: bad n_polys");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  const size_t n = size_t(1) << log_n, half = n / 2;
  const int np = (which == 0 || which == 1) ? n_polys : 2;
  std::vector<Fr*> src(np, nullptr), dst(np, nullptr);
  int32_t st;
  for (int i = 0; i < np; i++) {
    if ((st = dev_alloc(c, n * sizeof(Fr), (void**)&src[i]))) return st;
    JA_LAUNCH(c, KC_MISC, k_fill_pseudo<<<kSMs * 8, kBlock, 0, c->stream>>>(src[i], n, 17u + i));
    if (which <= 1) { if ((st = dev_alloc(c, half * sizeof(Fr), (void**)&dst[i]))) return st; }
  }
  JA_CUDA(cudaGetLastError());
  const uint64_t rr[4] = {0, 0, 0x0123456789abcdefull, 0x0fedcba987654321ull};
  const Challenge ch = to_challenge(rr);
  ja_spliteq* eq = nullptr;
  if (which == 2 || which == 4) {
    std::vector<uint64_t> w((size_t)log_n * 4);
    for (int i = 0; i < log_n; i++) { w[4 * i] = 0; w[4 * i + 1] = 0; w[4 * i + 2] = 0x9e3779b97f4a7c15ull * (i + 1); w[4 * i + 3] = 0x0123456789abcdefull + i; }
    if ((st = ja_spliteq_new(c, w.data(), (size_t)log_n, JA_LOW_TO_HIGH, nullptr, &eq))) return st;
  }
  auto launch = [&]() {
    if (which <= 1) {
      BindArgs args;
      for (int i = 0; i < np; i++) { args.in[i] = src[i]; args.out[i] = dst[i]; }
      dim3 grid(grid_for(half), (unsigned)np);
      if (which == 0) JA_LAUNCH(c, KC_BIND, k_bind<true><<<grid, kBlock, 0, c->stream>>>(args, ch, half));
      else            JA_LAUNCH(c, KC_BIND, k_bind<false><<<grid, kBlock, 0, c->stream>>>(args, ch, half));
    } else {
      EvalPolys P; P.p[0] = src[0]; P.p[1] = src[1]; P.p[2] = P.p[3] = P.p[4] = P.p[5] = nullptr; P.aux = nullptr; P.shift = 0;
      if (which == 2) launch_s<2>(c, P, eq, half);
      else if (which == 4) launch_s<0>(c, P, eq, half);
      else {
        unsigned grid = grid_for(half);
        if (grid > (unsigned)kSMs * 4) grid = kSMs * 4;
        JA_LAUNCH(c, KC_ROUND_EVAL_DOT, k_round_eval_dot<2><<<grid, kBlock, 0, c->stream>>>(P, half, c->d_partials, c->d_counter, c->d_out));
      }
    }
  };
  for (int i = 0; i < 3; i++) launch();   // warm-up
  JA_CUDA(cudaEventRecord(c->ev0, c->stream));
  for (int i = 0; i < iters; i++) launch();
  JA_CUDA(cudaEventRecord(c->ev1, c->stream));
  JA_CUDA(cudaEventSynchronize(c->ev1));
  JA_CUDA(cudaGetLastError());
  float ms = 0;
  JA_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  *out_ms = ms / iters;
  for (int i = 0; i < np; i++) { dev_free(c, src[i]); dev_free(c, dst[i]); }
  if (eq) ja_spliteq_free(c, eq);
  return JA_OK;
}

// ---- calibration -----------------------------------------------------------------------------------------
int32_t ja_calibrate_fr_mul(ja_ctx* c, int32_t iters, double* out_mul_per_s) {
  JA_REQUIRE(c && out_mul_per_s && iters > 0, "ja_calibrate_fr_mul: bad argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  Fr* d = nullptr;
  int32_t st = dev_alloc(c, kBlock * sizeof(Fr), (void**)&d);
  if (st) return st;
  cudaEvent_t e0, e1;
  JA_CUDA(cudaEventCreate(&e0)); JA_CUDA(cudaEventCreate(&e1));
  const unsigned grid = kSMs * 8;
  JA_LAUNCH(c, KC_MISC, k_calib_mul<FrParams><<<grid, kBlock, 0, c->stream>>>(d, 16));   // warm-up
  JA_CUDA(cudaEventRecord(e0, c->stream));
  JA_LAUNCH(c, KC_MISC, k_calib_mul<FrParams><<<grid, kBlock, 0, c->stream>>>(d, iters));
  JA_CUDA(cudaEventRecord(e1, c->stream));
  JA_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  JA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *out_mul_per_s = (double)grid * kBlock * 4.0 * iters / (ms * 1e-3);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  dev_free(c, d);
  return JA_OK;
}

}  // extern "C"
