// C ABI of HyperKZG::open (joltworks/src/poly/commitment/hyperkzg/mod.rs:400-447, :231-280, :192-229) split at the
// two Fiat–Shamir interaction points, plus an all-in-one entry that runs the library's own Blake2b transcript.
// Everything that touches 2^l-sized data runs on the device; only commitments (l-1 + 3 points) and the 3*l
// evaluations cross to the host for the transcript.
#include "common.hpp"
#include "hyperkzg_kernels.cuh"
#include "transcript_host.hpp"

struct ja_hkzg {
  const ja_srs* srs = nullptr;
  size_t n = 0;
  int ell = 0;
  Fr* P = nullptr;   // [poly_0 (n) | poly_1 (n/2) | ... | poly_{ell-1} (2)], 2n - 2 coefficients
};

static inline size_t poly_off(size_t n, int k) { return 2 * n - (2 * n >> k); }   // sum_{i<k} n >> i

static void build_pow_tab(const FrH& u, FrH* tab) {
  tab[0] = u;
  for (int j = 1; j < kPowTab; j++) tab[j] = host::sqr(tab[j - 1]);
}

extern "C" {

void ja_hyperkzg_open_free(ja_ctx* c, ja_hkzg* h) {
  if (!c || !h) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  dev_free(c, h->P);
  delete h;
}

int32_t ja_hyperkzg_open_begin(ja_ctx* c, const ja_srs* srs, const ja_poly* poly, const uint64_t* point, size_t ell,
                               ja_hkzg** out, uint64_t* com_xy, int32_t* com_inf) {
  JA_REQUIRE(c && srs && poly && point && out, "ja_hyperkzg_open_begin: null argument");
  JA_REQUIRE(ell >= 1 && ell < 31, "ja_hyperkzg_open_begin: bad number of variables");
  JA_REQUIRE(poly->len == (size_t(1) << ell), "ja_hyperkzg_open_begin: polynomial length must be 2^ell (mod.rs:408)");
  JA_REQUIRE(ell == 1 || com_xy, "ja_hyperkzg_open_begin: null output");
  if (poly->len > srs->n)
    return fail(JA_ERR_KEY_LENGTH, "KeyLengthError: SRS has " + std::to_string(srs->n) + " powers, polynomial needs " +
                                       std::to_string(poly->len));
  for (size_t i = 0; i < ell; i++)
    JA_REQUIRE(point[4 * i] == 0 && point[4 * i + 1] == 0, "ja_hyperkzg_open_begin: point must hold MontU128Challenge limbs {0,0,lo,hi}");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  ja_hkzg* h = new ja_hkzg();
  h->srs = srs; h->n = poly->len; h->ell = (int)ell;
  const size_t n = h->n;
  int32_t st = dev_alloc(c, 2 * n * sizeof(Fr), (void**)&h->P);
  if (st) { delete h; return st; }
  JA_CUDA(cudaMemcpyAsync(h->P, poly->data(), n * sizeof(Fr), cudaMemcpyDeviceToDevice, c->stream));
  // Phase 1 (mod.rs:413-428): Pi[j] = point[ell-i-1] * (prev[2j+1] - prev[2j]) + prev[2j]
  for (size_t i = 0; i + 1 < ell; i++) {
    const size_t half = n >> (i + 1);
    BindArgs args;
    args.in[0] = h->P + poly_off(n, (int)i);
    args.out[0] = h->P + poly_off(n, (int)i + 1);
    const Challenge ch = to_challenge(point + 4 * (ell - i - 1));
    JA_LAUNCH(c, KC_BIND, k_bind<true><<<dim3(grid_for(half), 1), kBlock, 0, c->stream>>>(args, ch, half));
  }
  JA_CUDA(cudaGetLastError());
  // commit_variable_batch(polys[1..]) (kzg.rs:227-243): one bucket pipeline for the l-1 folded polynomials
  if (ell > 1) {
    std::vector<MsmJob> jobs;
    for (size_t k = 1; k < ell; k++) jobs.push_back(MsmJob{h->P + poly_off(n, (int)k), n >> k, 0, 254, 0, 1});
    st = ja_msm_run(c, srs, jobs, com_xy, com_inf);
    if (st) { ja_hyperkzg_open_free(c, h); return st; }
  }
  *out = h;
  return JA_OK;
}

int32_t ja_hyperkzg_open_evals(ja_ctx* c, ja_hkzg* h, const uint64_t r[4], uint64_t* v_out) {
  JA_REQUIRE(c && h && r && v_out, "ja_hyperkzg_open_evals: null argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  const FrH rr = host::from_limbs(r);
  const FrH u[3] = {rr, host::neg(rr), host::sqr(rr)};   // mod.rs:441
  FrH* tab_h = reinterpret_cast<FrH*>(c->h_pinned);
  for (int p = 0; p < 3; p++) build_pow_tab(u[p], tab_h + p * kPowTab);
  Fr* d_tab = nullptr; Fr* d_v = nullptr;
  int32_t st = dev_alloc(c, 3 * kPowTab * sizeof(Fr), (void**)&d_tab);
  if (st) return st;
  st = dev_alloc(c, 3 * (size_t)h->ell * sizeof(Fr), (void**)&d_v);
  if (st) return st;
  JA_CUDA(cudaMemcpyAsync(d_tab, tab_h, 3 * kPowTab * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  for (int k = 0; k < h->ell; k++) {
    const size_t len = h->n >> k;
    int log_s = 8;
    while (log_s < 18 && (size_t(8) << log_s) < len) log_s++;
    const unsigned grid = 1u << (log_s - 8);
    JA_LAUNCH(c, KC_HKZG_EVAL, k_univariate_eval3<<<grid, kBlock, 0, c->stream>>>(h->P + poly_off(h->n, k), len, d_tab, log_s, c->d_partials,
                                                      c->d_counter, d_v + 3 * k));
  }
  JA_CUDA(cudaGetLastError());
  std::vector<FrH> tmp(3 * (size_t)h->ell);
  JA_CUDA(cudaMemcpyAsync(tmp.data(), d_v, tmp.size() * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  JA_CUDA(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < h->ell; k++)
    for (int p = 0; p < 3; p++) memcpy(v_out + 4 * ((size_t)p * h->ell + k), tmp[3 * k + p].l, 32);   // v[point][poly]
  dev_free(c, d_tab); dev_free(c, d_v);
  return JA_OK;
}

int32_t ja_hyperkzg_open_witness(ja_ctx* c, ja_hkzg* h, const uint64_t r[4], const uint64_t* q_powers, uint64_t* w_xy,
                                 int32_t* w_inf) {
  JA_REQUIRE(c && h && r && q_powers && w_xy, "ja_hyperkzg_open_witness: null argument");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  const size_t n = h->n;
  const int ell = h->ell;
  const FrH rr = host::from_limbs(r);
  const FrH u[3] = {rr, host::neg(rr), host::sqr(rr)};
  // B = sum_k q^k polys[k]  (mod.rs:262-270)
  Fr *d_q = nullptr, *d_B = nullptr, *d_H = nullptr, *d_tab = nullptr, *d_F = nullptr, *d_C = nullptr;
  int32_t st;
  if ((st = dev_alloc(c, (size_t)ell * sizeof(Fr), (void**)&d_q))) return st;
  if ((st = dev_alloc(c, n * sizeof(Fr), (void**)&d_B))) return st;
  if ((st = dev_alloc(c, 3 * n * sizeof(Fr), (void**)&d_H))) return st;
  if ((st = dev_alloc(c, 3 * kPowTab * sizeof(Fr), (void**)&d_tab))) return st;
  JA_REQUIRE((size_t)ell * 32 + 3 * kPowTab * 32 <= kPinnedBytes, "ja_hyperkzg_open_witness: staging overflow");
  FrH* stage = reinterpret_cast<FrH*>(c->h_pinned);
  memcpy(stage, q_powers, (size_t)ell * 32);
  for (int p = 0; p < 3; p++) build_pow_tab(u[p], stage + ell + p * kPowTab);
  JA_CUDA(cudaMemcpyAsync(d_q, stage, (size_t)ell * 32, cudaMemcpyHostToDevice, c->stream));
  JA_CUDA(cudaMemcpyAsync(d_tab, stage + ell, 3 * kPowTab * 32, cudaMemcpyHostToDevice, c->stream));
  unsigned grid = grid_for(n);
  if (grid > (unsigned)kSMs * 8) grid = kSMs * 8;
  JA_LAUNCH(c, KC_HKZG_LINCOMB, k_hkzg_lincomb<<<grid, kBlock, 0, c->stream>>>(h->P, n, ell, d_q, d_B));
  // h_i = witness polynomial of B at u_i (mod.rs:213-229), i = 0..2
  int log_l = 5;
  while ((size_t(1) << log_l) > n) log_l--;
  const size_t nchunks = n >> log_l;
  const size_t nb = (nchunks + kWitBlock - 1) / kWitBlock;
  int log_per = 0;
  while ((size_t(1024) << log_per) < nb) log_per++;
  if ((st = dev_alloc(c, nb * sizeof(Fr), (void**)&d_F))) return st;
  if ((st = dev_alloc(c, (nb + 1) * sizeof(Fr), (void**)&d_C))) return st;
  for (int p = 0; p < 3; p++) {
    const Fr* tab = d_tab + p * kPowTab;
    JA_LAUNCH(c, KC_HKZG_WITNESS, k_witness_block_sums<<<(unsigned)nb, kWitBlock, 0, c->stream>>>(d_B, nchunks, log_l, tab, d_F));
    JA_LAUNCH(c, KC_HKZG_WITNESS, k_witness_carry<<<1, 1024, 0, c->stream>>>(d_F, nb, log_per, log_l + 8, tab, d_C));
    JA_LAUNCH(c, KC_HKZG_WITNESS, k_witness_write<<<(unsigned)nb, kWitBlock, 0, c->stream>>>(d_B, nchunks, log_l, tab, d_C, d_H + (size_t)p * n));
  }
  JA_CUDA(cudaGetLastError());
  // commit_batch(h) (kzg.rs:195-223)
  std::vector<MsmJob> jobs;
  for (int p = 0; p < 3; p++) jobs.push_back(MsmJob{d_H + (size_t)p * n, n, 0, 254, 0, 1});
  st = ja_msm_run(c, h->srs, jobs, w_xy, w_inf);
  dev_free(c, d_q); dev_free(c, d_B); dev_free(c, d_H); dev_free(c, d_tab); dev_free(c, d_F); dev_free(c, d_C);
  return st;
}

// HyperKZG::open with the library's Blake2b transcript (state and round counter are read and written back), for
// callers that do not own a transcript object of their own (C++ host driver, bench, tests).
int32_t ja_hyperkzg_open(ja_ctx* c, const ja_srs* srs, const ja_poly* poly, const uint64_t* point, size_t ell,
                         uint8_t transcript_state[32], uint32_t* n_rounds, uint64_t* com_xy, int32_t* com_inf,
                         uint64_t* w_xy, int32_t* w_inf, uint64_t* v_out) {
  JA_REQUIRE(transcript_state && n_rounds && com_inf && w_inf && v_out, "ja_hyperkzg_open: null argument");
  ja_hkzg* h = nullptr;
  int32_t st = ja_hyperkzg_open_begin(c, srs, poly, point, ell, &h, com_xy, com_inf);
  if (st) return st;
  host::Blake2bTranscript t(transcript_state, *n_rounds);
  t.append_points(com_xy, com_inf, ell - 1);                       // mod.rs:439
  const FrH r = t.challenge_scalar();                              // mod.rs:440
  st = ja_hyperkzg_open_evals(c, h, r.l, v_out);
  if (st) { ja_hyperkzg_open_free(c, h); return st; }
  t.append_scalars(reinterpret_cast<const FrH*>(v_out), 3 * ell);  // mod.rs:258-259 (v flattened point-major)
  std::vector<FrH> q = t.challenge_scalar_powers(ell);             // mod.rs:260
  st = ja_hyperkzg_open_witness(c, h, r.l, reinterpret_cast<const uint64_t*>(q.data()), w_xy, w_inf);
  ja_hyperkzg_open_free(c, h);
  if (st) return st;
  t.append_points(w_xy, w_inf, 3);                                 // mod.rs:276
  (void)t.challenge_scalar();                                      // mod.rs:277
  memcpy(transcript_state, t.state, 32);
  *n_rounds = t.n_rounds;
  return JA_OK;
}

}  // extern "C"
