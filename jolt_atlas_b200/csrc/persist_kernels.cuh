// Round-resident sumcheck kernels: EVERY remaining round of a sumcheck instance (or of the RaVirtual + Booleanity pair of an
// RA one-hot check) runs inside ONE cooperative launch.  The Fiat-Shamir transcript stays on the host (a Blake2b
// compression costs 5.2 us on one B200 thread against ~0.1 us on a host core: scripts/micro/persist_probe.cu,
// profiles/r2_persist_probe.txt); what disappears from the per-round critical path is the kernel boundary (~2.5 us of
// retire + start even for a pre-enqueued kernel), the launch call on the host and the cold operand loads:
//
//   round i:   [operands of the first pair already in registers]  wait for r_(i-1)  ->  bind + evaluate  ->  sums published
//              with store_tagged into the instance's host-mapped slot (tag0 + i)  ->  prefetch for round i+1
//   host:      spins on the slot, assembles / hashes the round polynomial, posts r_i into entry i of the call's mailbox
//
// Challenges travel host -> device through a per-call array of 16-byte entries in host-mapped memory (one per round, zeroed
// before the launch; bit 29 of word 3 = valid, bit 31 = abort).  Only block 0 polls host memory (PCIe reads do not overlap);
// it republishes every entry in device memory where the other active blocks pick it up from L2.
// Work split: round i runs on W_i = min(gridDim, ceil(G_i / pairs-per-block)) blocks over contiguous slices; W never grows,
// blocks that are no longer needed exit.  While W stays the same and divides G, a block's slice of round i is exactly what it
// wrote in round i-1 (pair g reads elements 4g..4g+3, written by pairs 2g and 2g+1): the data is block-local and the operand
// prefetch ahead of the challenge wait is legal.  When W changes, blocks read what OTHER blocks wrote: the writers' fence
// + atomic of the cross-block tail (grid_sum_ex / prod_tail), the host round trip and the reader's fence after the
// challenge order those accesses, and the body starts without a prefetch.  All polynomial loads are ld.global.cg.
// The bodies are the ones of the per-round kernels (fused_kernels.cuh); results are bit-identical by construction.
#pragma once
#include "fused_kernels.cuh"

namespace ja {

constexpr int kRrMaxRounds = 64;
constexpr int kRrMaxSPolys = 5;
constexpr unsigned long long kRrTimeoutNs = 30000000000ull;

struct RrWaiter {
  const uint4* host;             // entry in host-mapped memory; nullptr: the challenge is already in r
  uint4* relay;                  // the entry's twin in device memory
  int* s_abort;                  // shared flag of the block: an abort word arrived (or the wait timed out)
  JA_DEV bool operator()(Challenge& r) const {
    if (!host) return true;
    __shared__ uint4 s_v;
    if (threadIdx.x == 0) {
      const bool relay_block = blockIdx.x == 0;
      const uint4* src = relay_block ? host : relay;
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      uint4 v;
      for (unsigned int it = 0;; it++) {
        v = ld_volatile_v4(src);
        if (v.w & 0xa0000000u) break;                                   // valid (bit 29) or abort (bit 31)
        if ((it & 1023u) == 1023u) {
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t1 - t0 > kRrTimeoutNs) { v = make_uint4(0u, 0u, 0u, 0x80000000u); break; }
        }
      }
      if (relay_block) st_volatile_v4(relay, v);
      s_v = v;
      if (v.w >> 31) *s_abort = 1;
    }
    __syncthreads();
    const uint4 v = s_v;
    r.c[0] = v.x; r.c[1] = v.y; r.c[2] = v.z; r.c[3] = v.w & 0x1fffffffu;
    return (v.w >> 31) == 0;
  }
};

// LowToHigh split-eq table cursor: the device twin of ja_spliteq_bind's bookkeeping (split_eq_poly.rs:331-372)
struct RrEq {
  const Fr* out_levels; const Fr* in_levels;
  int out_len, in_len, ci, m;
  JA_DEV const Fr* e_out() const { return out_levels + ((size_t(1) << (out_len - 1)) - 1); }
  JA_DEV const Fr* e_in() const { return in_levels + ((size_t(1) << (in_len - 1)) - 1); }
  JA_DEV void bind() {
    ci -= 1;
    if (m / 2 < ci && in_len > 1) in_len--;
    else if (0 < ci && out_len > 1) out_len--;
  }
};

struct RrCommon {
  const uint4* mail_host;        // [rounds] entries, host-mapped (device address)
  uint4* mail_relay;             // [rounds] twins, device memory, zeroed before the launch
  int rounds;                    // rounds this launch runs
  int first_fused;               // the first round binds r0 first (hand-over from the per-round engine)
  Challenge r0;
  unsigned long long n_first;    // length of the arrays the first round reads
};

// Blocks of a round.  A body runs `ppp` pairs per pass of one block; at most `wmax` blocks; `single_max` or fewer pairs run on ONE
// block (no cross-block tail).  The width is KEPT from round to round while it divides the pairs (then every block's slice is
// what it wrote itself: block-local data, prefetch ahead of the challenge wait, no fence) and only shrinks otherwise.
struct RrSplit { unsigned int ppp, wmax, single_max; };
JA_DEV unsigned int rr_width(unsigned long long G, const RrSplit& s, unsigned int w_prev, bool first) {
  if (G <= s.single_max) return 1;
  if (!first && G >= w_prev && G % w_prev == 0) return w_prev;
  unsigned long long cand = (G + s.ppp - 1) / s.ppp;
  if (cand < 1) cand = 1;
  const unsigned int cap = first ? s.wmax : (w_prev < s.wmax ? w_prev : s.wmax);      // blocks that left do not come back
  return cand > cap ? cap : (unsigned int)cand;
}

// ---- family S single instance (ADD / SUB / MUL / SQUARE / IDENT), LowToHigh ---------------------------------------------------------
struct RrSArgs {
  RrCommon c;
  RrSplit split;
  Fr* buf[kRrMaxSPolys][2];      // [polynomial][0 = array the first round reads, 1 = the other ping-pong buffer]; unused entries repeat polynomial 0
  int np;
  const Fr* gammas;              // scalars of the body (RSQRT, LIN3)
  RrEq eq;                       // state at the first round
  Fr* partials; unsigned int* counter;
  Fr* slot_vals; unsigned int tag0;
};

template <int KID>
__global__ void __launch_bounds__(kBlock) k_rr_s(const __grid_constant__ RrSArgs a) {
  __shared__ int s_abort;
  if (threadIdx.x == 0) s_abort = 0;
  __syncthreads();
  RrEq eq = a.eq;
  int cur = 0;                   // which buffer holds the array the round reads
  unsigned long long n = a.c.n_first;
  unsigned int w_prev = 0;
  Challenge r = a.c.r0;
  for (int i = 0; i < a.c.rounds; i++) {
    const bool fused = i > 0 || a.c.first_fused;
    const unsigned long long G = fused ? n / 4 : n / 2;
    const unsigned int W = rr_width(G, a.split, w_prev, i == 0);
    if (blockIdx.x >= W) return;
    const bool local = i > 0 && W == w_prev && G % W == 0;
    RrWaiter waiter{nullptr, nullptr, &s_abort};
    if (i > 0) { waiter.host = a.c.mail_host + (i - 1); waiter.relay = a.c.mail_relay + (i - 1); }
    if (i > 0 && !local) {
      if (!waiter(r)) return;
      __threadfence();
      waiter.host = nullptr;
    }
    FusedPolys P;
#pragma unroll
    for (int q = 0; q < kRrMaxSPolys; q++) { P.in[q] = a.buf[q][cur]; P.out[q] = a.buf[q][fused ? 1 - cur : cur]; }
    const size_t g_begin = (size_t)((unsigned long long)blockIdx.x * G / W), g_end = (size_t)((unsigned long long)(blockIdx.x + 1) * G / W);
    const Publish pub{a.slot_vals, nullptr, a.tag0 + (unsigned int)i};
    __syncthreads();
    if (fused) round_s_body<KID, true, RrWaiter, true>(P, a.np, r, waiter, eq.e_out(), eq.e_in(), eq.in_len - 1, g_begin, g_end, a.gammas, a.partials, a.counter, pub, blockIdx.x, W, 0);
    else round_s_body<KID, false, RrWaiter, true>(P, a.np, r, waiter, eq.e_out(), eq.e_in(), eq.in_len - 1, g_begin, g_end, a.gammas, a.partials, a.counter, pub, blockIdx.x, W, 0);
    __syncthreads();
    if (s_abort) return;
    if (fused) { cur = 1 - cur; n /= 2; }
    eq.bind();
    w_prev = W;
  }
  // final bind: the arrays have 2 entries; block 0 binds the last challenge and leaves the claims in element 0 of the other buffer
  if (blockIdx.x != 0) return;
  RrWaiter waiter{a.c.mail_host + (a.c.rounds - 1), a.c.mail_relay + (a.c.rounds - 1), &s_abort};
  if (!waiter(r)) return;
  __threadfence();
  if (threadIdx.x < a.np) {
    const Fr* z = a.buf[threadIdx.x][cur];
    const Fr a0 = fr_ld<true>(z), a1 = fr_ld<true>(z + 1);
    const Fr fin = fp_add<FrParams>(a0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a1, a0), r));
    fp_store(a.buf[threadIdx.x][1 - cur], fin);
    store_tagged(a.slot_vals, threadIdx.x, fin, a.tag0 + (unsigned int)a.c.rounds);     // the final claims travel like a round's sums
  }
}

// ---- family D single instance (DOT2 / DOT3), HighToLow in place, ONE block (contraction lengths are small) ------------------------------
struct RrDotArgs {
  RrCommon c;
  Fr* buf[3];
  Fr* partials; unsigned int* counter;
  Fr* slot_vals; unsigned int tag0;
};
template <int NP>
__global__ void __launch_bounds__(kBlock) k_rr_dot(const __grid_constant__ RrDotArgs a) {
  __shared__ int s_abort;
  if (threadIdx.x == 0) s_abort = 0;
  __syncthreads();
  unsigned long long n = a.c.n_first;
  Challenge r = a.c.r0;
  FusedPolys P;
#pragma unroll
  for (int q = 0; q < NP; q++) { P.in[q] = a.buf[q]; P.out[q] = a.buf[q]; }
  for (int i = 0; i < a.c.rounds; i++) {
    const bool fused = i > 0 || a.c.first_fused;
    const unsigned long long G = fused ? n / 4 : n / 2;
    RrWaiter waiter{nullptr, nullptr, &s_abort};
    if (i > 0) { waiter.host = a.c.mail_host + (i - 1); waiter.relay = a.c.mail_relay + (i - 1); }
    const Publish pub{a.slot_vals, nullptr, a.tag0 + (unsigned int)i};
    __syncthreads();
    if (fused) round_dot_body<NP, true, RrWaiter, true>(P, r, waiter, (size_t)G, a.partials, a.counter, pub, 0, 1);
    else round_dot_body<NP, false, RrWaiter, true>(P, r, waiter, (size_t)G, a.partials, a.counter, pub, 0, 1);
    __syncthreads();
    if (s_abort) return;
    if (fused) n /= 2;
  }
  RrWaiter waiter{a.c.mail_host + (a.c.rounds - 1), a.c.mail_relay + (a.c.rounds - 1), &s_abort};
  if (!waiter(r)) return;
  if (threadIdx.x < NP) {
    Fr* z = a.buf[threadIdx.x];
    const Fr a0 = fr_ld<true>(z), a1 = fr_ld<true>(z + 1);
    const Fr fin = fp_add<FrParams>(a0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a1, a0), r));
    fp_store(z, fin);
    store_tagged(a.slot_vals, threadIdx.x, fin, a.tag0 + (unsigned int)a.c.rounds);
  }
}

// ---- RA one-hot checks: RaVirtual (product of d <= 16) + Booleanity phase 2 over d polynomials, same length -----------------
// The two instances of a round are independent: they run CONCURRENTLY on two sub-grids (blocks [0, off_b): product,
// [off_b, ...): booleanity), each with its own width schedule, scratch and result slot - as the paired per-round launch does.
struct RrPairArgs {
  RrCommon c;
  int d;
  unsigned int off_b;            // first block of the booleanity sub-grid
  RrSplit split_a, split_wide, split_b;
  Fr* bufA[16][2];               // RaVirtual polynomials
  Fr* bufB[16][2];               // Booleanity H polynomials
  RrEq eqA, eqB;
  const Fr* gammas;              // Booleanity
  Fr* partialsA; unsigned int* counterA; Fr* slotA; unsigned int tagA0;
  Fr* partialsB; unsigned int* counterB; Fr* slotB; unsigned int tagB0;
  unsigned int wide_max_pairs;   // L == 16: rounds with at most this many pairs run the 64-threads-per-pair product
};

template <int L>
__global__ void __launch_bounds__(kBlock) k_rr_pair(const __grid_constant__ RrPairArgs a) {
  __shared__ int s_abort;
  if (threadIdx.x == 0) s_abort = 0;
  __syncthreads();
  const bool role_b = blockIdx.x >= a.off_b;
  const unsigned int bx = role_b ? blockIdx.x - a.off_b : blockIdx.x;
  RrEq eq = role_b ? a.eqB : a.eqA;
  Fr* const (*buf)[2] = role_b ? a.bufB : a.bufA;
  int cur = 0;
  unsigned long long n = a.c.n_first;
  unsigned int w_prev = 0;
  Challenge r = a.c.r0;
  for (int i = 0; i < a.c.rounds; i++) {
    const bool fused = i > 0 || a.c.first_fused;
    const unsigned long long G = fused ? n / 4 : n / 2;
    const bool wide = !role_b && L == 16 && G <= a.wide_max_pairs;
    const unsigned int W = rr_width(G, role_b ? a.split_b : (wide ? a.split_wide : a.split_a), w_prev, i == 0);
    if (bx >= W) return;
    RrWaiter waiter{nullptr, nullptr, &s_abort};
    if (i > 0) {
      waiter.host = a.c.mail_host + (i - 1); waiter.relay = a.c.mail_relay + (i - 1);
      if (!waiter(r)) return;
      if (W != w_prev || G % W != 0) __threadfence();       // this block reads what other blocks wrote
      waiter.host = nullptr;
    }
    FusedPolys P;
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int qq = q < a.d ? q : 0;
      P.in[q] = buf[qq][cur]; P.out[q] = buf[qq][fused ? 1 - cur : cur];
    }
    const size_t g_begin = (size_t)((unsigned long long)bx * G / W), g_end = (size_t)((unsigned long long)(bx + 1) * G / W);
    __syncthreads();
    if (!role_b) {
      const Publish pub{a.slotA, nullptr, a.tagA0 + (unsigned int)i};
      if (fused) {
        if (L == 16 && wide) round_prod16_wide_body<true, kBlock, RrWaiter, true>(P, a.d, r, waiter, eq.e_out(), eq.e_in(), eq.in_len - 1, g_begin, g_end, a.partialsA, a.counterA, pub, bx, W);
        else round_prod_body<L, false, true, kBlock, RrWaiter, true>(P, a.d, r, waiter, eq.e_out(), eq.e_in(), eq.in_len - 1, g_begin, g_end, a.partialsA, a.counterA, pub, bx, W);
      } else {
        if (L == 16 && wide) round_prod16_wide_body<false, kBlock, RrWaiter, true>(P, a.d, r, waiter, eq.e_out(), eq.e_in(), eq.in_len - 1, g_begin, g_end, a.partialsA, a.counterA, pub, bx, W);
        else round_prod_body<L, false, false, kBlock, RrWaiter, true>(P, a.d, r, waiter, eq.e_out(), eq.e_in(), eq.in_len - 1, g_begin, g_end, a.partialsA, a.counterA, pub, bx, W);
      }
    } else {
      const Publish pub{a.slotB, nullptr, a.tagB0 + (unsigned int)i};
      if (fused) round_bool_body<L, true, kBlock, RrWaiter, true>(P, a.d, r, waiter, eq.e_out(), eq.e_in(), eq.in_len - 1, g_begin, g_end, a.gammas, a.partialsB, a.counterB, pub, bx, W);
      else round_bool_body<L, false, kBlock, RrWaiter, true>(P, a.d, r, waiter, eq.e_out(), eq.e_in(), eq.in_len - 1, g_begin, g_end, a.gammas, a.partialsB, a.counterB, pub, bx, W);
    }
    __syncthreads();
    if (fused) { cur = 1 - cur; n /= 2; }
    eq.bind();
    w_prev = W;
  }
  // final bind by block 0 of each sub-grid
  if (bx != 0) return;
  RrWaiter waiter{a.c.mail_host + (a.c.rounds - 1), a.c.mail_relay + (a.c.rounds - 1), &s_abort};
  if (!waiter(r)) return;
  __threadfence();
  if (threadIdx.x < a.d) {
    Fr* const* b = buf[threadIdx.x];
    const Fr a0 = fr_ld<true>(b[cur]), a1 = fr_ld<true>(b[cur] + 1);
    const Fr fin = fp_add<FrParams>(a0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a1, a0), r));
    fp_store(b[1 - cur], fin);
    store_tagged(role_b ? a.slotB : a.slotA, threadIdx.x, fin, (role_b ? a.tagB0 : a.tagA0) + (unsigned int)a.c.rounds);
  }
}

}  // namespace ja
