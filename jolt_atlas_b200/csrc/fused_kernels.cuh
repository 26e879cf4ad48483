// Round kernels of the sumcheck engine (sumcheck.cu): "bind the previous challenge, then evaluate the next round" in ONE
// pass, and publication of the reduced sums straight into host-mapped pinned memory.
//
// A sumcheck round on the reference is compute_message (reduce over the hypercube) followed, after the transcript, by
// ingest_challenge (bind every MLE).  Launched separately that is two passes over each polynomial and two launches per
// round; fused, round j+1's kernel reads the length-n arrays once, writes the bound length-n/2 arrays and emits the
// sums over them: 48 n bytes per polynomial and round, the algorithmic minimum, and one launch.
// Publication: the finishing block writes the sums to mapped host memory, fences at system scope and then stores a
// sequence number the host spins on - no cudaMemcpyAsync, no cudaStreamSynchronize on the per-round critical path.
//   LowToHigh (family S, product-of-d):  pair g of the bound array comes from in[4g .. 4g+3]
//   HighToLow (family D):                pair (i, i+G) of the bound array comes from in[i], in[i+2G], in[i+G], in[i+3G] (in place)
#pragma once
#include "poly_kernels.cuh"

namespace ja {

struct Publish {
  Fr* vals;                      // device address of the mapped host slot (kMaxOut elements, 48 bytes each when tagged)
  volatile unsigned int* seq;    // device address of the slot's sequence word (fence + flag protocol: k_round_open only)
  unsigned int value;            // the round's tag / sequence value (never 0)
};
// Two publication protocols.  Tagged (every round kernel but k_round_open): store_tagged (poly_kernels.cuh), no fence.
// Fence + flag (k_round_open, k_round_open_rows): plain stores, __threadfence_system, then the sequence word.
// publish_flag: call from every thread of the block that wrote pub.vals; only warp 0 writes the sums.
JA_DEV void publish_flag(const Publish& pub) {
  if (threadIdx.x < 32) {
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) *pub.seq = pub.value;
  }
}

// Final MLE claims of a batch: up to 32 single field elements scattered over device buffers -> pinned host memory
// (device-addressable under unified addressing) in one launch instead of one 32-byte cudaMemcpyAsync each.
struct CollectArgs {
  const Fr* src[32];
  Fr* dst[32];
};
static __global__ void __launch_bounds__(64) k_collect_finals(CollectArgs a, int n) {
  const int i = threadIdx.x >> 1, h = threadIdx.x & 1;
  if (i < n) reinterpret_cast<uint4*>(a.dst[i])[h] = __ldg(reinterpret_cast<const uint4*>(a.src[i]) + h);
}

// ---- challenge mailbox --------------------------------------------------------------------------------------------------
// The kernel of round j+1 is enqueued BEFORE the host has derived r_j (right behind round j's kernel), so that the
// launch overhead and latency are off the Fiat-Shamir critical path: it starts as soon as round j's kernel retires and
// waits on the device for the challenge.  A mailbox entry is ONE 16-byte vector in host-mapped memory: the 125-bit
// challenge in words 0..3 and a 3-bit tag in the free top bits of word 3 (bits 29-30: use counter of the entry, cycling
// 1,2,3 so that consecutive uses differ and zeroed memory never matches; bit 31: abort, the host hit an error).  One
// aligned 16-byte load both polls and delivers the challenge: a read of host memory is a PCIe round trip of ~1 us and
// they do NOT overlap (measured: 16 blocks reading the entry 17 us, 128 blocks 117 us; five 4-byte reads 5 us), so only
// thread 0 of block (0, 0) polls the host entry; it republishes the vector in device memory, where the other blocks pick
// it up from L2.  p == nullptr: no mailbox, the challenge is the kernel parameter.
struct MailRef {
  const uint4* p = nullptr;      // device address of the host-mapped entry
  uint4* dev = nullptr;          // the entry's twin in device memory (relay target)
  uint32_t tag = 0;              // 1..3
  unsigned int* ticket = nullptr;// relay election word of the entry (device memory)
  uint32_t use = 0;              // unique id of this use of the entry (the context's mailbox sequence number, never 0)
};
JA_DEV uint4 ld_volatile_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
JA_DEV void st_volatile_v4(uint4* p, const uint4& v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// Relay election: the FIRST block that reaches the mailbox (atomic exchange of the entry's ticket word with this use's id)
// polls the host entry and republishes it in device memory; every later block reads the twin.  No block waits on a block
// that may not be resident yet (CUDA gives no ordering between the blocks of a grid), so a grid larger than the resident
// capacity, MPS time-slicing or a debugger cannot wedge it.  Give-up path: after kMailTimeoutNs (longer than the host's own
// 20 s wait) the waiting block treats the round as aborted and returns without publishing - no trap, the context survives
// and the host reports the missing publication as an error.
constexpr unsigned long long kMailTimeoutNs = 30000000000ull;
JA_DEV bool mail_wait(Challenge& r, const MailRef& m) {
  if (!m.p) return true;
  __shared__ uint4 s_mail;
  if (threadIdx.x == 0) {
    const bool relay = atomicExch(m.ticket, m.use) != m.use;
    const uint4* src = relay ? m.p : m.dev;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint4 v;
    for (unsigned int it = 0;; it++) {
      v = ld_volatile_v4(src);
      if (((v.w >> 29) & 3u) == m.tag) break;
      if ((it & 255u) == 255u) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > kMailTimeoutNs) { v = make_uint4(0u, 0u, 0u, (m.tag << 29) | 0x80000000u); break; }   // the host is gone: abort
      }
    }
    if (relay) st_volatile_v4(m.dev, v);
    s_mail = v;
  }
  __syncthreads();
  const uint4 v = s_mail;
  r.c[0] = v.x; r.c[1] = v.y; r.c[2] = v.z; r.c[3] = v.w & 0x1fffffffu;
  return (v.w >> 31) == 0;
}

// How a round body obtains the challenge it binds first.  MailWaiter: the pre-launched one-round kernels (mailbox above).
// The round-resident kernels (persist_kernels.cuh) pass their own waiter; NoWait: the challenge is already in `r`.
struct MailWaiter {
  MailRef m;
  JA_DEV bool operator()(Challenge& r) const { return mail_wait(r, m); }
};
struct NoWait {
  JA_DEV bool operator()(Challenge&) const { return true; }
};
// Polynomial loads of a round body.  CG = false: ld.global.nc (the arrays are read-only for the one-round kernels);
// CG = true: ld.global.cg - the round-resident kernels read arrays that earlier rounds of the SAME kernel wrote (possibly
// from another SM), which .nc must not be used for and a stale L1 line must not serve.
template <bool CG> JA_DEV Fr fr_ld(const Fr* p) {
  if (CG) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 lo = __ldcg(q), hi = __ldcg(q + 1);
    Fr r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w; r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
  }
  return fp_load(p);
}

// The same collection with tagged publication (store_tagged) into the context's mapped value buffer: the host picks the claims
// up as soon as they land, without a stream synchronisation at the end of every sumcheck call.
struct CollectIdxArgs {
  const Fr* src[32];
  unsigned int idx[32];
};
static __global__ void __launch_bounds__(32) k_collect_finals_tagged(CollectIdxArgs a, int n, Fr* mapped, unsigned int tag) {
  const int i = threadIdx.x;
  if (i < n) store_tagged(mapped, (int)a.idx[i], fp_load(a.src[i]), tag);
}

struct FusedPolys {
  const Fr* in[kMaxProdPolys];
  Fr* out[kMaxProdPolys];        // FUSED only: bound arrays (LowToHigh: other ping-pong buffer; HighToLow: == in)
};

// (lo, hi) = elements (2g, 2g+1) of the array the round evaluates
template <bool FUSED, bool CG = false>
JA_DEV void load_pair_l2h(const Fr* __restrict__ in, Fr* __restrict__ out, size_t g, const Challenge& r, Fr& lo, Fr& hi) {
  if (FUSED) {
    const Fr a0 = fr_ld<CG>(in + 4 * g), a1 = fr_ld<CG>(in + 4 * g + 1), a2 = fr_ld<CG>(in + 4 * g + 2), a3 = fr_ld<CG>(in + 4 * g + 3);
    lo = fp_add<FrParams>(a0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a1, a0), r));
    hi = fp_add<FrParams>(a2, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a3, a2), r));
    fp_store(out + 2 * g, lo);
    fp_store(out + 2 * g + 1, hi);
  } else {
    lo = fr_ld<CG>(in + 2 * g);
    hi = fr_ld<CG>(in + 2 * g + 1);
  }
}

// ---- family S (split-eq weighted, LowToHigh) ----------------------------------------------------------------------
// KID: 0 ADD, 1 SUB, 2 MUL, 3 SQUARE, 6 IDENT (ids of include/jolt_atlas_b200.h), 7 BOOLEANITY phase 2
//      (booleanity.rs:254-301: [sum_i gamma_i h_i0 (h_i0 - 1), sum_i gamma_i (dh_i)^2] over n_polys one-hot chunks).
//      8 IFF, 9 DIV, 10 RSQRT, 11 LIN3: the three-to-five-operand bodies of poly_kernels.cuh (SGen), operands bound polynomial by polynomial
template <int KID> struct SOut { static constexpr int N = (KID == 2 || KID == 3 || KID == 7 || KID == 8 || KID == 9 || KID == 10) ? 2 : 1; };
template <int KID> struct SPolys { static constexpr int N = KID >= 7 ? 0 : ((KID == 3 || KID == 6) ? 1 : 2); };   // register-staged operand polynomials
template <int KID> struct SGenPolys { static constexpr int N = 1; };
template <> struct SGenPolys<8> { static constexpr int N = 3; };
template <> struct SGenPolys<9> { static constexpr int N = 4; };
template <> struct SGenPolys<10> { static constexpr int N = 5; };
template <> struct SGenPolys<11> { static constexpr int N = 3; };

// The body works on the pairs [g_begin, g_end) of block bx of nb; k_round_s (one round per launch) and the round-resident
// kernels (persist_kernels.cuh: every round of a sumcheck in one launch) both run it.
template <int KID, bool FUSED, class WAITER, bool CG>
JA_DEV void round_s_body(const FusedPolys& P, int n_polys, Challenge r, const WAITER& waiter, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in,
                         int bits_in, size_t g_begin, size_t g_end, const Fr* __restrict__ gammas, Fr* partials, unsigned int* counter,
                         const Publish& pub, unsigned int bx, unsigned int nb, size_t g_off) {
  constexpr int NOUT = SOut<KID>::N;
  constexpr int NP = SPolys<KID>::N;      // polynomials with register-staged operands
  const size_t mask_in = (size_t(1) << bits_in) - 1;
  // The operands of the thread's first pair (and its eq weight) do not depend on the challenge: a pre-launched kernel
  // issues these loads BEFORE it waits on the mailbox, so their latency is spent while the host is still hashing.
  Fr a[NP > 0 ? NP : 1][4];
  Fr ei = fp_zero<FrParams>();
  {
    const size_t g = g_begin + threadIdx.x;
    if (g < g_end) {
      ei = fp_load(e_in + ((g + g_off) & mask_in));
#pragma unroll
      for (int q = 0; q < NP; q++) {
        if (FUSED) {
          const Fr* __restrict__ z = P.in[q] + 4 * g;
          a[q][0] = fr_ld<CG>(z); a[q][1] = fr_ld<CG>(z + 1); a[q][2] = fr_ld<CG>(z + 2); a[q][3] = fr_ld<CG>(z + 3);
        } else {
          a[q][0] = fr_ld<CG>(P.in[q] + 2 * g); a[q][1] = fr_ld<CG>(P.in[q] + 2 * g + 1);
        }
      }
    }
  }
  if (FUSED && !waiter(r)) return;
  Fr outer[NOUT], inner[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; k++) { outer[k] = fp_zero<FrParams>(); inner[k] = fp_zero<FrParams>(); }
  size_t cur_xout = ~size_t(0);
  for (size_t g = g_begin + threadIdx.x; g < g_end; g += kBlock) {
    const size_t x_out = (g + g_off) >> bits_in;
    if (x_out != cur_xout) {
      if (cur_xout != ~size_t(0)) {
        const Fr eo = fp_load(e_out + cur_xout);
#pragma unroll
        for (int k = 0; k < NOUT; k++) {
          outer[k] = fp_add<FrParams>(outer[k], fp_mul<FrParams>(eo, inner[k]));
          inner[k] = fp_zero<FrParams>();
        }
      }
      cur_xout = x_out;
    }
    if (g != g_begin + threadIdx.x) {
      // every global load of the pair (both polynomials) is issued before the first store: the bound arrays may alias
      // nothing the compiler can prove, and a load -> bind -> store -> load chain costs ~2 us per polynomial on a small slab
      ei = fp_load(e_in + ((g + g_off) & mask_in));
#pragma unroll
      for (int q = 0; q < NP; q++) {
        if (FUSED) {
          const Fr* __restrict__ z = P.in[q] + 4 * g;
          a[q][0] = fr_ld<CG>(z); a[q][1] = fr_ld<CG>(z + 1); a[q][2] = fr_ld<CG>(z + 2); a[q][3] = fr_ld<CG>(z + 3);
        } else {
          a[q][0] = fr_ld<CG>(P.in[q] + 2 * g); a[q][1] = fr_ld<CG>(P.in[q] + 2 * g + 1);
        }
      }
    }
    Fr v[NOUT];
    if constexpr (KID >= 8) {
      constexpr int NG = SGenPolys<KID>::N;
      Fr lo[NG], hi[NG];
#pragma unroll
      for (int q = 0; q < NG; q++) load_pair_l2h<FUSED, CG>(P.in[q], P.out[q], g, r, lo[q], hi[q]);
      SGen<KID>::eval(lo, hi, gammas, v);
    } else if (KID == 7) {
      v[0] = fp_zero<FrParams>(); v[1] = fp_zero<FrParams>();
      for (int q = 0; q < n_polys; q++) {
        Fr h0, h1;
        load_pair_l2h<FUSED, CG>(P.in[q], P.out[q], g, r, h0, h1);
        const Fr gm = fp_load(gammas + q);
        const Fr b = fp_sub<FrParams>(h1, h0);
        v[0] = fp_add<FrParams>(v[0], fp_mul<FrParams>(fp_mul<FrParams>(gm, h0), fp_sub<FrParams>(h0, fp_one<FrParams>())));
        v[1] = fp_add<FrParams>(v[1], fp_mul<FrParams>(fp_mul<FrParams>(gm, b), b));
      }
    } else {
      Fr lo[NP > 0 ? NP : 1], hi[NP > 0 ? NP : 1];
#pragma unroll
      for (int q = 0; q < NP; q++) {
        if (FUSED) {
          lo[q] = fp_add<FrParams>(a[q][0], fp_mul_challenge<FrParams>(fp_sub<FrParams>(a[q][1], a[q][0]), r));
          hi[q] = fp_add<FrParams>(a[q][2], fp_mul_challenge<FrParams>(fp_sub<FrParams>(a[q][3], a[q][2]), r));
          fp_store(P.out[q] + 2 * g, lo[q]); fp_store(P.out[q] + 2 * g + 1, hi[q]);
        } else {
          lo[q] = a[q][0]; hi[q] = a[q][1];
        }
      }
      if (KID == 6) v[0] = lo[0];
      else if (KID == 3) { const Fr d = fp_sub<FrParams>(hi[0], lo[0]); v[0] = fp_sqr<FrParams>(lo[0]); v[NOUT - 1] = fp_sqr<FrParams>(d); }
      else if (KID == 0) v[0] = fp_add<FrParams>(lo[0], lo[NP - 1]);
      else if (KID == 1) v[0] = fp_sub<FrParams>(lo[0], lo[NP - 1]);
      else { v[0] = fp_mul<FrParams>(lo[0], lo[NP - 1]); v[NOUT - 1] = fp_mul<FrParams>(fp_sub<FrParams>(hi[0], lo[0]), fp_sub<FrParams>(hi[NP - 1], lo[NP - 1])); }
    }
#pragma unroll
    for (int k = 0; k < NOUT; k++) inner[k] = fp_add<FrParams>(inner[k], fp_mul<FrParams>(ei, v[k]));
  }
  if (cur_xout != ~size_t(0)) {
    const Fr eo = fp_load(e_out + cur_xout);
#pragma unroll
    for (int k = 0; k < NOUT; k++) outer[k] = fp_add<FrParams>(outer[k], fp_mul<FrParams>(eo, inner[k]));
  }
  grid_sum_ex<NOUT>(outer, partials, counter, pub.vals, bx, nb, pub.value);
}


template <int KID, bool FUSED>
__global__ void __launch_bounds__(kBlock, 2)
k_round_s(FusedPolys P, int n_polys, Challenge r, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in, int bits_in,
          size_t G, size_t tiles_per_block, const Fr* __restrict__ gammas, Fr* partials, unsigned int* counter, Publish pub,
          size_t g_off = 0 /* first global pair of this GPU's hypercube slice (multi-GPU); eq tables are indexed globally */,
          MailRef mail = MailRef{}) {
  const size_t g_begin = (size_t)blockIdx.x * tiles_per_block * kBlock;
  size_t g_end = g_begin + tiles_per_block * kBlock;
  if (g_end > G) g_end = G;
  round_s_body<KID, FUSED, MailWaiter, false>(P, n_polys, r, MailWaiter{mail}, e_out, e_in, bits_in, g_begin, g_end, gammas, partials, counter, pub,
                                              blockIdx.x, gridDim.x, g_off);
}

// ---- product of d <= 16 linear factors, warp-transposed (see k_round_eval_prod_t), with the fused bind -------------
// Operand loads of a thread's FIRST pair are issued before the mailbox wait (they do not depend on the challenge): on a
// pre-launched kernel the L2 / HBM latency is then spent while the host is still hashing.
// BLOCK = 128 on small slabs (one warp per scheduler: a dependent chain of Montgomery products issues at one IMAD.WIDE
// per 4 cycles per scheduler, so two resident warps double the latency of the chain).

// cross-block tail shared by the product bodies: L per-block sums -> partials -> last block -> tagged publication
template <int L, int BLOCK>
JA_DEV void prod_tail(Fr tot /* valid in threads < L */, Fr* partials, unsigned int* counter, const Publish& pub, unsigned int bx, unsigned int nb) {
  constexpr int GPB = BLOCK / L;
  const int li = threadIdx.x & (L - 1);
  if (nb == 1) {
    if (threadIdx.x < L) store_tagged(pub.vals, threadIdx.x, tot, pub.value);
    return;
  }
  if (threadIdx.x < L) fp_store(partials + (size_t)bx * L + threadIdx.x, tot);
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicInc(counter, nb - 1) == nb - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  Fr acc = fp_zero<FrParams>();
  for (unsigned b = threadIdx.x / L; b < nb; b += GPB) {
    const uint4* q = reinterpret_cast<const uint4*>(partials + (size_t)b * L + li);
    const uint4 lo = __ldcg(q), hi = __ldcg(q + 1);               // L2 (coherent after the fence), two 16-byte loads
    Fr t;
    t.l[0] = lo.x; t.l[1] = lo.y; t.l[2] = lo.z; t.l[3] = lo.w; t.l[4] = hi.x; t.l[5] = hi.y; t.l[6] = hi.z; t.l[7] = hi.w;
    acc = fp_add<FrParams>(acc, t);
  }
  tot = block_sum_by_lane<L, BLOCK>(acc);
  if (threadIdx.x < L) store_tagged(pub.vals, threadIdx.x, tot, pub.value);
}

template <int L, bool SAME, bool FUSED, int BLOCK = kBlock, class WAITER = MailWaiter, bool CG = false>
JA_DEV void round_prod_body(const FusedPolys& P, int d, Challenge r, const WAITER& waiter, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in,
                            int bits_in, size_t g_begin, size_t g_end, Fr* partials /* [nb][L] */, unsigned int* counter,
                            const Publish& pub, unsigned int bx, unsigned int nb, size_t g_off = 0) {
  constexpr int GPB = BLOCK / L;
  const int li = threadIdx.x & (L - 1);
  const int group = threadIdx.x / L;
  const bool pad = li >= d;
  // SAME (x^d of one MLE): every lane reads polynomial 0; only lane 0 writes the bound array
  const int pi = (SAME || pad) ? 0 : li;
  const Fr* __restrict__ zin = P.in[pi];
  Fr* __restrict__ zout = P.out[pi];
  const size_t mask_in = (size_t(1) << bits_in) - 1;
  Fr a0, a1, a2, a3;
  {
    const size_t g = g_begin + group;
    const size_t gl = g < g_end ? g : g_begin;
    if (FUSED) { a0 = fr_ld<CG>(zin + 4 * gl); a1 = fr_ld<CG>(zin + 4 * gl + 1); a2 = fr_ld<CG>(zin + 4 * gl + 2); a3 = fr_ld<CG>(zin + 4 * gl + 3); }
    else { a0 = fr_ld<CG>(zin + 2 * gl); a1 = fr_ld<CG>(zin + 2 * gl + 1); a2 = a0; a3 = a1; }
  }
  if (FUSED && !waiter(r)) return;
  Fr outer = fp_zero<FrParams>(), inner = fp_zero<FrParams>();
  size_t cur_xout = ~size_t(0);
  for (size_t base = g_begin; base < g_end; base += GPB) {
    const size_t g = base + group;
    const bool active = g < g_end;
    const size_t gl = active ? g : g_begin;
    if (base != g_begin) {
      if (FUSED) { a0 = fr_ld<CG>(zin + 4 * gl); a1 = fr_ld<CG>(zin + 4 * gl + 1); a2 = fr_ld<CG>(zin + 4 * gl + 2); a3 = fr_ld<CG>(zin + 4 * gl + 3); }
      else { a0 = fr_ld<CG>(zin + 2 * gl); a1 = fr_ld<CG>(zin + 2 * gl + 1); }
    }
    Fr p0, dp;
    if (pad) { p0 = fp_one<FrParams>(); dp = fp_zero<FrParams>(); }
    else if (FUSED) {
      p0 = fp_add<FrParams>(a0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a1, a0), r));
      const Fr p1 = fp_add<FrParams>(a2, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a3, a2), r));
      if (active && (!SAME || li == 0)) { fp_store(zout + 2 * gl, p0); fp_store(zout + 2 * gl + 1, p1); }
      dp = fp_sub<FrParams>(p1, p0);
    } else {
      p0 = a0;
      dp = fp_sub<FrParams>(a1, p0);
    }
    const Fr pv = lane_product<L>(p0, dp, pad, d, li);
    if (active) {
      const size_t x_out = (g + g_off) >> bits_in;
      if (x_out != cur_xout) {
        if (cur_xout != ~size_t(0)) {
          outer = fp_add<FrParams>(outer, fp_mul<FrParams>(fp_load(e_out + cur_xout), inner));
          inner = fp_zero<FrParams>();
        }
        cur_xout = x_out;
      }
      inner = fp_add<FrParams>(inner, fp_mul<FrParams>(fp_load(e_in + ((g + g_off) & mask_in)), pv));
    }
  }
  if (cur_xout != ~size_t(0)) outer = fp_add<FrParams>(outer, fp_mul<FrParams>(fp_load(e_out + cur_xout), inner));
  const Fr tot = block_sum_by_lane<L, BLOCK>(outer);
  prod_tail<L, BLOCK>(tot, partials, counter, pub, bx, nb);
}

// Latency variant for small slabs of a product of 9..16 factors: 64 threads per pair, thread (li, s) = (polynomial, quarter of
// the 16 evaluation points X = 1..15 and "infinity").  Each thread forms its polynomial's 4 values, then the 16 lanes of a
// quarter multiply across polynomials: exchange on bit 3 (4 values -> 2, two products), on bit 2 (2 -> 1, one product) and a
// multiplying butterfly on bits 1, 0: 5 dependent products per thread instead of 15, at 1.3x the total work and 4x the
// (L2-resident) loads.  Block = 128 threads = 2 pairs per pass of its loop.  It is a LATENCY form: on large slabs both forms reach
// ~0.72 of the field-mul peak and this one does 1.6x the products (600 vs 394 us at 2^16 pairs), so the launcher uses it up
// to kWideMaxPairs only (JA_BIGWIDE=1 enables it on large slabs for experiments).
constexpr int kWideBlock = 128;
constexpr size_t kWideMaxPairs = 256;      // measured on B200: 64 threads per pair wins up to 2^8 pairs, loses from 2^10 (scripts/small_probe.py)
constexpr size_t kSmallMaxPairs = 1024;    // 128-thread blocks up to here
constexpr size_t kBigWideMinPairs = 2048;  // JA_BIGWIDE=1 only: the 64-threads-per-pair form with several pairs per block
// sub-grids of the large-slab form: product blocks (3 of the 4 resident blocks per SM) and booleanity blocks
static inline __host__ __device__ size_t big_wide_ppb_prod(size_t G) { size_t p = (G + (size_t)kSMs * 3 - 1) / ((size_t)kSMs * 3); return (p + 1) & ~size_t(1); }
static inline __host__ __device__ size_t big_wide_ppb_bool(size_t G) { size_t p = (G + (size_t)kSMs - 1) / (size_t)kSMs; return (p + 7) & ~size_t(7); }
template <bool FUSED, int BLOCK = kWideBlock, class WAITER = MailWaiter, bool CG = false>
JA_DEV void round_prod16_wide_body(const FusedPolys& P, int d, Challenge r, const WAITER& waiter, const Fr* __restrict__ e_out,
                                   const Fr* __restrict__ e_in, int bits_in, size_t g_begin, size_t g_end,
                                   Fr* partials /* [nb][16] */, unsigned int* counter,
                                   const Publish& pub, unsigned int bx, unsigned int nb, size_t g_off = 0) {
  constexpr int L = 16, PPB = BLOCK / 64;
  const int t = threadIdx.x & 63, li = t & 15, s = t >> 4, group = threadIdx.x >> 6;
  const bool pad = li >= d;
  const bool hi8 = (li & 8) != 0, hi4 = (li & 4) != 0;
  const size_t mask_in = (size_t(1) << bits_in) - 1;
  const Fr* __restrict__ zin = P.in[pad ? 0 : li];
  Fr* __restrict__ zout = P.out[pad ? 0 : li];
  Fr a0, a1, a2, a3;
  {
    const size_t g = g_begin + group;
    const size_t gl = g < g_end ? g : g_begin;
    if (FUSED) { a0 = fr_ld<CG>(zin + 4 * gl); a1 = fr_ld<CG>(zin + 4 * gl + 1); a2 = fr_ld<CG>(zin + 4 * gl + 2); a3 = fr_ld<CG>(zin + 4 * gl + 3); }
    else { a0 = fr_ld<CG>(zin + 2 * gl); a1 = fr_ld<CG>(zin + 2 * gl + 1); a2 = a0; a3 = a1; }
  }
  if (FUSED && !waiter(r)) return;
  Fr outer = fp_zero<FrParams>(), inner = fp_zero<FrParams>();
  size_t cur_xout = ~size_t(0);
  for (size_t base = g_begin; base < g_end; base += PPB) {      // uniform trip count: every lane joins the shuffles
    const size_t g = base + group;
    const bool active = g < g_end;
    const size_t gl = active ? g : g_begin;
    if (base != g_begin) {
      if (FUSED) { a0 = fr_ld<CG>(zin + 4 * gl); a1 = fr_ld<CG>(zin + 4 * gl + 1); a2 = fr_ld<CG>(zin + 4 * gl + 2); a3 = fr_ld<CG>(zin + 4 * gl + 3); }
      else { a0 = fr_ld<CG>(zin + 2 * gl); a1 = fr_ld<CG>(zin + 2 * gl + 1); }
    }
    Fr p0, dp;
    if (pad) { p0 = fp_one<FrParams>(); dp = fp_zero<FrParams>(); }
    else if (FUSED) {
      p0 = fp_add<FrParams>(a0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a1, a0), r));
      const Fr p1 = fp_add<FrParams>(a2, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a3, a2), r));
      if (active && s == 0) { fp_store(zout + 2 * gl, p0); fp_store(zout + 2 * gl + 1, p1); }
      dp = fp_sub<FrParams>(p1, p0);
    } else {
      p0 = a0;
      dp = fp_sub<FrParams>(a1, p0);
    }
    // my points k = 4 s + j: X = k + 1 for k < 15, k = 15 the leading coefficient (pad lanes: the constant 1)
    const Fr dp4 = fp_dbl<FrParams>(fp_dbl<FrParams>(dp));
    Fr off = (s & 1) ? dp4 : fp_zero<FrParams>();
    if (s & 2) off = fp_add<FrParams>(off, fp_dbl<FrParams>(dp4));
    const Fr v0 = fp_add<FrParams>(fp_add<FrParams>(p0, dp), off);
    const Fr v1 = fp_add<FrParams>(v0, dp), v2 = fp_add<FrParams>(v1, dp);
    const Fr v3 = s == 3 ? (pad ? p0 : dp) : fp_add<FrParams>(v2, dp);
    const Fr n0 = fp_mul<FrParams>(fr_select(hi8, v2, v0), fr_shfl_xor(fr_select(hi8, v0, v2), 8));
    const Fr n1 = fp_mul<FrParams>(fr_select(hi8, v3, v1), fr_shfl_xor(fr_select(hi8, v1, v3), 8));
    Fr w = fp_mul<FrParams>(fr_select(hi4, n1, n0), fr_shfl_xor(fr_select(hi4, n0, n1), 4));
    w = fp_mul<FrParams>(w, fr_shfl_xor(w, 2));
    w = fp_mul<FrParams>(w, fr_shfl_xor(w, 1));
    // point k = 4 s + 2 hi8 + hi4 is complete in all four lanes that share (s, hi8, hi4); lane (li & 3) == 0 weighs it
    if (active && (li & 3) == 0) {
      const size_t x_out = (g + g_off) >> bits_in;
      if (x_out != cur_xout) {
        if (cur_xout != ~size_t(0)) {
          outer = fp_add<FrParams>(outer, fp_mul<FrParams>(fp_load(e_out + cur_xout), inner));
          inner = fp_zero<FrParams>();
        }
        cur_xout = x_out;
      }
      inner = fp_add<FrParams>(inner, fp_mul<FrParams>(fp_load(e_in + ((g + g_off) & mask_in)), w));
    }
  }
  if (cur_xout != ~size_t(0)) outer = fp_add<FrParams>(outer, fp_mul<FrParams>(fp_load(e_out + cur_xout), inner));
  __shared__ Fr s_red[PPB][L];
  if ((li & 3) == 0) s_red[group][4 * s + (hi8 ? 2 : 0) + (hi4 ? 1 : 0)] = outer;
  __syncthreads();
  Fr tot = fp_zero<FrParams>();
  if (threadIdx.x < L) {
#pragma unroll
    for (int gp = 0; gp < PPB; gp++) tot = fp_add<FrParams>(tot, s_red[gp][threadIdx.x]);
  }
  prod_tail<L, BLOCK>(tot, partials, counter, pub, bx, nb);
}

template <int L, bool SAME, bool FUSED>
__global__ void __launch_bounds__(kBlock, L == 16 ? 2 : 0)      // 0 = unspecified; L = 16: two blocks per SM (128 registers, 128 B of spills) beat one at 168
k_round_prod(FusedPolys P, int d, Challenge r, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in, int bits_in, size_t G,
             size_t pairs_per_block, Fr* partials /* [gridDim.x][L] */, unsigned int* counter, Publish pub, size_t g_off = 0,
             MailRef mail = MailRef{}) {
  const size_t g_begin = (size_t)blockIdx.x * pairs_per_block;
  const size_t g_end = g_begin + pairs_per_block < G ? g_begin + pairs_per_block : G;
  round_prod_body<L, SAME, FUSED>(P, d, r, MailWaiter{mail}, e_out, e_in, bits_in, g_begin, g_end, partials, counter, pub, blockIdx.x, gridDim.x, g_off);
}
template <bool FUSED>
__global__ void __launch_bounds__(kWideBlock)
k_round_prod16_wide(FusedPolys P, int d, Challenge r, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in, int bits_in, size_t G,
                    size_t pairs_per_block, Fr* partials /* [gridDim.x][16] */, unsigned int* counter, Publish pub, size_t g_off = 0,
                    MailRef mail = MailRef{}) {
  const size_t g_begin = (size_t)blockIdx.x * pairs_per_block;
  const size_t g_end = g_begin + pairs_per_block < G ? g_begin + pairs_per_block : G;
  round_prod16_wide_body<FUSED>(P, d, r, MailWaiter{mail}, e_out, e_in, bits_in, g_begin, g_end, partials, counter, pub, blockIdx.x, gridDim.x, g_off);
}

// ---- booleanity phase 2, lane-parallel (booleanity.rs:254-301) ---------------------------------------------------------
// [sum_i gamma_i h_i0 (h_i0 - 1), sum_i gamma_i (dh_i)^2] per pair: a group of L = next_pow2(d) lanes owns one pair, lane i
// loads polynomial i (with the fused bind), forms its two terms (4 products) and the group adds them up with shuffles —
// d times more threads and d times shorter dependency chains than one thread looping over the d polynomials.
template <int L, bool FUSED, int BLOCK = kBlock, class WAITER = MailWaiter, bool CG = false>
JA_DEV void round_bool_body(const FusedPolys& P, int d, Challenge r, const WAITER& waiter, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in,
                            int bits_in, size_t g_begin, size_t g_end, const Fr* __restrict__ gammas, Fr* partials,
                            unsigned int* counter, const Publish& pub, unsigned int bx, unsigned int nb) {
  constexpr int GPB = BLOCK / L;
  const int li = threadIdx.x & (L - 1);
  const int group = threadIdx.x / L;
  const bool pad = li >= d;
  const Fr* __restrict__ zin = P.in[pad ? 0 : li];
  Fr* __restrict__ zout = P.out[pad ? 0 : li];
  const Fr gm = pad ? fp_zero<FrParams>() : fp_load(gammas + li);
  const size_t mask_in = (size_t(1) << bits_in) - 1;
  Fr a0, a1, a2, a3;
  {
    const size_t g = g_begin + group;
    const size_t gl = g < g_end ? g : g_begin;
    if (FUSED) { a0 = fr_ld<CG>(zin + 4 * gl); a1 = fr_ld<CG>(zin + 4 * gl + 1); a2 = fr_ld<CG>(zin + 4 * gl + 2); a3 = fr_ld<CG>(zin + 4 * gl + 3); }
    else { a0 = fr_ld<CG>(zin + 2 * gl); a1 = fr_ld<CG>(zin + 2 * gl + 1); a2 = a0; a3 = a1; }
  }
  if (FUSED && !waiter(r)) return;
  Fr outer[2], inner[2];
#pragma unroll
  for (int k = 0; k < 2; k++) { outer[k] = fp_zero<FrParams>(); inner[k] = fp_zero<FrParams>(); }
  size_t cur_xout = ~size_t(0);
  for (size_t base = g_begin; base < g_end; base += GPB) {      // uniform trip count: every lane joins the shuffles
    const size_t g = base + group;
    const bool active = g < g_end;
    const size_t gl = active ? g : g_begin;
    if (base != g_begin) {
      if (FUSED) { a0 = fr_ld<CG>(zin + 4 * gl); a1 = fr_ld<CG>(zin + 4 * gl + 1); a2 = fr_ld<CG>(zin + 4 * gl + 2); a3 = fr_ld<CG>(zin + 4 * gl + 3); }
      else { a0 = fr_ld<CG>(zin + 2 * gl); a1 = fr_ld<CG>(zin + 2 * gl + 1); }
    }
    Fr v0 = fp_zero<FrParams>(), v1 = fp_zero<FrParams>();
    if (!pad) {
      Fr h0, h1;
      if (FUSED) {
        h0 = fp_add<FrParams>(a0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a1, a0), r));
        h1 = fp_add<FrParams>(a2, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a3, a2), r));
        if (active) { fp_store(zout + 2 * gl, h0); fp_store(zout + 2 * gl + 1, h1); }
      } else {
        h0 = a0;
        h1 = a1;
      }
      const Fr b = fp_sub<FrParams>(h1, h0);
      v0 = fp_mul<FrParams>(fp_mul<FrParams>(gm, h0), fp_sub<FrParams>(h0, fp_one<FrParams>()));
      v1 = fp_mul<FrParams>(fp_mul<FrParams>(gm, b), b);
    }
#pragma unroll
    for (int dlt = L / 2; dlt >= 1; dlt >>= 1) {
      v0 = fp_add<FrParams>(v0, fr_shfl_xor(v0, dlt));
      v1 = fp_add<FrParams>(v1, fr_shfl_xor(v1, dlt));
    }
    if (active && li == 0) {
      const size_t x_out = g >> bits_in;
      if (x_out != cur_xout) {
        if (cur_xout != ~size_t(0)) {
          const Fr eo = fp_load(e_out + cur_xout);
#pragma unroll
          for (int k = 0; k < 2; k++) { outer[k] = fp_add<FrParams>(outer[k], fp_mul<FrParams>(eo, inner[k])); inner[k] = fp_zero<FrParams>(); }
        }
        cur_xout = x_out;
      }
      const Fr ei = fp_load(e_in + (g & mask_in));
      inner[0] = fp_add<FrParams>(inner[0], fp_mul<FrParams>(ei, v0));
      inner[1] = fp_add<FrParams>(inner[1], fp_mul<FrParams>(ei, v1));
    }
  }
  if (cur_xout != ~size_t(0)) {
    const Fr eo = fp_load(e_out + cur_xout);
#pragma unroll
    for (int k = 0; k < 2; k++) outer[k] = fp_add<FrParams>(outer[k], fp_mul<FrParams>(eo, inner[k]));
  }
  grid_sum_ex<2>(outer, partials, counter, pub.vals, bx, nb, pub.value);
}

template <int L, bool FUSED>
__global__ void __launch_bounds__(kBlock)
k_round_bool(FusedPolys P, int d, Challenge r, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in, int bits_in, size_t G,
             size_t pairs_per_block, const Fr* __restrict__ gammas, Fr* partials, unsigned int* counter, Publish pub,
             MailRef mail = MailRef{}) {
  const size_t g_begin = (size_t)blockIdx.x * pairs_per_block;
  const size_t g_end = g_begin + pairs_per_block < G ? g_begin + pairs_per_block : G;
  round_bool_body<L, FUSED>(P, d, r, MailWaiter{mail}, e_out, e_in, bits_in, g_begin, g_end, gammas, partials, counter, pub, blockIdx.x, gridDim.x);
}

// ---- RA one-hot checks: RaVirtual (product of d) and Booleanity phase 2 of the same batch in ONE launch -------------------
// The two instances of a round are independent; launched back to back on one stream they serialise two latency-bound
// kernels.  blockIdx.y selects the body, each with its own sub-grid, scratch and result slot.  Blocks beyond a body's
// sub-grid exit at once (block (0, 0), the mailbox relay, always belongs to body A).
// BLOCK = kWideBlock: the small-slab variant (128-thread blocks); WIDE: the product of 9..16 factors runs 64 threads per pair.
struct PairArgs {
  FusedPolys P;
  int d;
  const Fr* e_out; const Fr* e_in;
  int bits_in;
  unsigned int nb;               // blocks of this body
  unsigned long long G, ppb;
  const Fr* gammas;
  Fr* partials; unsigned int* counter;
  Publish pub;
};
template <int L, bool FUSED, int BLOCK = kBlock, bool WIDE = false>
__global__ void __launch_bounds__(BLOCK, (L == 16 && BLOCK == kBlock && !WIDE) ? 2 : 0)
k_round_prod_bool(PairArgs A, PairArgs B, Challenge r, MailRef mail = MailRef{}) {
  const MailWaiter waiter{mail};
  if (blockIdx.y == 0) {
    if (blockIdx.x >= A.nb) return;
    const size_t g_begin = (size_t)blockIdx.x * (size_t)A.ppb;
    const size_t g_end = g_begin + (size_t)A.ppb < (size_t)A.G ? g_begin + (size_t)A.ppb : (size_t)A.G;
    if constexpr (L == 16 && BLOCK == kWideBlock && WIDE)
      round_prod16_wide_body<FUSED>(A.P, A.d, r, waiter, A.e_out, A.e_in, A.bits_in, g_begin, g_end, A.partials, A.counter, A.pub, blockIdx.x, A.nb);
    else
      round_prod_body<L, false, FUSED, BLOCK>(A.P, A.d, r, waiter, A.e_out, A.e_in, A.bits_in, g_begin, g_end, A.partials, A.counter, A.pub, blockIdx.x, A.nb);
  } else {
    if (blockIdx.x >= B.nb) return;
    const size_t g_begin = (size_t)blockIdx.x * (size_t)B.ppb;
    const size_t g_end = g_begin + (size_t)B.ppb < (size_t)B.G ? g_begin + (size_t)B.ppb : (size_t)B.G;
    round_bool_body<L, FUSED, BLOCK>(B.P, B.d, r, waiter, B.e_out, B.e_in, B.bits_in, g_begin, g_end, B.gammas, B.partials, B.counter, B.pub, blockIdx.x, B.nb);
  }
}

// ---- family D (plain products at X in {0,2,3}, HighToLow), fused bind in place ------------------------------------------
template <int NPOLY, bool FUSED, class WAITER, bool CG>
JA_DEV void round_dot_body(const FusedPolys& P, Challenge r, const WAITER& waiter, size_t G /* pairs of the evaluated array: it has 2G entries */,
                           Fr* partials, unsigned int* counter, const Publish& pub, unsigned int bx, unsigned int nb) {
  if (FUSED && !waiter(r)) return;
  constexpr int NOUT = NPOLY;
  Fr acc[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; k++) acc[k] = fp_zero<FrParams>();
  const size_t stride = (size_t)nb * blockDim.x;
  for (size_t i = (size_t)bx * blockDim.x + threadIdx.x; i < G; i += stride) {
    Fr prod[NOUT];
#pragma unroll
    for (int q = 0; q < NPOLY; q++) {
      Fr a, b;
      if (FUSED) {
        Fr* z = P.out[q];      // in place
        const Fr a0 = CG ? fr_ld<true>(z + i) : fp_load_rw(z + i), a1 = CG ? fr_ld<true>(z + i + 2 * G) : fp_load_rw(z + i + 2 * G),
                 b0 = CG ? fr_ld<true>(z + i + G) : fp_load_rw(z + i + G), b1 = CG ? fr_ld<true>(z + i + 3 * G) : fp_load_rw(z + i + 3 * G);
        a = fp_add<FrParams>(a0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a1, a0), r));
        b = fp_add<FrParams>(b0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(b1, b0), r));
        fp_store(z + i, a);
        fp_store(z + i + G, b);
      } else {
        a = fr_ld<CG>(P.in[q] + i);
        b = fr_ld<CG>(P.in[q] + i + G);
      }
      const Fr m = fp_sub<FrParams>(b, a);
      const Fr e = fp_add<FrParams>(b, m);   // X = 2
      if (q == 0) { prod[0] = a; prod[1] = e; } else { prod[0] = fp_mul<FrParams>(prod[0], a); prod[1] = fp_mul<FrParams>(prod[1], e); }
      if (NOUT == 3) {
        const Fr e3 = fp_add<FrParams>(e, m);  // X = 3
        if (q == 0) prod[2] = e3; else prod[2] = fp_mul<FrParams>(prod[2], e3);
      }
    }
#pragma unroll
    for (int k = 0; k < NOUT; k++) acc[k] = fp_add<FrParams>(acc[k], prod[k]);
  }
  grid_sum_ex<NOUT>(acc, partials, counter, pub.vals, bx, nb, pub.value);
}
template <int NPOLY, bool FUSED>
__global__ void __launch_bounds__(kBlock)
k_round_dot(FusedPolys P, Challenge r, size_t G, Fr* partials, unsigned int* counter, Publish pub, MailRef mail = MailRef{}) {
  round_dot_body<NPOLY, FUSED, MailWaiter, false>(P, r, MailWaiter{mail}, G, partials, counter, pub, blockIdx.x, gridDim.x);
}

// ---- opening reduction, HighToLow (opening_reduction.rs:355-403 dense, :630-673 one-hot cycle rounds) --------------------
// q_i(0) = sum_{j < half} E_in[j >> bits_out] * E_out[j & mask] * P_i[j] for d polynomials that share one opening point
// (blockIdx.y = polynomial; one sum per polynomial: each is its own sumcheck instance).  FUSED: bind the previous
// challenge in place first (P[j] <- P[j] + r (P[j + 2 half] - P[j]) for j < 2 half).
// counters: [0, d) per-row block counters, [63] rows finished.
template <bool FUSED>
__global__ void __launch_bounds__(kBlock)
k_round_open(FusedPolys P, Challenge r, const Fr* __restrict__ e_out, const Fr* __restrict__ e_in, int bits_out, size_t half,
             Fr* partials /* [gridDim.y][gridDim.x] */, unsigned int* counters, Publish pub) {
  const int row = blockIdx.y;
  Fr* __restrict__ z = P.out[row];
  const size_t mask_out = (size_t(1) << bits_out) - 1;
  Fr acc[1];
  acc[0] = fp_zero<FrParams>();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += stride) {
    Fr h;
    if (FUSED) {
      const Fr a0 = fp_load_rw(z + j), a1 = fp_load_rw(z + j + 2 * half), b0 = fp_load_rw(z + j + half), b1 = fp_load_rw(z + j + 3 * half);
      h = fp_add<FrParams>(a0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a1, a0), r));
      const Fr h2 = fp_add<FrParams>(b0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(b1, b0), r));
      fp_store(z + j, h);
      fp_store(z + j + half, h2);
    } else {
      h = fp_load(P.in[row] + j);
    }
    const Fr w = fp_mul<FrParams>(fp_load(e_in + (j >> bits_out)), fp_load(e_out + (j & mask_out)));
    acc[0] = fp_add<FrParams>(acc[0], fp_mul<FrParams>(w, h));
  }
  // row-level reduction: the row's blocks behave like a 1-D grid of their own
  __shared__ Fr s_part[kBlock / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Fr v = fr_warp_sum(acc[0]);
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  Fr tot = fp_zero<FrParams>();
  if (warp == 0) {
    tot = lane < (kBlock >> 5) ? s_part[lane] : fp_zero<FrParams>();
    tot = fr_warp_sum(tot);
  }
  Fr* row_part = partials + (size_t)row * gridDim.x;
  if (gridDim.x > 1) {
    if (threadIdx.x == 0) fp_store(row_part + blockIdx.x, tot);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicInc(counters + row, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    Fr a = fp_zero<FrParams>();
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
      const volatile uint32_t* q = reinterpret_cast<const volatile uint32_t*>(row_part + b);
      Fr t;
#pragma unroll
      for (int i = 0; i < 8; i++) t.l[i] = q[i];
      a = fp_add<FrParams>(a, t);
    }
    a = fr_warp_sum(a);
    __syncthreads();
    if (lane == 0) s_part[warp] = a;
    __syncthreads();
    if (warp == 0) {
      tot = lane < (kBlock >> 5) ? s_part[lane] : fp_zero<FrParams>();
      tot = fr_warp_sum(tot);
    }
  }
  // publish the row's sum; the last row to finish raises the flag
  if (threadIdx.x == 0) {
    fp_store(pub.vals + row, tot);
    __threadfence_system();
    const bool all_done = gridDim.y == 1 || atomicInc(counters + 63, gridDim.y - 1) == gridDim.y - 1;
    if (all_done) { __threadfence_system(); *pub.seq = pub.value; }
  }
}

// ---- opening reduction, every group of a batch in ONE launch -----------------------------------------------------------
// The batched opening reduction runs hundreds of one-hot instances (opening_proof.rs:500-532): one k_round_open launch per
// group and round would be launch-bound (74 groups x 18 rounds for a nanoGPT proof).  Here blockIdx.y walks a device
// table of rows (one polynomial each, with its own eq tables and length); every row is the same HighToLow body as
// k_round_open.  Row sums go to a flat host-mapped array; the last row to finish raises the flag.
struct OpenRow {
  Fr* z;
  const Fr* e_out;
  const Fr* e_in;
  unsigned long long half;       // pairs evaluated: j < half
  unsigned int bits_out;
  unsigned int fused;            // bind the previous challenge first
  unsigned int out_index;        // slot of the row's sum in the host-mapped array
  unsigned int pad;
};
static __global__ void __launch_bounds__(kBlock)
k_round_open_rows(const OpenRow* __restrict__ rows, unsigned int n_rows /* active rows of the round, all launches */,
                  unsigned int row_off /* first row of this launch */, Challenge r, Fr* partials /* [gridDim.y][gridDim.x] */,
                  unsigned int* counters /* [n_rows] + 1 */, Fr* dev_vals /* [total_rows] row sums, device */, unsigned int total_rows,
                  Fr* host_vals, volatile unsigned int* host_seq, unsigned int seq_value) {
  const unsigned int ri = row_off + blockIdx.y;
  const OpenRow row = rows[ri];
  const size_t half = (size_t)row.half;
  unsigned int nblk = (unsigned int)((half + kBlock - 1) / kBlock);
  if (nblk > gridDim.x) nblk = gridDim.x;
  if (blockIdx.x >= nblk) return;
  Fr* __restrict__ z = row.z;
  const size_t mask_out = (size_t(1) << row.bits_out) - 1;
  Fr acc = fp_zero<FrParams>();
  const size_t stride = (size_t)nblk * blockDim.x;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += stride) {
    Fr h;
    if (row.fused) {
      const Fr a0 = fp_load_rw(z + j), a1 = fp_load_rw(z + j + 2 * half), b0 = fp_load_rw(z + j + half), b1 = fp_load_rw(z + j + 3 * half);
      h = fp_add<FrParams>(a0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(a1, a0), r));
      const Fr h2 = fp_add<FrParams>(b0, fp_mul_challenge<FrParams>(fp_sub<FrParams>(b1, b0), r));
      fp_store(z + j, h);
      fp_store(z + j + half, h2);
    } else {
      h = fp_load_rw(z + j);
    }
    const Fr w = fp_mul<FrParams>(fp_load(row.e_in + (j >> row.bits_out)), fp_load(row.e_out + (j & mask_out)));
    acc = fp_add<FrParams>(acc, fp_mul<FrParams>(w, h));
  }
  __shared__ Fr s_part[kBlock / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Fr v = fr_warp_sum(acc);
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  Fr tot = fp_zero<FrParams>();
  if (warp == 0) {
    tot = lane < (kBlock >> 5) ? s_part[lane] : fp_zero<FrParams>();
    tot = fr_warp_sum(tot);
  }
  if (nblk > 1) {
    Fr* row_part = partials + (size_t)blockIdx.y * gridDim.x;
    if (threadIdx.x == 0) fp_store(row_part + blockIdx.x, tot);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicInc(counters + ri, nblk - 1) == nblk - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    Fr a = fp_zero<FrParams>();
    for (unsigned b = threadIdx.x; b < nblk; b += blockDim.x) {
      const volatile uint32_t* q = reinterpret_cast<const volatile uint32_t*>(row_part + b);
      Fr t;
#pragma unroll
      for (int i = 0; i < 8; i++) t.l[i] = q[i];
      a = fp_add<FrParams>(a, t);
    }
    a = fr_warp_sum(a);
    __syncthreads();
    if (lane == 0) s_part[warp] = a;
    __syncthreads();
    if (warp == 0) {
      tot = lane < (kBlock >> 5) ? s_part[lane] : fp_zero<FrParams>();
      tot = fr_warp_sum(tot);
    }
  }
  // Row sums stay on the device; the block that finishes the LAST row ships the whole value array to the host in one
  // coalesced burst and raises the flag (hundreds of rows each doing their own 32-byte PCIe write + system fence cost
  // ~0.4 ms per round at 756 rows).
  __syncthreads();                                   // s_last is reused below
  if (threadIdx.x == 0) {
    fp_store(dev_vals + row.out_index, tot);
    __threadfence();
    s_last = n_rows == 1 || atomicInc(counters + n_rows, n_rows - 1) == n_rows - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const volatile uint4* src = reinterpret_cast<const volatile uint4*>(dev_vals);
  uint4* dst = reinterpret_cast<uint4*>(host_vals);
  for (unsigned int i = threadIdx.x; i < 2 * total_rows; i += blockDim.x) {
    uint4 v;
    v.x = src[i].x; v.y = src[i].y; v.z = src[i].z; v.w = src[i].w;
    dst[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *host_seq = seq_value;
}

}  // namespace ja
