// 254-bit prime-field arithmetic for sm_100a: BN254 scalar field Fr and base field Fq.
//
// Device replacement for the arithmetic the reference gets from ark-ff (external) through
// joltworks/src/field/ark.rs:16-298 (JoltField for ark_bn254::Fr) and
// joltworks/src/field/challenge/mont_ark_u128.rs:25-92 (125-bit challenge held as Montgomery
// limbs [0,0,lo,hi]; `F * challenge` == mul_hi_bigint_u128, macros.rs:274-286).
//
// Representation: 8 x u32 little-endian limbs, Montgomery form with R = 2^256 — byte-identical
// to ark's BigInt<4> (4 x u64 LE), so host buffers are passed through unchanged.
// All results are canonical (< p), which is what makes GPU output bit-identical to the CPU prover.
//
// Multiplication is an operand-scanning Montgomery product on two staggered accumulators
// (EVEN at bit 0, ODD at bit 32).  Every 32x32->64 partial product lands on a 64-bit aligned slot
// of one of the two accumulators, so ptxas can fuse each mad.lo.cc/madc.hi.cc pair into a single
// IMAD.WIDE.U32(.X) with a predicate carry chain; the only extra work per row is one IADD3 for
// the stray limb, one IMAD for the quotient digit and two carry folds.
#pragma once
#include <cstdint>

namespace ja {

struct FrParams {
  static constexpr uint32_t P0 = 0xf0000001u, P1 = 0x43e1f593u, P2 = 0x79b97091u, P3 = 0x2833e848u,
                            P4 = 0x8181585du, P5 = 0xb85045b6u, P6 = 0xe131a029u, P7 = 0x30644e72u;
  static constexpr uint32_t INV = 0xefffffffu;  // -p^-1 mod 2^32
  // R mod p (Montgomery one)
  static constexpr uint32_t R0 = 0x4ffffffbu, R1 = 0xac96341cu, R2 = 0x9f60cd29u, R3 = 0x36fc7695u,
                            R4 = 0x7879462eu, R5 = 0x666ea36fu, R6 = 0x9a07df2fu, R7 = 0x0e0a77c1u;
  // R^2 mod p
  static constexpr uint32_t S0 = 0xae216da7u, S1 = 0x1bb8e645u, S2 = 0xe35c59e3u, S3 = 0x53fe3ab1u,
                            S4 = 0x53bb8085u, S5 = 0x8c49833du, S6 = 0x7f4e44a5u, S7 = 0x0216d0b1u;
};

struct FqParams {
  static constexpr uint32_t P0 = 0xd87cfd47u, P1 = 0x3c208c16u, P2 = 0x6871ca8du, P3 = 0x97816a91u,
                            P4 = 0x8181585du, P5 = 0xb85045b6u, P6 = 0xe131a029u, P7 = 0x30644e72u;
  static constexpr uint32_t INV = 0xe4866389u;
  static constexpr uint32_t R0 = 0xc58f0d9du, R1 = 0xd35d438du, R2 = 0xf5c70b3du, R3 = 0x0a78eb28u,
                            R4 = 0x7879462cu, R5 = 0x666ea36fu, R6 = 0x9a07df2fu, R7 = 0x0e0a77c1u;
  static constexpr uint32_t S0 = 0x538afa89u, S1 = 0xf32cfc5bu, S2 = 0xd44501fbu, S3 = 0xb5e71911u,
                            S4 = 0x0a417ff6u, S5 = 0x47ab1effu, S6 = 0xcab8351fu, S7 = 0x06d89f71u;
};

template <class M>
struct alignas(16) Fp {
  uint32_t l[8];
};

using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

#define JA_DEV __device__ __forceinline__
// Montgomery products stay force-inlined: making them real calls (one code copy per kernel, -DJA_MUL_CALL) to relieve the
// instruction cache of the latency-bound small launches was measured SLOWER everywhere (ABI spills around the calls:
// single-instance rounds 16.9 -> 19.6 us, fused ADD at 2^24 0.63 -> 0.40 of HBM).
#ifdef JA_MUL_CALL
#define JA_MUL_DEV __device__ __noinline__
#else
#define JA_MUL_DEV __device__ __forceinline__
#endif

template <class M> JA_DEV Fp<M> fp_zero() { Fp<M> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
template <class M> JA_DEV Fp<M> fp_one() {
  Fp<M> r; r.l[0] = M::R0; r.l[1] = M::R1; r.l[2] = M::R2; r.l[3] = M::R3;
  r.l[4] = M::R4; r.l[5] = M::R5; r.l[6] = M::R6; r.l[7] = M::R7; return r; }
template <class M> JA_DEV Fp<M> fp_r2() {
  Fp<M> r; r.l[0] = M::S0; r.l[1] = M::S1; r.l[2] = M::S2; r.l[3] = M::S3;
  r.l[4] = M::S4; r.l[5] = M::S5; r.l[6] = M::S6; r.l[7] = M::S7; return r; }

template <class M> JA_DEV bool fp_is_zero(const Fp<M>& a) {
  return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7]) == 0; }
template <class M> JA_DEV bool fp_eq(const Fp<M>& a, const Fp<M>& b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a.l[i] ^ b.l[i];
  return d == 0; }

// r = a - p if a >= p else a   (a < 2p)
template <class M> JA_DEV void fp_final_sub(uint32_t* a) {
  uint32_t t[8], borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
        "=r"(borrow)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "n"(M::P0), "n"(M::P1), "n"(M::P2), "n"(M::P3), "n"(M::P4), "n"(M::P5), "n"(M::P6), "n"(M::P7));
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = borrow ? a[i] : t[i];
}

template <class M> JA_DEV Fp<M> fp_add(const Fp<M>& a, const Fp<M>& b) {
  Fp<M> r;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
        "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  fp_final_sub<M>(r.l);  // a+b < 2p < 2^255: no carry out of limb 7
  return r;
}

template <class M> JA_DEV Fp<M> fp_sub(const Fp<M>& a, const Fp<M>& b) {
  Fp<M> r; uint32_t borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7]),
        "=r"(borrow)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
        "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  // add back p masked by the borrow (borrow is 0 or 0xffffffff)
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, %15;"
      : "+r"(r.l[0]), "+r"(r.l[1]), "+r"(r.l[2]), "+r"(r.l[3]), "+r"(r.l[4]), "+r"(r.l[5]), "+r"(r.l[6]), "+r"(r.l[7])
      : "r"(M::P0 & borrow), "r"(M::P1 & borrow), "r"(M::P2 & borrow), "r"(M::P3 & borrow),
        "r"(M::P4 & borrow), "r"(M::P5 & borrow), "r"(M::P6 & borrow), "r"(M::P7 & borrow));
  return r;
}

template <class M> JA_DEV Fp<M> fp_neg(const Fp<M>& a) {
  return fp_sub<M>(fp_zero<M>(), a);
}

template <class M> JA_DEV Fp<M> fp_dbl(const Fp<M>& a) { return fp_add<M>(a, a); }

// ---- staggered-accumulator rows -------------------------------------------------------------
// acc[0..7] += x[0,2,4,6] * y  (64-bit aligned slots), carry-out returned in acc[8] (added).
#define JA_ROW_ACC(acc, x0, x2, x4, x6, y)                                                      \
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"                                                      \
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                                     \
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"                                                    \
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                                    \
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"                                                    \
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                                    \
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"                                                    \
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                                    \
      "addc.u32 %8, %8, 0;"                                                                     \
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),     \
        "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8])                                                \
      : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y))

// Montgomery product core.  NB = number of b limbs consumed (8 = full product a*b*2^-256;
// 4 = "challenge" product a*b'*2^-128 for b = b' << 128).  Output < 2p in r[0..7].
template <class M, int NB>
JA_DEV void fp_mont_rows(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t E[9], O[9];
#pragma unroll
  for (int i = 0; i < 9; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
  for (int i = 0; i < NB; i++) {
    if (i > 0) {
      // divide by 2^32: EVEN' = ODD + E[1] (carry joins the ODD' chain), ODD' = EVEN >> 64
      uint32_t nE[9], nO[9];
      asm("add.cc.u32 %0, %9, %10;\n\t"
          "madc.lo.cc.u32 %1, %11, %15, %16;\n\t"
          "madc.hi.cc.u32 %2, %11, %15, %17;\n\t"
          "madc.lo.cc.u32 %3, %12, %15, %18;\n\t"
          "madc.hi.cc.u32 %4, %12, %15, %19;\n\t"
          "madc.lo.cc.u32 %5, %13, %15, %20;\n\t"
          "madc.hi.cc.u32 %6, %13, %15, %21;\n\t"
          "madc.lo.cc.u32 %7, %14, %15, %22;\n\t"
          "madc.hi.u32 %8, %14, %15, 0;"
          : "=r"(nE[0]), "=r"(nO[0]), "=r"(nO[1]), "=r"(nO[2]), "=r"(nO[3]), "=r"(nO[4]), "=r"(nO[5]),
            "=r"(nO[6]), "=r"(nO[7])
          : "r"(O[0]), "r"(E[1]), "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(b[i]),
            "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]));
#pragma unroll
      for (int k = 1; k < 8; k++) nE[k] = O[k];
      nE[8] = 0; nO[8] = 0;
#pragma unroll
      for (int k = 0; k < 9; k++) { E[k] = nE[k]; O[k] = nO[k]; }
    } else {
      JA_ROW_ACC(O, a[1], a[3], a[5], a[7], b[i]);
    }
    JA_ROW_ACC(E, a[0], a[2], a[4], a[6], b[i]);
    uint32_t m = E[0] * M::INV;
    {
      const uint32_t p0 = M::P0, p2 = M::P2, p4 = M::P4, p6 = M::P6;
      JA_ROW_ACC(E, p0, p2, p4, p6, m);
      const uint32_t p1 = M::P1, p3 = M::P3, p5 = M::P5, p7 = M::P7;
      JA_ROW_ACC(O, p1, p3, p5, p7, m);
    }
  }
  // final divide by 2^32 and merge: r = E[1] + O + ((E >> 64) << 32)
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]),
        "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]));
}

// ---- delayed reduction (the reference's mul_unreduced::<9> + from_montgomery_reduce, field/mod.rs:286-310) ----------------
// A 512-bit accumulator takes up to 16 unreduced products of canonical elements (16 p^2 < 2^512); one reduction per flush
// instead of one per product: 64 IMAD.WIDE per accumulated product instead of 128.
struct FpWide { uint32_t l[16]; };
JA_DEV FpWide fpw_zero() { FpWide w;
#pragma unroll
  for (int i = 0; i < 16; i++) w.l[i] = 0; return w; }
// acc += a * b as integers (a, b < p)
template <class M> JA_DEV void fpw_mul_acc(FpWide& acc, const Fp<M>& a, const Fp<M>& b) {
  uint32_t E[17], O[17];
#pragma unroll
  for (int i = 0; i < 17; i++) { E[i] = 0; O[i] = 0; }
  // partial product a[j] * b[i] sits at limb i + j: even positions accumulate in E (64-bit slots at limbs 0, 2, ...), odd
  // positions in O (the same slots shifted by one limb), so every mad.lo.cc / madc.hi.cc pair is one IMAD.WIDE
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    JA_ROW_ACC((E + i), a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
    JA_ROW_ACC((O + i), a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
    JA_ROW_ACC((O + i), a.l[0], a.l[2], a.l[4], a.l[6], b.l[i + 1]);
    JA_ROW_ACC((E + i + 2), a.l[1], a.l[3], a.l[5], a.l[7], b.l[i + 1]);
  }
  // T = E + (O << 32); acc += T
  uint32_t T[16];
  asm("add.cc.u32 %0, %16, 0;\n\t"
      "addc.cc.u32 %1, %17, %32;\n\t"  "addc.cc.u32 %2, %18, %33;\n\t"  "addc.cc.u32 %3, %19, %34;\n\t"
      "addc.cc.u32 %4, %20, %35;\n\t"  "addc.cc.u32 %5, %21, %36;\n\t"  "addc.cc.u32 %6, %22, %37;\n\t"
      "addc.cc.u32 %7, %23, %38;\n\t"  "addc.cc.u32 %8, %24, %39;\n\t"  "addc.cc.u32 %9, %25, %40;\n\t"
      "addc.cc.u32 %10, %26, %41;\n\t" "addc.cc.u32 %11, %27, %42;\n\t" "addc.cc.u32 %12, %28, %43;\n\t"
      "addc.cc.u32 %13, %29, %44;\n\t" "addc.cc.u32 %14, %30, %45;\n\t" "addc.u32 %15, %31, %46;"
      : "=r"(T[0]), "=r"(T[1]), "=r"(T[2]), "=r"(T[3]), "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7]), "=r"(T[8]), "=r"(T[9]),
        "=r"(T[10]), "=r"(T[11]), "=r"(T[12]), "=r"(T[13]), "=r"(T[14]), "=r"(T[15])
      : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]), "r"(E[10]),
        "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
        "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]),
        "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
  asm("add.cc.u32 %0, %0, %16;\n\t"
      "addc.cc.u32 %1, %1, %17;\n\t"   "addc.cc.u32 %2, %2, %18;\n\t"   "addc.cc.u32 %3, %3, %19;\n\t"
      "addc.cc.u32 %4, %4, %20;\n\t"   "addc.cc.u32 %5, %5, %21;\n\t"   "addc.cc.u32 %6, %6, %22;\n\t"
      "addc.cc.u32 %7, %7, %23;\n\t"   "addc.cc.u32 %8, %8, %24;\n\t"   "addc.cc.u32 %9, %9, %25;\n\t"
      "addc.cc.u32 %10, %10, %26;\n\t" "addc.cc.u32 %11, %11, %27;\n\t" "addc.cc.u32 %12, %12, %28;\n\t"
      "addc.cc.u32 %13, %13, %29;\n\t" "addc.cc.u32 %14, %14, %30;\n\t" "addc.u32 %15, %15, %31;"
      : "+r"(acc.l[0]), "+r"(acc.l[1]), "+r"(acc.l[2]), "+r"(acc.l[3]), "+r"(acc.l[4]), "+r"(acc.l[5]), "+r"(acc.l[6]), "+r"(acc.l[7]),
        "+r"(acc.l[8]), "+r"(acc.l[9]), "+r"(acc.l[10]), "+r"(acc.l[11]), "+r"(acc.l[12]), "+r"(acc.l[13]), "+r"(acc.l[14]), "+r"(acc.l[15])
      : "r"(T[0]), "r"(T[1]), "r"(T[2]), "r"(T[3]), "r"(T[4]), "r"(T[5]), "r"(T[6]), "r"(T[7]), "r"(T[8]), "r"(T[9]), "r"(T[10]),
        "r"(T[11]), "r"(T[12]), "r"(T[13]), "r"(T[14]), "r"(T[15]));
}
// acc * 2^-256 mod p, canonical: (lo + hi 2^256) 2^-256 = mont(lo, 1) + mont(hi, R)   (both operands < 2^256 are fine for the
// row reduction: its output is < (2^256 p + 2^256 p) / 2^256 = 2p)
template <class M> JA_DEV Fp<M> fpw_reduce(const FpWide& acc) {
  const uint32_t one_raw[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  const Fp<M> r_one = fp_one<M>();
  Fp<M> lo, hi;
  fp_mont_rows<M, 8>(lo.l, acc.l, one_raw);
  fp_final_sub<M>(lo.l);
  fp_mont_rows<M, 8>(hi.l, acc.l + 8, r_one.l);
  fp_final_sub<M>(hi.l);
  return fp_add<M>(lo, hi);
}

template <class M> JA_MUL_DEV Fp<M> fp_mul(const Fp<M>& a, const Fp<M>& b) {
  Fp<M> r;
  fp_mont_rows<M, 8>(r.l, a.l, b.l);
  fp_final_sub<M>(r.l);
  return r;
}
template <class M> JA_DEV Fp<M> fp_sqr(const Fp<M>& a) { return fp_mul<M>(a, a); }

// a * c where c is a 125-bit challenge held as Montgomery limbs [0,0,lo,hi] (u64) = u32 limbs
// [0,0,0,0,c0,c1,c2,c3]:  a*c*2^-256 == a*c'*2^-128, i.e. only 4 of the 8 rows are needed.
// (field/challenge/macros.rs:274-286 `mul_hi_bigint_u128`)
struct Challenge { uint32_t c[4]; };
template <class M> JA_MUL_DEV Fp<M> fp_mul_challenge(const Fp<M>& a, const Challenge& ch) {
  Fp<M> r;
  fp_mont_rows<M, 4>(r.l, a.l, ch.c);
  fp_final_sub<M>(r.l);
  return r;
}

// Montgomery form of a signed 64-bit integer (field/ark.rs:125-162 from_i32/from_i64): |v| * R^2 * R^-1, negated if v<0.
template <class M> JA_DEV Fp<M> fp_from_i64(long long v) {
  unsigned long long u = v < 0 ? (unsigned long long)(-(v + 1)) + 1ull : (unsigned long long)v;
  Fp<M> t = fp_zero<M>();
  t.l[0] = (uint32_t)u; t.l[1] = (uint32_t)(u >> 32);
  Fp<M> r = fp_mul<M>(t, fp_r2<M>());
  return v < 0 ? fp_neg<M>(r) : r;
}

// a * u64 small scalar (plain integer, not Montgomery): a*s mod p.  Uses s*R as the b operand would
// cost a full product; instead compute (a * [s0,s1]) with 2 rows then fix the 2^-64 by R^2-free trick:
// a*s = mont(a, s*R) and s*R = mont(s, R^2).  Kept simple: two products.
template <class M> JA_DEV Fp<M> fp_mul_u64(const Fp<M>& a, unsigned long long s) {
  Fp<M> t = fp_zero<M>();
  t.l[0] = (uint32_t)s; t.l[1] = (uint32_t)(s >> 32);
  return fp_mul<M>(a, fp_mul<M>(t, fp_r2<M>()));
}

// ---- 256-bit vectorised global access -------------------------------------------------------
template <class M> JA_DEV Fp<M> fp_load(const Fp<M>* p) {
  Fp<M> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 lo = __ldg(q), hi = __ldg(q + 1);
  r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
  r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
  return r;
}
// Coherent variant (plain ld.global) for operands that alias an OUTPUT of the same kernel (the in-place HighToLow binds):
// PTX requires memory read through ld.global.nc to stay read-only for the whole kernel.
template <class M> JA_DEV Fp<M> fp_load_rw(const Fp<M>* p) {
  Fp<M> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 lo = q[0], hi = q[1];
  r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
  r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
  return r;
}
template <class M> JA_DEV void fp_store(Fp<M>* p, const Fp<M>& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

}  // namespace ja
