// TEST INFRASTRUCTURE ONLY — CPU oracle (C++ restatement).  Never linked into the product library.
//
// BN254 Fr / Fq arithmetic: 4 x u64 Montgomery limbs, the same in-memory form as ark_ff::Fp<MontBackend, 4>
// (external: a16z/arkworks-algebra@76bb3a4, un-vendored).  Restates the call sites in
//   joltworks/src/field/ark.rs:76-297 (from_u64/from_i64/from_i128, mul_u64, inverse, from_bytes)
//   joltworks/src/field/challenge/mont_ark_u128.rs:51-92, macros.rs:274-286 (F x Challenge)
// Parity unpinned at the byte level: the reference holds no KATs for this path; pinned against the
// Python twin (oracle/pyref) and the reference's equivalence invariants in tests/.
#pragma once
#include <cstdint>
#include <cstring>

namespace orc {

typedef unsigned __int128 u128;

struct FrCfg {
  static constexpr uint64_t P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr uint64_t INV = 0xc2e1f593efffffffull;
  static constexpr uint64_t R[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
  static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
};
struct FqCfg {
  static constexpr uint64_t P[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr uint64_t INV = 0x87d20782e4866389ull;
  static constexpr uint64_t R[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
  static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
};

template <class C>
struct Fp {
  uint64_t l[4];

  static Fp zero() { return Fp{{0, 0, 0, 0}}; }
  static Fp one() { return Fp{{C::R[0], C::R[1], C::R[2], C::R[3]}}; }
  static Fp r2() { return Fp{{C::R2[0], C::R2[1], C::R2[2], C::R2[3]}}; }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  bool operator==(const Fp& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
  bool operator!=(const Fp& o) const { return !(*this == o); }

  static bool geq_p(const uint64_t* a) {
    for (int i = 3; i >= 0; i--) {
      if (a[i] > C::P[i]) return true;
      if (a[i] < C::P[i]) return false;
    }
    return true;
  }
  static void sub_p(uint64_t* a) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - C::P[i] - (uint64_t)b; a[i] = (uint64_t)t; b = (t >> 64) & 1; }
  }
  Fp operator+(const Fp& o) const {
    Fp r; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)l[i] + o.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (geq_p(r.l)) sub_p(r.l);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r; u128 b = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)l[i] - o.l[i] - (uint64_t)b; r.l[i] = (uint64_t)t; b = (t >> 64) & 1; }
    if (b) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.l[i] + C::P[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
    return r;
  }
  Fp operator-() const { return is_zero() ? *this : zero() - *this; }
  Fp& operator+=(const Fp& o) { *this = *this + o; return *this; }
  Fp& operator-=(const Fp& o) { *this = *this - o; return *this; }
  Fp& operator*=(const Fp& o) { *this = *this * o; return *this; }

  // CIOS Montgomery product (what ark-ff's MontBackend::mul_assign computes; result canonical)
  Fp operator*(const Fp& o) const {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      u128 c = 0;
      for (int j = 0; j < 4; j++) { c += (u128)l[j] * o.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * C::INV;
      c = (u128)m * C::P[0] + t[0]; c >>= 64;
      for (int j = 1; j < 4; j++) { c += (u128)m * C::P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fp r{{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_p(r.l)) sub_p(r.l);
    return r;
  }
  Fp sqr() const { return *this * *this; }
  Fp dbl() const { return *this + *this; }

  static Fp from_u64(uint64_t v) { Fp t{{v, 0, 0, 0}}; return t * r2(); }
  static Fp from_i64(int64_t v) { return v < 0 ? -from_u64((uint64_t)(-(v + 1)) + 1) : from_u64((uint64_t)v); }
  static Fp from_canonical(const uint64_t in[4]) { Fp t; memcpy(t.l, in, 32); return t * r2(); }
  void to_canonical(uint64_t out[4]) const { Fp o{{1, 0, 0, 0}}; Fp r = *this * o; memcpy(out, r.l, 32); }
  // raw limbs reinterpretation (MontU128Challenge -> Fr via from_bigint_unchecked, mont_ark_u128.rs:79-84)
  static Fp from_raw(const uint64_t* p) { Fp r; memcpy(r.l, p, 32); return r; }

  Fp pow(const uint64_t e[4]) const {
    Fp r = one();
    for (int i = 255; i >= 0; i--) { r = r.sqr(); if ((e[i / 64] >> (i % 64)) & 1) r = r * *this; }
    return r;
  }
  Fp inv() const { uint64_t e[4] = {C::P[0] - 2, C::P[1], C::P[2], C::P[3]}; return pow(e); }
  Fp mul_pow_2(unsigned k) const { Fp r = *this; for (unsigned i = 0; i < k; i++) r = r.dbl(); return r; }  // field/mod.rs:274-284
};

typedef Fp<FrCfg> Fr;
typedef Fp<FqCfg> Fq;

}  // namespace orc
