// TEST INFRASTRUCTURE ONLY — CPU oracle (C++ restatement): Blake2b-256 (RFC 7693; the reference uses the
// `blake2 0.10.6` crate, external) and the Fiat–Shamir transcript of joltworks/src/transcripts/blake2b.rs:11-258,
// plus UniPoly (joltworks/src/poly/unipoly.rs) and gaussian elimination (utils/gaussian_elimination.rs:9-70).
// Parity unpinned at the byte level upstream; Blake2b itself is pinned against hashlib in tests/.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "field.hpp"

namespace orc {

// ---------------------------------------------------------------- Blake2b (RFC 7693), digest 32, no key
struct Blake2b256 {
  uint64_t h[8];
  uint8_t buf[128];
  size_t buflen = 0;
  uint64_t t = 0;
  static constexpr uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                     0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
  Blake2b256() { for (int i = 0; i < 8; i++) h[i] = IV[i]; h[0] ^= 0x01010000ull ^ 32ull; }
  static inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
  void compress(const uint8_t* block, bool last) {
    static const uint8_t S[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    uint64_t m[16], v[16];
    memcpy(m, block, 128);
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = IV[i]; }
    v[12] ^= t; v[14] = last ? ~v[14] : v[14];
    auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
      v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32);
      v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 24);
      v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16);
      v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 63);
    };
    for (int r = 0; r < 12; r++) {
      const uint8_t* s = S[r];
      G(0, 4, 8, 12, m[s[0]], m[s[1]]); G(1, 5, 9, 13, m[s[2]], m[s[3]]);
      G(2, 6, 10, 14, m[s[4]], m[s[5]]); G(3, 7, 11, 15, m[s[6]], m[s[7]]);
      G(0, 5, 10, 15, m[s[8]], m[s[9]]); G(1, 6, 11, 12, m[s[10]], m[s[11]]);
      G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
  }
  void update(const uint8_t* p, size_t n) {
    while (n) {
      if (buflen == 128) { t += 128; compress(buf, false); buflen = 0; }
      size_t k = 128 - buflen; if (k > n) k = n;
      memcpy(buf + buflen, p, k); buflen += k; p += k; n -= k;
    }
  }
  void finalize(uint8_t out[32]) {
    t += buflen;
    memset(buf + buflen, 0, 128 - buflen);
    compress(buf, true);
    memcpy(out, h, 32);
  }
};

// ---------------------------------------------------------------- transcript (blake2b.rs)
struct Transcript {
  uint8_t state[32];
  uint32_t n_rounds = 0;
  explicit Transcript(const std::string& label) {   // :81-100
    uint8_t pad[32] = {0};
    memcpy(pad, label.data(), label.size());
    Blake2b256 h; h.update(pad, 32); h.finalize(state);
  }
  Transcript(const uint8_t st[32], uint32_t rounds) : n_rounds(rounds) { memcpy(state, st, 32); }   // resume a running transcript
  Blake2b256 hasher() const {                       // :31-37
    Blake2b256 h; h.update(state, 32);
    uint8_t packed[32] = {0};
    packed[28] = (uint8_t)(n_rounds >> 24); packed[29] = (uint8_t)(n_rounds >> 16);
    packed[30] = (uint8_t)(n_rounds >> 8); packed[31] = (uint8_t)n_rounds;
    h.update(packed, 32);
    return h;
  }
  void finish(Blake2b256& h) { h.finalize(state); n_rounds++; }
  void append_message(const std::string& msg) {     // :109-122
    uint8_t pad[32] = {0}; memcpy(pad, msg.data(), msg.size());
    Blake2b256 h = hasher(); h.update(pad, 32); finish(h);
  }
  void append_bytes(const uint8_t* p, size_t n) { Blake2b256 h = hasher(); h.update(p, n); finish(h); }   // :124-128
  void append_u64(uint64_t x) {                     // :130-136
    uint8_t b[32] = {0};
    for (int i = 0; i < 8; i++) b[31 - i] = (uint8_t)(x >> (8 * i));
    append_bytes(b, 32);
  }
  void append_scalar(const Fr& x) {                 // :138-146: 32-byte big-endian canonical integer
    uint64_t c[4]; x.to_canonical(c);
    uint8_t b[32];
    for (int i = 0; i < 32; i++) b[31 - i] = (uint8_t)(c[i / 8] >> (8 * (i % 8)));
    append_bytes(b, 32);
  }
  void append_scalars(const std::vector<Fr>& xs) {  // :158-164
    append_message("begin_append_vector");
    for (auto& x : xs) append_scalar(x);
    append_message("end_append_vector");
  }
  // affine point as canonical big-endian x||y; infinity = 64 zero bytes (:166-187)
  void append_point(bool inf, const Fq& x, const Fq& y) {
    uint8_t b[64] = {0};
    if (!inf) {
      uint64_t cx[4], cy[4]; x.to_canonical(cx); y.to_canonical(cy);
      for (int i = 0; i < 32; i++) { b[31 - i] = (uint8_t)(cx[i / 8] >> (8 * (i % 8))); b[63 - i] = (uint8_t)(cy[i / 8] >> (8 * (i % 8))); }
    }
    append_bytes(b, 64);
  }
  void challenge_bytes32(uint8_t out[32]) { Blake2b256 h = hasher(); uint8_t r[32]; h.finalize(r); memcpy(out, r, 32); memcpy(state, r, 32); n_rounds++; }
  // :197-202 + mont_ark_u128.rs:51-63 — returns challenge limbs {0,0,lo,hi}
  void challenge_optimized(uint64_t out[4]) {
    uint8_t r[32]; challenge_bytes32(r);
    uint64_t lo, hi; memcpy(&lo, r, 8); memcpy(&hi, r + 8, 8);   // LE u128 of the first 16 bytes
    out[0] = 0; out[1] = 0; out[2] = lo; out[3] = hi & (0xffffffffffffffffull >> 3);
  }
  // :204-215 challenge_scalar: Fr::from_le_bytes_mod_order(reverse(first 16 bytes)) == big-endian integer < 2^128
  Fr challenge_scalar() {
    uint8_t r[32]; challenge_bytes32(r);
    uint64_t c[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; i++) c[i / 8] |= (uint64_t)r[15 - i] << (8 * (i % 8));
    return Fr::from_canonical(c);
  }
  std::vector<Fr> challenge_vector(size_t n) { std::vector<Fr> v(n); for (auto& x : v) x = challenge_scalar(); return v; }
  std::vector<Fr> challenge_scalar_powers(size_t n) {   // :224-231
    Fr q = challenge_scalar();
    std::vector<Fr> v(n, Fr::one());
    for (size_t i = 1; i < n; i++) v[i] = v[i - 1] * q;
    return v;
  }
};

// ---------------------------------------------------------------- UniPoly (unipoly.rs)
inline std::vector<Fr> gaussian_elimination(std::vector<std::vector<Fr>>& m) {   // gaussian_elimination.rs:9-70
  const size_t size = m.size();
  for (size_t i = 0; i + 1 < size; i++)
    for (size_t j = i; j + 1 < size; j++)
      if (!m[i][i].is_zero()) {
        Fr f = m[j + 1][i] * m[i][i].inv();
        for (size_t k = i; k < size + 1; k++) { Fr tmp = m[i][k]; m[j + 1][k] -= f * tmp; }
      }
  for (size_t i = size - 1; i >= 1; i--)
    if (!m[i][i].is_zero())
      for (size_t j = i; j >= 1; j--) {
        Fr f = m[j - 1][i] * m[i][i].inv();
        for (size_t k = size + 1; k-- > 0;) { Fr tmp = m[i][k]; m[j - 1][k] -= f * tmp; }
      }
  std::vector<Fr> res(size);
  for (size_t i = 0; i < size; i++) res[i] = m[i][size] * m[i][i].inv();
  return res;
}

struct UniPoly {
  std::vector<Fr> coeffs;
  static UniPoly from_coeff(std::vector<Fr> c) {     // :39-52
    while (!c.empty() && c.back().is_zero()) c.pop_back();
    if (c.empty()) c.push_back(Fr::zero());
    return UniPoly{c};
  }
  static UniPoly from_evals(const std::vector<Fr>& e) {   // :55-92, :136-153
    const size_t n = e.size();
    if (n == 3) {
      Fr two_inv = Fr::from_u64(2).inv();
      Fr c2 = (e[0] - e[1] - e[1] + e[2]) * two_inv;
      Fr c1 = e[1] - e[0] - c2;
      return UniPoly{{e[0], c1, c2}};
    }
    if (n == 4) {
      Fr two_inv = Fr::from_u64(2).inv(), six_inv = Fr::from_u64(6).inv();
      Fr c3 = (e[3] - e[0] + (e[1] - e[2]) * Fr::from_u64(3)) * six_inv;
      Fr c2 = (e[0] - e[1] - e[1] + e[2]) * two_inv - c3 - c3 - c3;
      Fr c1 = e[1] - e[0] - c2 - c3;
      return UniPoly{{e[0], c1, c2, c3}};
    }
    std::vector<std::vector<Fr>> rows(n);
    for (size_t i = 0; i < n; i++) {
      Fr x = Fr::from_u64(i), pw = Fr::one();
      for (size_t j = 0; j < n; j++) { rows[i].push_back(pw); pw *= x; }
      rows[i].push_back(e[i]);
    }
    return from_coeff(gaussian_elimination(rows));
  }
  static UniPoly from_evals_and_hint(const Fr& hint, const std::vector<Fr>& evals) {   // :96-101
    std::vector<Fr> e = evals;
    e.insert(e.begin() + 1, hint - e[0]);
    return from_evals(e);
  }
  static UniPoly from_evals_toom(const std::vector<Fr>& e) {   // :104-134 (no trimming)
    const size_t n = e.size();
    std::vector<std::vector<Fr>> rows(n);
    for (size_t i = 0; i + 1 < n; i++) {
      Fr x = Fr::from_u64(i), pw = Fr::one();
      for (size_t j = 0; j < n; j++) { rows[i].push_back(pw); pw *= x; }
      rows[i].push_back(e[i]);
    }
    rows[n - 1].assign(n - 1, Fr::zero());
    rows[n - 1].push_back(Fr::one());
    rows[n - 1].push_back(e[n - 1]);
    return UniPoly{gaussian_elimination(rows)};
  }
  Fr evaluate(const Fr& r) const {   // :219-245
    Fr acc = coeffs[0], pw = r;
    for (size_t i = 1; i < coeffs.size(); i++) { acc += pw * coeffs[i]; pw = pw * r; }
    return acc;
  }
  std::vector<Fr> compress() const {   // :307-318
    if (coeffs.size() < 2) return coeffs;
    std::vector<Fr> c; c.push_back(coeffs[0]);
    c.insert(c.end(), coeffs.begin() + 2, coeffs.end());
    return c;
  }
  UniPoly scaled(const Fr& s) const { std::vector<Fr> c = coeffs; for (auto& x : c) x *= s; return from_coeff(c); }   // :455-461
  void add_assign(const UniPoly& o) {   // :400-412
    for (size_t i = 0; i < coeffs.size() && i < o.coeffs.size(); i++) coeffs[i] += o.coeffs[i];
    if (coeffs.size() < o.coeffs.size()) coeffs.insert(coeffs.end(), o.coeffs.begin() + coeffs.size(), o.coeffs.end());
  }
};

inline void append_compressed(Transcript& t, const std::vector<Fr>& c) {   // :550-558
  t.append_message("UniPoly_begin");
  for (auto& x : c) t.append_scalar(x);
  t.append_message("UniPoly_end");
}

}  // namespace orc
