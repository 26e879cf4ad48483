// Host-side Fiat–Shamir transcript of the product library (the C++ host driver and the all-in-one
// HyperKZG / sumcheck entry points run it between kernel launches; a Rust caller keeps using its own
// Blake2bTranscript and the split entry points instead).
//
// Byte rules follow joltworks/src/transcripts/blake2b.rs:
//   new(label)            state = H(label || 0-pad to 32)                                    :81-100
//   every operation       H(state || 0^28 || n_rounds_be32 || payload), then n_rounds += 1   :31-37
//   append_message        payload = msg right-padded to 32 bytes                             :109-122
//   append_scalar         payload = 32-byte big-endian canonical integer                     :138-146
//   append_u64            payload = 24 zero bytes || big-endian u64                          :130-136
//   append_point          payload = x_be || y_be (64 zero bytes for infinity)                :166-187
//   append_scalars/points wrapped in "begin_append_vector" / "end_append_vector"             :158-164,:189-195
//   challenge_u128        little-endian u128 of the first 16 digest bytes                    :197-202
//   challenge_scalar      big-endian integer of the first 16 digest bytes, as Fr             :204-215
//   challenge_scalar_optimized   challenge_u128 & (u128::MAX >> 3) as Montgomery limbs {0,0,lo,hi}   :233-238,
//                                field/challenge/mont_ark_u128.rs:51-63
// Blake2b-256 is RFC 7693 (unkeyed, 32-byte digest), the `blake2 0.10.6` crate in the reference.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include "fr_host.hpp"

namespace ja {
namespace host {

namespace b2 {
static const uint64_t kIV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
static constexpr uint8_t kSigma[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

static inline uint64_t ror(uint64_t x, unsigned n) { return (x >> n) | (x << (64 - n)); }

// One round with the message schedule resolved at compile time (R is a template parameter, so every m[sigma] is a fixed
// register / stack slot): the transcript hashes ~20 blocks per sumcheck round on the Fiat-Shamir critical path, and a
// table-driven round loop cost 2.5x as much.
#define JA_B2_G(a, b, c, d, x, y)                                  \
  v[a] += v[b] + (x); v[d] = ror(v[d] ^ v[a], 32); v[c] += v[d]; v[b] = ror(v[b] ^ v[c], 24); \
  v[a] += v[b] + (y); v[d] = ror(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = ror(v[b] ^ v[c], 63);
template <int R>
static inline __attribute__((always_inline)) void round_fn(uint64_t (&v)[16], const uint64_t (&m)[16]) {
  constexpr int S = R % 10;
  JA_B2_G(0, 4, 8, 12, m[kSigma[S][0]], m[kSigma[S][1]])   JA_B2_G(1, 5, 9, 13, m[kSigma[S][2]], m[kSigma[S][3]])
  JA_B2_G(2, 6, 10, 14, m[kSigma[S][4]], m[kSigma[S][5]])  JA_B2_G(3, 7, 11, 15, m[kSigma[S][6]], m[kSigma[S][7]])
  JA_B2_G(0, 5, 10, 15, m[kSigma[S][8]], m[kSigma[S][9]])  JA_B2_G(1, 6, 11, 12, m[kSigma[S][10]], m[kSigma[S][11]])
  JA_B2_G(2, 7, 8, 13, m[kSigma[S][12]], m[kSigma[S][13]]) JA_B2_G(3, 4, 9, 14, m[kSigma[S][14]], m[kSigma[S][15]])
}
#undef JA_B2_G

// one compression of a 128-byte block; `bytes_so_far` counts the message bytes up to and including this block
static inline void compress(uint64_t h[8], const uint8_t block[128], uint64_t bytes_so_far, bool final_block) {
  uint64_t m[16], v[16];
#if defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__
  memcpy(m, block, 128);
#else
  for (int i = 0; i < 16; i++) {
    uint64_t w = 0;
    for (int k = 7; k >= 0; k--) w = (w << 8) | block[8 * i + k];
    m[i] = w;
  }
#endif
  for (int i = 0; i < 8; i++) { v[i] = h[i]; v[8 + i] = kIV[i]; }
  v[12] ^= bytes_so_far;
  if (final_block) v[14] = ~v[14];
  round_fn<0>(v, m); round_fn<1>(v, m); round_fn<2>(v, m); round_fn<3>(v, m); round_fn<4>(v, m); round_fn<5>(v, m);
  round_fn<6>(v, m); round_fn<7>(v, m); round_fn<8>(v, m); round_fn<9>(v, m); round_fn<10>(v, m); round_fn<11>(v, m);
  for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
}

#if defined(__x86_64__)
// AVX2 compression (rows of the 4 x 4 state in one 256-bit register each, as in the BLAKE2 reference's blake2b-round.h), chosen at
// run time when the CPU has it: ~3x the scalar code, and the transcript is ~20 compressions per sumcheck round.
namespace avx2 {
#define JA_B2_TGT __attribute__((target("avx2"), always_inline)) static inline
JA_B2_TGT __m256i rot32(__m256i x) { return _mm256_shuffle_epi32(x, _MM_SHUFFLE(2, 3, 0, 1)); }
JA_B2_TGT __m256i rot24(__m256i x) {
  return _mm256_shuffle_epi8(x, _mm256_setr_epi8(3, 4, 5, 6, 7, 0, 1, 2, 11, 12, 13, 14, 15, 8, 9, 10, 3, 4, 5, 6, 7, 0, 1, 2, 11, 12, 13, 14, 15, 8, 9, 10));
}
JA_B2_TGT __m256i rot16(__m256i x) {
  return _mm256_shuffle_epi8(x, _mm256_setr_epi8(2, 3, 4, 5, 6, 7, 0, 1, 10, 11, 12, 13, 14, 15, 8, 9, 2, 3, 4, 5, 6, 7, 0, 1, 10, 11, 12, 13, 14, 15, 8, 9));
}
JA_B2_TGT __m256i rot63(__m256i x) { return _mm256_or_si256(_mm256_srli_epi64(x, 63), _mm256_add_epi64(x, x)); }
template <int R>
JA_B2_TGT void round_fn(__m256i& a, __m256i& b, __m256i& c, __m256i& d, const uint64_t (&m)[16]) {
  constexpr int S = R % 10;
#define JA_M(i) (long long)m[kSigma[S][i]]
  // column step: G(0,4,8,12) G(1,5,9,13) G(2,6,10,14) G(3,7,11,15) in the four lanes
  a = _mm256_add_epi64(_mm256_add_epi64(a, _mm256_set_epi64x(JA_M(6), JA_M(4), JA_M(2), JA_M(0))), b);
  d = rot32(_mm256_xor_si256(d, a)); c = _mm256_add_epi64(c, d); b = rot24(_mm256_xor_si256(b, c));
  a = _mm256_add_epi64(_mm256_add_epi64(a, _mm256_set_epi64x(JA_M(7), JA_M(5), JA_M(3), JA_M(1))), b);
  d = rot16(_mm256_xor_si256(d, a)); c = _mm256_add_epi64(c, d); b = rot63(_mm256_xor_si256(b, c));
  // diagonalise: lane i of b <- lane i+1, c <- lane i+2, d <- lane i+3
  b = _mm256_permute4x64_epi64(b, _MM_SHUFFLE(0, 3, 2, 1));
  c = _mm256_permute4x64_epi64(c, _MM_SHUFFLE(1, 0, 3, 2));
  d = _mm256_permute4x64_epi64(d, _MM_SHUFFLE(2, 1, 0, 3));
  // diagonal step: G(0,5,10,15) G(1,6,11,12) G(2,7,8,13) G(3,4,9,14)
  a = _mm256_add_epi64(_mm256_add_epi64(a, _mm256_set_epi64x(JA_M(14), JA_M(12), JA_M(10), JA_M(8))), b);
  d = rot32(_mm256_xor_si256(d, a)); c = _mm256_add_epi64(c, d); b = rot24(_mm256_xor_si256(b, c));
  a = _mm256_add_epi64(_mm256_add_epi64(a, _mm256_set_epi64x(JA_M(15), JA_M(13), JA_M(11), JA_M(9))), b);
  d = rot16(_mm256_xor_si256(d, a)); c = _mm256_add_epi64(c, d); b = rot63(_mm256_xor_si256(b, c));
  b = _mm256_permute4x64_epi64(b, _MM_SHUFFLE(2, 1, 0, 3));
  c = _mm256_permute4x64_epi64(c, _MM_SHUFFLE(1, 0, 3, 2));
  d = _mm256_permute4x64_epi64(d, _MM_SHUFFLE(0, 3, 2, 1));
#undef JA_M
}
__attribute__((target("avx2"))) static inline void compress(uint64_t h[8], const uint8_t block[128], uint64_t bytes_so_far, bool final_block) {
  uint64_t m[16];
  memcpy(m, block, 128);
  __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(h));
  __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(h + 4));
  __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(kIV));
  __m256i d = _mm256_xor_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(kIV + 4)),
                               _mm256_set_epi64x(0, final_block ? -1ll : 0ll, 0, (long long)bytes_so_far));
  const __m256i a0 = a, b0 = b;
  round_fn<0>(a, b, c, d, m); round_fn<1>(a, b, c, d, m); round_fn<2>(a, b, c, d, m); round_fn<3>(a, b, c, d, m);
  round_fn<4>(a, b, c, d, m); round_fn<5>(a, b, c, d, m); round_fn<6>(a, b, c, d, m); round_fn<7>(a, b, c, d, m);
  round_fn<8>(a, b, c, d, m); round_fn<9>(a, b, c, d, m); round_fn<10>(a, b, c, d, m); round_fn<11>(a, b, c, d, m);
  _mm256_storeu_si256(reinterpret_cast<__m256i*>(h), _mm256_xor_si256(a0, _mm256_xor_si256(a, c)));
  _mm256_storeu_si256(reinterpret_cast<__m256i*>(h + 4), _mm256_xor_si256(b0, _mm256_xor_si256(b, d)));
}
#undef JA_B2_TGT
}  // namespace avx2
static inline bool have_avx2() { static const bool v = __builtin_cpu_supports("avx2"); return v; }
#endif

static inline void compress_any(uint64_t h[8], const uint8_t block[128], uint64_t bytes_so_far, bool final_block) {
#if defined(__x86_64__)
  if (have_avx2()) { avx2::compress(h, block, bytes_so_far, final_block); return; }
#endif
  compress(h, block, bytes_so_far, final_block);
}

// digest of a whole message held in memory
static inline void blake2b_256(const uint8_t* msg, size_t len, uint8_t out[32]) {
  uint64_t h[8];
  for (int i = 0; i < 8; i++) h[i] = kIV[i];
  h[0] ^= 0x01010020ull;   // digest length 32, no key, fanout = depth = 1
  size_t done = 0;
  uint8_t block[128];
  while (len - done > 128) {
    compress_any(h, msg + done, done + 128, false);
    done += 128;
  }
  const size_t rest = len - done;
  memset(block, 0, 128);
  if (rest) memcpy(block, msg + done, rest);
  compress_any(h, block, len, true);
  for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(h[i] >> (8 * k));
}
}  // namespace b2

static const uint64_t FQ_P[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t FQ_INV = 0x87d20782e4866389ull;

// Montgomery limbs -> canonical integer limbs for an arbitrary 4-limb modulus (one REDC pass)
static inline void mont_to_canonical(const uint64_t a[4], const uint64_t p[4], uint64_t inv, uint64_t out[4]) {
  uint64_t t[5] = {a[0], a[1], a[2], a[3], 0};
  for (int i = 0; i < 4; i++) {
    const uint64_t m = t[0] * inv;
    u128 c = (u128)m * p[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = (uint64_t)(c >> 64);
  }
  bool ge = t[4] != 0;
  if (!ge) { ge = true; for (int i = 3; i >= 0; i--) { if (t[i] > p[i]) break; if (t[i] < p[i]) { ge = false; break; } } }
  if (ge) { u128 b = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - p[i] - (uint64_t)b; t[i] = (uint64_t)d; b = (d >> 64) & 1; } }
  memcpy(out, t, 32);
}
static inline void be32(const uint64_t canon[4], uint8_t out[32]) {
  for (int i = 0; i < 4; i++) { const uint64_t w = __builtin_bswap64(canon[3 - i]); memcpy(out + 8 * i, &w, 8); }
}

class Blake2bTranscript {
 public:
  uint8_t state[32];
  uint32_t n_rounds = 0;

  explicit Blake2bTranscript(const char* label) {
    uint8_t pad[32] = {0};
    size_t n = strlen(label);
    memcpy(pad, label, n > 32 ? 32 : n);
    b2::blake2b_256(pad, 32, state);
  }
  Blake2bTranscript(const uint8_t st[32], uint32_t rounds) : n_rounds(rounds) { memcpy(state, st, 32); }

  void append_message(const char* msg) {
    uint8_t pad[32] = {0};
    size_t n = strlen(msg);
    memcpy(pad, msg, n > 32 ? 32 : n);
    absorb(pad, 32);
  }
  void append_bytes(const uint8_t* p, size_t n) { absorb(p, n); }
  void append_u64(uint64_t x) {
    uint8_t b[32] = {0};
    for (int i = 0; i < 8; i++) b[31 - i] = (uint8_t)(x >> (8 * i));
    absorb(b, 32);
  }
  void append_scalar(const FrH& x) {
    uint64_t c[4]; to_canonical(x, c);
    uint8_t b[32]; be32(c, b);
    absorb(b, 32);
  }
  void append_scalars(const FrH* xs, size_t n) {
    append_message("begin_append_vector");
    for (size_t i = 0; i < n; i++) append_scalar(xs[i]);
    append_message("end_append_vector");
  }
  // affine G1 point as Montgomery Fq limbs x||y
  void append_point(const uint64_t xy[8], bool is_inf) {
    uint8_t b[64] = {0};
    if (!is_inf) {
      uint64_t cx[4], cy[4];
      mont_to_canonical(xy, FQ_P, FQ_INV, cx);
      mont_to_canonical(xy + 4, FQ_P, FQ_INV, cy);
      be32(cx, b); be32(cy, b + 32);
    }
    absorb(b, 64);
  }
  void append_points(const uint64_t* xy, const int32_t* is_inf, size_t n) {
    append_message("begin_append_vector");
    for (size_t i = 0; i < n; i++) append_point(xy + 8 * i, is_inf && is_inf[i]);
    append_message("end_append_vector");
  }
  // 125-bit challenge as Montgomery limbs {0, 0, lo, hi}
  void challenge_scalar_optimized(uint64_t out[4]) {
    uint8_t d[32]; squeeze(d);
    uint64_t lo = 0, hi = 0;
    for (int i = 7; i >= 0; i--) { lo = (lo << 8) | d[i]; hi = (hi << 8) | d[8 + i]; }
    out[0] = 0; out[1] = 0; out[2] = lo; out[3] = hi & (~0ull >> 3);
  }
  FrH challenge_scalar() {
    uint8_t d[32]; squeeze(d);
    uint64_t c[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; i++) c[i / 8] |= (uint64_t)d[15 - i] << (8 * (i % 8));
    return from_canonical(c);
  }
  std::vector<FrH> challenge_vector(size_t n) {
    std::vector<FrH> v(n);
    for (auto& x : v) x = challenge_scalar();
    return v;
  }
  std::vector<FrH> challenge_scalar_powers(size_t n) {
    const FrH q = challenge_scalar();
    std::vector<FrH> v(n, FR_ONE);
    for (size_t i = 1; i < n; i++) v[i] = mul(v[i - 1], q);
    return v;
  }

 private:
  void absorb(const uint8_t* payload, size_t n) {
    uint8_t small[128];                                   // state || 0^28 || round || payload: one block for every scalar / label
    std::vector<uint8_t> big;
    uint8_t* msg = small;
    if (64 + n > sizeof(small)) { big.assign(64 + n, 0); msg = big.data(); }
    memcpy(msg, state, 32);
    memset(msg + 32, 0, 28);
    msg[60] = (uint8_t)(n_rounds >> 24); msg[61] = (uint8_t)(n_rounds >> 16);
    msg[62] = (uint8_t)(n_rounds >> 8); msg[63] = (uint8_t)n_rounds;
    if (n) memcpy(msg + 64, payload, n);
    b2::blake2b_256(msg, 64 + n, state);
    n_rounds++;
  }
  void squeeze(uint8_t out[32]) {
    absorb(nullptr, 0);
    memcpy(out, state, 32);
  }
};

}  // namespace host
}  // namespace ja
