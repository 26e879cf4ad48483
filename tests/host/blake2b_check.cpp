// CPU check of the product transcript's Blake2b-256 (csrc/transcript_host.hpp: unrolled scalar rounds and the AVX2 rows, whichever the CPU
// selects, plus the forced scalar path): prints "<len> <msg hex> <digest hex>" lines that tests/test_host_glue.py compares with hashlib.
#include <cstdio>
#include <random>
#include "fr_host.hpp"
#include "transcript_host.hpp"
using namespace ja; using namespace ja::host;
int main(int argc, char** argv) {
  const bool scalar = argc > 1;      // any argument: call the scalar compression directly
  // digests of messages of several lengths, hex, for comparison with hashlib.blake2b(digest_size=32)
  std::mt19937_64 rng(5);
  for (size_t len : {0, 1, 32, 64, 96, 127, 128, 129, 255, 256, 300}) {
    std::vector<uint8_t> msg(len);
    for (auto& b : msg) b = (uint8_t)rng();
    uint8_t out[32];
    if (!scalar) b2::blake2b_256(msg.data(), len, out);
    else {                                   // RFC 7693 with the scalar compression only
      uint64_t h[8];
      for (int i = 0; i < 8; i++) h[i] = b2::kIV[i];
      h[0] ^= 0x01010020ull;
      size_t done = 0;
      uint8_t block[128];
      while (len - done > 128) { b2::compress(h, msg.data() + done, done + 128, false); done += 128; }
      memset(block, 0, 128);
      if (len - done) memcpy(block, msg.data() + done, len - done);
      b2::compress(h, block, len, true);
      for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(h[i] >> (8 * k));
    }
    printf("%zu ", len);
    for (auto b : msg) printf("%02x", b);
    printf(" ");
    for (int i = 0; i < 32; i++) printf("%02x", out[i]);
    printf("\n");
  }
  // be32
  uint64_t c[4] = {0x0807060504030201ull, 0x100f0e0d0c0b0a09ull, 0x1817161514131211ull, 0x201f1e1d1c1b1a19ull};
  uint8_t o[32]; be32(c, o);
  printf("be32 "); for (int i = 0; i < 32; i++) printf("%02x", o[i]); printf("\n");
}
