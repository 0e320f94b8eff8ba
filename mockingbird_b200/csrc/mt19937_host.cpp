// MT19937 block generator (host only; compiled by g++ so that the twist / tempering loops get AVX2 / AVX-512 clones).
// Bit-compatible with std::mt19937 / at::mt19937 (aten/src/ATen/core/MT19937RNGEngine.h): same recurrence, same tempering.
#include "mt19937_host.h"

#include <cstring>

namespace mb {
namespace {
constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

// one block twist: s[0..623] <- next 624 state words (std::mt19937 / at::mt19937::next_state)
__attribute__((target_clones("avx512f", "avx2", "default"))) void mt_twist(uint32_t* __restrict__ s) {
  // i in [0, 227): reads old s[i+1], old s[i+397]
  for (int i = 0; i < kN - kM; ++i) {
    const uint32_t y = (s[i] & kUpper) | (s[i + 1] & kLower);
    s[i] = s[i + kM] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
  }
  // i in [227, 623): reads old s[i+1], NEW s[i-227]; chunks of <= 227 are independent
  for (int base = kN - kM; base < kN - 1; base += kN - kM) {
    const int end = base + (kN - kM) < kN - 1 ? base + (kN - kM) : kN - 1;
    for (int i = base; i < end; ++i) {
      const uint32_t y = (s[i] & kUpper) | (s[i + 1] & kLower);
      s[i] = s[i - (kN - kM)] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
    }
  }
  const uint32_t y = (s[kN - 1] & kUpper) | (s[0] & kLower);
  s[kN - 1] = s[kM - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
}

__attribute__((target_clones("avx512f", "avx2", "default"))) void mt_temper(const uint32_t* __restrict__ s,
                                                                             uint32_t* __restrict__ out, int n) {
  for (int i = 0; i < n; ++i) {
    uint32_t y = s[i];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    out[i] = y;
  }
}

}  // namespace

void mt_fill(MtPos& g, uint32_t* out, size_t n, bool temper) {
  while (n > 0) {
    if (g.avail == 0) {
      mt_twist(g.s);
      g.idx = 0;
      g.avail = kN;
    }
    const int m = (size_t)g.avail < n ? g.avail : (int)n;
    if (temper) mt_temper(g.s + g.idx, out, m);
    else memcpy(out, g.s + g.idx, sizeof(uint32_t) * (size_t)m);
    g.idx += m;
    g.avail -= m;
    out += m;
    n -= (size_t)m;
  }
}


}  // namespace mb
