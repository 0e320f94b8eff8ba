// host MT19937 stream position + filler (mt19937_host.cpp)
#pragma once
#include <cstddef>
#include <cstdint>

namespace mb {

struct MtPos {
  uint32_t s[624];
  int idx;    // next output index in the current block
  int avail;  // outputs left in the current block (at::mt19937: left_ - 1)
};

// n tempered 32-bit outputs continuing from g (g is advanced)
void mt_fill(MtPos& g, uint32_t* out, size_t n);

}  // namespace mb
