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

// n 32-bit outputs continuing from g (g is advanced).  temper = false stores the UNTEMPERED state words (the caller applies
// the tempering - mt_stream.cu does it on the device, which halves the host work per draw)
void mt_fill(MtPos& g, uint32_t* out, size_t n, bool temper = true);

}  // namespace mb
