// pana4_emu.cpp -- CPU replay of pana_kernel<4, 12> (rawspeed_b200/csrc/pana.cuh): the packet
// arithmetic and the section-swap addressing are the kernel's own source (pana4_core.h); the
// loop below mirrors the kernel's per-thread program (load 16 bytes through the swap, decode,
// report zeros, store 14 pixels at linear pixel index 14 * unit).  Test infrastructure: the
// parity of the real kernel is the GPU test's job (tests/test_gpu_panasonic.py, V4 cases).
#include "../../rawspeed_b200/csrc/pana4_core.h"

#include <cstring>

using namespace rsb200;

extern "C" int pana4_emu_run(const uint8_t* in, uint64_t in_offset, uint8_t* out, uint64_t out_offset,
                             uint32_t out_pitch, uint32_t width, uint32_t height, uint32_t split,
                             int zero_is_not_bad, uint32_t* zero_list, uint32_t zero_cap,
                             uint32_t* zero_count) {
  const uint64_t area = (uint64_t)width * height;
  const uint32_t units = (uint32_t)(area / 14);
  const uint8_t* base = in + in_offset;
  uint32_t nz = 0;
  for (uint32_t ul = 0; ul < units; ++ul) {
    const uint32_t blk = ul >> 10, o = (ul & 1023u) * 16u;
    uint32_t w[4] = {0, 0, 0, 0};
    if ((split & 7u) == 0) {
      std::memcpy(&w[0], base + pana4_src(blk, o, split), 8);
      std::memcpy(&w[2], base + pana4_src(blk, o + 8u, split), 8);
    } else {
      for (uint32_t i = 0; i < 16; ++i)
        w[i >> 2] |= (uint32_t)base[pana4_src(blk, o + i, split)] << (8u * (i & 3u));
    }
    uint32_t px[14];
    const uint32_t zeros = pana4_packet(w, px);
    const uint32_t idx0 = ul * 14u;
    const uint32_t row = idx0 / width, col0 = idx0 - row * width;
    if (zeros && !zero_is_not_bad) {
      for (uint32_t z = zeros; z; z &= z - 1u) {
        const uint32_t i = (uint32_t)__builtin_ctz(z);
        if (nz < zero_cap)
          zero_list[nz] = (row << 16) | (col0 + i);
        ++nz;
      }
    }
    uint16_t* o16 = reinterpret_cast<uint16_t*>(out + out_offset + (uint64_t)row * out_pitch) + col0;
    for (int i = 0; i < 14; ++i)
      o16[i] = (uint16_t)px[i];
  }
  *zero_count = nz;
  return 0;
}
