// phaseone_emu.cpp -- CPU replay of the third version of K8 (rawspeed_b200/csrc/phaseone.cuh:
// p1_walk_entry + p1_decode_entry), compiled by g++ against tests/emu/cuda_emu.h and run in the plan's
// order.  Test infrastructure (no GPU needed); parity of the real kernels is the GPU tests' job.
#include "cuda_emu.h"

#include "../../rawspeed_b200/csrc/phaseone.cuh"

#include <vector>

using namespace rsb200;

// strips: (offset, size, row) triples; returns the job's failure flag in *bad
extern "C" int p1_emu_run(const uint8_t* in, uint64_t in_total, const uint64_t* offs, const uint32_t* sizes,
                          const uint32_t* rows, int nstrips, int width, int out_pitch, uint8_t* out,
                          uint32_t* bad, int reverse, int first_form) {
  // the file at a 4-byte aligned address, exactly in_total bytes readable (+ the slack the ABI promises)
  // (128-byte aligned like a device allocation: the "lines" form of the walk reads whole aligned lines)
  std::vector<uint8_t> buf((size_t)in_total + 512);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(buf.data()) + 127) & ~(uintptr_t)127);
  memcpy(base, in, (size_t)in_total);
  std::vector<P1StripDev> st((size_t)nstrips);
  for (int i = 0; i < nstrips; ++i) {
    st[(size_t)i].in_offset = offs[i];
    st[(size_t)i].in_size = sizes[i];
    st[(size_t)i].row = rows[i];
    st[(size_t)i].job = 0;
    st[(size_t)i].pad = 0;
  }
  P1JobDev jb{0, (uint32_t)out_pitch, (uint32_t)width};
  const uint32_t gstride = ((uint32_t)width / 8u + 1u + 3u) & ~3u; // (rows of descriptors at multiples of 16 bytes)
  std::vector<uint32_t> gdesc_buf((size_t)gstride * (size_t)nstrips + 4, 0xCDCDCDCDu), rowflag((size_t)nstrips, 0xCDu);
  uint32_t* const gdesc_p = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(gdesc_buf.data()) + 15) & ~(uintptr_t)15);
  struct { uint32_t* p; uint32_t* data() const { return p; } } gdesc{gdesc_p};
  *bad = 0;
  const uint32_t nbw = ((uint32_t)nstrips + P1W_NT - 1) / P1W_NT;
  for (uint32_t b = 0; b < nbw; ++b)
    cuemu::run_cta(b, nbw, P1W_NT, sizeof(P1WalkShared), reverse != 0, [&](uint8_t* smem) {
      P1WalkShared& wsh = *reinterpret_cast<P1WalkShared*>(smem);
      if (first_form == 1)
        p1_walk_entry<false, 0>(wsh, base, st.data(), (uint32_t)nstrips, &jb, gstride, gdesc.data(), rowflag.data());
      else if (first_form == 2)
        p1_walk_entry<true, 0>(wsh, base, st.data(), (uint32_t)nstrips, &jb, gstride, gdesc.data(), rowflag.data());
      else if (first_form == 3)
        p1_walk_entry<true, 2>(wsh, base, st.data(), (uint32_t)nstrips, &jb, gstride, gdesc.data(), rowflag.data());
      else if (first_form == 5)
        p1_walk_entry<true, 1>(wsh, base, st.data(), (uint32_t)nstrips, &jb, gstride, gdesc.data(), rowflag.data());
      else if (first_form == 8)
        p1_walk_entry<true, 6>(wsh, base, st.data(), (uint32_t)nstrips, &jb, gstride, gdesc.data(), rowflag.data());
      else if (first_form == 6)
        p1_walk_entry<true, 4>(wsh, base, st.data(), (uint32_t)nstrips, &jb, gstride, gdesc.data(), rowflag.data());
      else
        p1_walk_entry<true, 5>(wsh, base, st.data(), (uint32_t)nstrips, &jb, gstride, gdesc.data(), rowflag.data());
    });
  const uint32_t nbd = ((uint32_t)nstrips * 32u + P1D_NT - 1) / P1D_NT;
  for (uint32_t b = 0; b < nbd; ++b)
    cuemu::run_cta(b, nbd, P1D_NT, sizeof(P1DecodeShared), reverse != 0, [&](uint8_t* smem) {
      p1_decode_entry(*reinterpret_cast<P1DecodeShared*>(smem), base, out, st.data(), (uint32_t)nstrips, &jb,
                      gstride, gdesc.data(), rowflag.data(), bad);
    });
  return 0;
}
