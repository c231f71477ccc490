// dngop_emu.cpp -- CPU replay of dngop_kernel (rawspeed_b200/csrc/dngop.cuh): the per-thread
// opcode walk (dngop_core.h) and the descriptor builder (dngop_host.h) are the library's own
// source; the loop below mirrors the kernel's thread program (load eight samples, walk the
// list, write back what changed).  Test infrastructure: the parity of the real kernel is the
// GPU test's job (tests/test_gpu_dngopcodes.py).
#include "../../rawspeed_b200/csrc/dngop_host.h"

#include <cstring>
#include <vector>

using namespace rsb200;

namespace {
struct SinkHost {
  std::vector<std::vector<uint32_t>>* lists;
  void hit(uint32_t slot, uint32_t row, uint32_t col) { (*lists)[slot].push_back((row << 16) | col); }
};
} // namespace

// bad_out: for opcode k with a list, its positions are appended as [k, count, positions...];
// returns the number of 32-bit words written there (or -1 and a message)
extern "C" int dngop_emu_run(uint8_t* img, const rsb200_dngop_job* jobs, int njobs,
                             const rsb200_dng_op* ops, int nops, const uint16_t* tables, int ntables,
                             const uint32_t* deltas, int ndeltas, uint32_t* bad_out, uint32_t bad_cap,
                             char* err, int errlen) {
  std::vector<DngOpJobDev> hj((size_t)njobs);
  std::vector<DngOpDev> ho((size_t)nops + 1);
  std::vector<int> slot_of((size_t)nops + 1, -1);
  int nslots = 0;
  uint64_t units = 0;
  if (const char* why = dngop_build(jobs, njobs, ops, nops, ntables, ndeltas, hj.data(), ho.data(),
                                    slot_of.data(), &nslots, &units)) {
    std::strncpy(err, why, (size_t)errlen - 1);
    err[errlen - 1] = 0;
    return -1;
  }
  std::vector<std::vector<uint32_t>> lists((size_t)nslots);
  SinkHost sink{&lists};
  for (uint32_t u = 0; u < (uint32_t)units; ++u) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (hj[mid].unit_begin <= u)
        lo = mid;
      else
        hi = mid - 1;
    }
    const DngOpJobDev jb = hj[lo];
    const uint32_t ul = u - jb.unit_begin;
    const uint32_t r = jb.row0 + ul / jb.groups, g = ul % jb.groups;
    const uint32_t s0 = g * 8u;
    uint8_t* const rowp = img + jb.offset + (uint64_t)r * jb.pitch;
    uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0}, old[8];
    if (jb.is_f32) {
      std::memcpy(&v[0], rowp + (uint64_t)s0 * 4u, 16);
      if (s0 + 4u < jb.samples)
        std::memcpy(&v[4], rowp + (uint64_t)s0 * 4u + 16u, 16);
    } else {
      uint16_t t[8];
      std::memcpy(t, rowp + (uint64_t)s0 * 2u, 16);
      for (int i = 0; i < 8; ++i)
        v[i] = t[i];
    }
    std::memcpy(old, v, sizeof v);
    dngop_apply_group(ho.data() + jb.first_op, jb.num_ops, tables, deltas, jb, r, s0, v, sink);
    const bool lo_changed = std::memcmp(v, old, 16) != 0, hi_changed = std::memcmp(v + 4, old + 4, 16) != 0;
    if (jb.is_f32) {
      if (lo_changed)
        std::memcpy(rowp + (uint64_t)s0 * 4u, &v[0], 16);
      if (hi_changed)
        std::memcpy(rowp + (uint64_t)s0 * 4u + 16u, &v[4], 16);
    } else if (lo_changed || hi_changed) {
      uint16_t t[8];
      for (int i = 0; i < 8; ++i)
        t[i] = (uint16_t)v[i];
      std::memcpy(rowp + (uint64_t)s0 * 2u, t, 16);
    }
  }
  uint32_t w = 0;
  for (int k = 0; k < nops; ++k) {
    if (slot_of[k] < 0)
      continue;
    const auto& l = lists[(size_t)slot_of[k]];
    if ((uint64_t)w + 2 + l.size() > bad_cap) {
      std::strncpy(err, "bad_out too small", (size_t)errlen - 1);
      return -1;
    }
    bad_out[w++] = (uint32_t)k;
    bad_out[w++] = (uint32_t)l.size();
    for (uint32_t p : l)
      bad_out[w++] = p;
  }
  return (int)w;
}
