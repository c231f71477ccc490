// badpix_emu.cpp -- CPU replay of badpix_kernel (rawspeed_b200/csrc/badpix.cuh): the bitmap /
// list builder (badpix_host.h) and the per-pixel interpolation (badpix_core.h) are the
// library's own source; the loop below is the kernel's grid.  Values are computed for every
// listed pixel BEFORE any is written when `two_phase` is set, and pixel by pixel otherwise:
// both must agree (only good pixels are read), which is what makes the kernel race-free.
#include "../../rawspeed_b200/csrc/badpix_host.h"

#include <cstring>

using namespace rsb200;

extern "C" int badpix_emu_run(uint8_t* img, const rsb200_badpix_job* jobs, int njobs,
                              const uint32_t* positions, uint32_t npositions, int two_phase,
                              char* err, int errlen) {
  std::vector<BadPixJobDev> hj((size_t)njobs);
  std::vector<uint8_t> maps;
  std::vector<uint32_t> list;
  for (int i = 0; i < njobs; ++i)
    if (const char* why = badpix_build(jobs[i], positions, npositions, jobs[i].prior_map, &hj[i], &maps,
                                       &list)) {
      std::strncpy(err, why, (size_t)errlen - 1);
      err[errlen - 1] = 0;
      return -1;
    }
  std::vector<uint32_t> vals(list.size());
  for (int pass = 0; pass < 2; ++pass)
    for (uint32_t i = 0; i < (uint32_t)list.size(); ++i) {
      int lo = 0, hi = njobs - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (hj[mid].first <= i)
          lo = mid;
        else
          hi = mid - 1;
      }
      const BadPixJobDev jb = hj[lo];
      const int x = (int)(list[i] & 0xFFFFu), y = (int)(list[i] >> 16);
      uint16_t* px = reinterpret_cast<uint16_t*>(img + jb.offset + (size_t)y * jb.pitch) + x;
      if (two_phase) {
        if (pass == 0)
          vals[i] = badpix_value(img, jb, maps.data() + jb.map_offset, x, y);
        else
          *px = (uint16_t)vals[i];
      } else if (pass == 0) {
        *px = (uint16_t)badpix_value(img, jb, maps.data() + jb.map_offset, x, y);
      }
    }
  return (int)list.size();
}
