// lookup_host.h -- host side of K12: rsb200_lookup_job -> LookupJobDev (no CUDA here: shared
// with the CPU replay in tests/emu).
#pragma once

#include "../../include/rawspeed_b200.h"
#include "lookup_core.h"

namespace rsb200 {

inline const char* lookup_build_job(const rsb200_lookup_job& j, int ntables, uint32_t quad_begin,
                                    LookupJobDev* d) {
  if ((j.offset & 15) || (j.pitch & 15) || !j.width || !j.height || j.width > 65535 ||
      j.height > 65535 || j.cpp < 1 || j.cpp > 4 || (uint64_t)j.width * j.cpp * 2 > j.pitch)
    return "malformed image descriptor";
  if (j.table >= (uint32_t)ntables)
    return "table index outside the plan's tables";
  LookupJobDev o{};
  o.offset = j.offset;
  o.pitch = j.pitch;
  o.width = j.width;
  o.height = j.height;
  o.ncols = j.width * j.cpp;
  o.ngroups = (o.ncols + 7) / 8;
  o.table = j.table;
  o.quad_begin = quad_begin;
  *d = o;
  return nullptr;
}

inline uint32_t lookup_job_quads(const rsb200_lookup_job& j) { return (j.height + SCALE_ROWS - 1) / SCALE_ROWS; }

} // namespace rsb200
