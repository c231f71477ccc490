// dngop_host.h -- host side of K10: rsb200_dngop_job / rsb200_dng_op -> DngOpJobDev / DngOpDev
// (no CUDA here, so the CPU replay in tests/emu builds its descriptors with the code the
// library uses).
#pragma once

#include "../../include/rawspeed_b200.h"
#include "dngop_core.h"

#include <algorithm>
#include <vector>

namespace rsb200 {

// dev_jobs[njobs], dev_ops[nops], slot_of_op[nops] (-1 = no list) are filled; *nslots and *units
// accumulate.  Returns nullptr, or why a descriptor is refused.
inline const char* dngop_build(const rsb200_dngop_job* jobs, int njobs, const rsb200_dng_op* ops,
                               int nops, int ntables, int ndeltas, DngOpJobDev* dev_jobs,
                               DngOpDev* dev_ops, int* slot_of_op, int* nslots, uint64_t* units) {
  std::vector<char> used((size_t)nops, 0);
  for (int i = 0; i < njobs; ++i) {
    const rsb200_dngop_job& j = jobs[i];
    if ((j.offset & 15) || (j.pitch & 15) || !j.width || !j.height || j.width > 65535 ||
        j.height > 65535 || j.cpp < 1 || j.cpp > 4 ||
        (uint64_t)j.width * j.cpp * (j.is_f32 ? 4u : 2u) > j.pitch ||
        (uint64_t)j.first_op + j.num_ops > (uint64_t)nops)
      return "malformed image descriptor";
    DngOpJobDev d{};
    d.offset = j.offset;
    d.pitch = j.pitch;
    d.cpp = j.cpp;
    d.is_f32 = j.is_f32 ? 1u : 0u;
    d.samples = j.width * j.cpp;
    d.groups = (d.samples + 7) / 8;
    d.first_op = j.first_op;
    d.num_ops = j.num_ops;
    uint32_t row0 = j.height, row1 = 0;
    for (uint32_t k = j.first_op; k < j.first_op + j.num_ops; ++k) {
      const rsb200_dng_op& o = ops[k];
      const bool delta = o.kind >= RSB200_DNGOP_OFFSET_ROW && o.kind <= RSB200_DNGOP_SCALE_COL;
      bool ok = o.kind <= RSB200_DNGOP_BAD_CONSTANT && !used[k] && o.top <= o.bottom &&
                o.left <= o.right && o.bottom <= j.height && o.right <= j.width && o.planes >= 1 &&
                (uint64_t)o.first_plane + o.planes <= j.cpp && o.row_pitch >= 1 && o.col_pitch >= 1;
      if (ok && o.kind == RSB200_DNGOP_LOOKUP)
        ok = !j.is_f32 && o.table < (uint32_t)ntables;
      if (ok && o.kind == RSB200_DNGOP_BAD_CONSTANT)
        ok = !j.is_f32 && j.cpp == 1;
      if (ok && delta) {
        const bool by_row = o.kind == RSB200_DNGOP_OFFSET_ROW || o.kind == RSB200_DNGOP_SCALE_ROW;
        const uint64_t span = by_row ? o.bottom - o.top : o.right - o.left;
        const uint64_t pitch = by_row ? o.row_pitch : o.col_pitch;
        const uint64_t need = span ? 1 + (span - 1) / pitch : 0;
        ok = (uint64_t)o.table + need <= (uint64_t)ndeltas;
      }
      if (!ok)
        return "malformed opcode (ROI / planes / pitch / table outside the image or the arrays, "
               "or an opcode shared by two images)";
      used[k] = 1;
      DngOpDev& h = dev_ops[k];
      h.kind = o.kind;
      h.top = o.top;
      h.left = o.left;
      h.bottom = o.bottom;
      h.right = o.right;
      h.first_plane = o.first_plane;
      h.planes = o.planes;
      h.row_pitch = o.row_pitch;
      h.col_pitch = o.col_pitch;
      h.table = o.table;
      h.value = o.value;
      h.slot = 0;
      slot_of_op[k] = -1;
      if (o.kind == RSB200_DNGOP_BAD_CONSTANT) {
        slot_of_op[k] = *nslots;
        h.slot = (uint32_t)(*nslots)++;
      }
      if (o.top < o.bottom && o.left < o.right) {
        row0 = std::min(row0, o.top);
        row1 = std::max(row1, o.bottom);
      }
    }
    if (row1 <= row0)
      row0 = row1 = 0; // nothing to do for this image
    d.row0 = row0;
    d.row1 = row1;
    d.unit_begin = (uint32_t)*units;
    *units += (uint64_t)(row1 - row0) * d.groups;
    if (*units > 0x7FFFFFFFull)
      return "too many samples";
    dev_jobs[i] = d;
  }
  return nullptr;
}

} // namespace rsb200
