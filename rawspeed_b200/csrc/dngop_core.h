// dngop_core.h -- the per-thread program of K10 (fused DNG opcode pass), written so that the
// same source compiles as device code (dngop.cuh) and as plain C++ (tests/emu/dngop_emu.cpp).
//
// Reference: DngOpcodes::PixelOpcode::applyOP and the opcodes built on it
//   common/DngOpcodes.cpp:390-409 (lattice walk), LookupOpcode :417-438, OffsetPerRowOrCol
//   :591-623, ScalePerRowOrCol :625-662, FixBadPixelsConstant::apply :172-183.
//
// The reference makes one pass over the image per opcode.  Every opcode here is a map of one
// sample (its result depends on the sample's own value and position only), and opcodes are
// applied in list order, so the list can be evaluated per sample in registers: one read and
// one write of the image for the whole list.  A thread owns eight consecutive samples of a
// row and walks the opcode list once.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define RSD_HD __host__ __device__ __forceinline__
#else
#define RSD_HD inline
#endif

namespace rsb200 {

enum : uint32_t {
  DNGOP_LOOKUP = 0,      // MapTable / MapPolynomial: v = table[v]
  DNGOP_OFFSET_ROW = 1,  // DeltaPerRow:    v = clamp16(delta[row index] + v)   | float: d + v
  DNGOP_OFFSET_COL = 2,  // DeltaPerColumn
  DNGOP_SCALE_ROW = 3,   // ScalePerRow:    v = clamp16((delta * v + 512) >> 10) | float: d * v
  DNGOP_SCALE_COL = 4,   // ScalePerColumn
  DNGOP_BAD_CONSTANT = 5 // FixBadPixelsConstant: report samples equal to `value`
};

struct DngOpDev {
  uint32_t kind;
  uint32_t top, left, bottom, right; // ROI in uncropped pixel coordinates
  uint32_t first_plane, planes;
  uint32_t row_pitch, col_pitch;
  uint32_t table; // LOOKUP: table index; OFFSET / SCALE: first element in `deltas`
  uint32_t value; // BAD_CONSTANT
  uint32_t slot;  // BAD_CONSTANT: index of its position list
};

struct DngOpJobDev {
  uint64_t offset;    // byte offset of row 0 of the uncropped image (multiple of 16)
  uint32_t pitch;     // bytes between rows (multiple of 16)
  uint32_t cpp;
  uint32_t is_f32;    // samples are floats (4 bytes) instead of uint16
  uint32_t row0;      // first / one-past-last row any opcode touches
  uint32_t row1;
  uint32_t groups;    // 8-sample groups per row = ceil(width * cpp / 8)
  uint32_t samples;   // width * cpp
  uint32_t first_op, num_ops;
  uint32_t unit_begin; // first (row, group) unit of this job
};

RSD_HD uint32_t dngop_clamp16(int32_t v) { return (uint32_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v)); }

// v[8]: samples s0 .. s0+7 of row r (uint16 values, or float bit patterns); Sink::hit(slot,
// row, col) receives FixBadPixelsConstant matches.
template <class Sink>
RSD_HD void dngop_apply_group(const DngOpDev* ops, uint32_t nops, const uint16_t* tables,
                              const uint32_t* deltas, const DngOpJobDev& jb, uint32_t r,
                              uint32_t s0, uint32_t (&v)[8], Sink& sink) {
  for (uint32_t k = 0; k < nops; ++k) {
    const DngOpDev op = ops[k];
    // row lattice: rows top, top + row_pitch, ... below bottom
    if (r < op.top || r >= op.bottom)
      continue;
    const uint32_t ry = r - op.top;
    const uint32_t yi = ry / op.row_pitch;
    if (yi * op.row_pitch != ry)
      continue;
    const bool by_row = op.kind == DNGOP_OFFSET_ROW || op.kind == DNGOP_SCALE_ROW;
    uint32_t col = s0 / jb.cpp, plane = s0 - col * jb.cpp;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int i = 0; i < 8; ++i) {
      const bool in_planes = plane - op.first_plane < op.planes; // (unsigned wrap: plane >= first)
      if (s0 + i < jb.samples && in_planes && col >= op.left && col < op.right) {
        const uint32_t cx = col - op.left;
        const uint32_t xi = op.col_pitch == 1 ? cx : cx / op.col_pitch;
        if (xi * op.col_pitch == cx) {
          const uint32_t sel = by_row ? yi : xi;
          switch (op.kind) {
          case DNGOP_LOOKUP:
            v[i] = tables[(size_t)op.table * 65536u + v[i]];
            break;
          case DNGOP_OFFSET_ROW:
          case DNGOP_OFFSET_COL:
            if (jb.is_f32) {
              union { uint32_t u; float f; } a, d;
              a.u = v[i];
              d.u = deltas[op.table + sel];
              a.f = d.f + a.f;
              v[i] = a.u;
            } else {
              v[i] = dngop_clamp16((int32_t)deltas[op.table + sel] + (int32_t)v[i]);
            }
            break;
          case DNGOP_SCALE_ROW:
          case DNGOP_SCALE_COL:
            if (jb.is_f32) {
              union { uint32_t u; float f; } a, d;
              a.u = v[i];
              d.u = deltas[op.table + sel];
              a.f = d.f * a.f;
              v[i] = a.u;
            } else {
              v[i] = dngop_clamp16(((int32_t)deltas[op.table + sel] * (int32_t)v[i] + 512) >> 10);
            }
            break;
          default: // DNGOP_BAD_CONSTANT (uint16, cpp == 1: checked by the plan)
            if (v[i] == op.value)
              sink.hit(op.slot, r, col);
            break;
          }
        }
      }
      if (++plane == jb.cpp) {
        plane = 0;
        ++col;
      }
    }
  }
}

} // namespace rsb200
