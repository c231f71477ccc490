// scale_host.h -- host side of K9: rsb200_scale_job -> ScaleJobDev (no CUDA here, so the CPU
// replay in tests/emu builds its jobs with exactly the code the library uses).
// Reference: RawImageDataU16::scaleValues / scaleValues_SSE2 / scaleValues_plain
// (common/RawImageDataU16.cpp:185-202, :204-292, :343-371): the float arithmetic below is
// the reference's, expression by expression.
#pragma once

#include "../../include/rawspeed_b200.h"
#include "scale_core.h"

namespace rsb200 {

// returns nullptr, or why the descriptor is refused; *mode: 0 SSE2 loop, 1 plain loop
inline const char* scale_build_job(const rsb200_scale_job& j, uint32_t quad_begin, ScaleJobDev* d,
                                   int* mode) {
  if ((j.offset & 15) || (j.pitch & 15) || !j.pitch)
    return "offset / pitch must be multiples of 16";
  if (!j.width || !j.height || j.width > 65535 || j.height > 65535 || j.cpp < 1 || j.cpp > 4)
    return "bad image dimensions";
  if ((uint64_t)j.width * j.cpp * 2 > j.pitch)
    return "pitch smaller than a row";
  if (!j.crop_w || !j.crop_h || (uint64_t)j.crop_x + j.crop_w > j.width ||
      (uint64_t)j.crop_y + j.crop_h > j.height)
    return "crop outside the image";
  if (j.path > RSB200_SCALE_PLAIN)
    return "bad path";
  const int depth_values = j.white_point - j.black_separate[0];
  if (depth_values == 0)
    return "white point equals the black level";
  const float app_scale = 65535.0F / (float)depth_values;
  const bool sse2 = j.path == RSB200_SCALE_AUTO ? app_scale < 63 : j.path == RSB200_SCALE_SSE2;
  for (int i = 0; i < 4; ++i)
    if (j.white_point == j.black_separate[i])
      return "white point equals a black level";
  ScaleJobDev o{};
  o.offset = j.offset;
  o.pitch = j.pitch;
  o.off_y = j.crop_y;
  o.crop_w = j.crop_w;
  o.crop_h = j.crop_h;
  o.full_fp = (int)(app_scale * 4.0F);
  o.half_fp = (int)(app_scale * 4095.0F);
  o.dither = j.dither ? 1u : 0u;
  o.quad_begin = quad_begin;
  if (sse2) {
    // whole uncropped rows, groups of 8 columns from column 0, x < roundDown(width, 8)
    o.group0 = 0;
    o.ngroups = j.width / 8;
    o.skip = 0;
    o.ncols = 8 * o.ngroups;
    for (uint32_t rp = 0; rp < 2; ++rp) {
      // sub_mul[] (:224-291): the pair for column parities 0 / 1 shares one 32-bit lane
      const int b0 = j.black_separate[2 * rp + (j.crop_x & 1)];
      const int b1 = j.black_separate[2 * rp + ((j.crop_x + 1) & 1)];
      const uint32_t m0 = (uint32_t)(int)(1024.0F * 65535.0F / (float)(j.white_point - b0));
      const uint32_t m1 = (uint32_t)(int)(1024.0F * 65535.0F / (float)(j.white_point - b1));
      const uint32_t mulv = m0 | (m1 << 16);
      const uint32_t subv = (uint32_t)b0 | ((uint32_t)b1 << 16);
      o.mul[2 * rp] = (int32_t)(mulv & 0xFFFFu);
      o.mul[2 * rp + 1] = (int32_t)(mulv >> 16);
      o.sub[2 * rp] = (int32_t)(subv & 0xFFFFu);
      o.sub[2 * rp + 1] = (int32_t)(subv >> 16);
    }
  } else {
    // the crop's samples only: [crop_x * cpp, (crop_x + crop_w) * cpp)
    const uint32_t col0 = j.crop_x * j.cpp;
    o.ncols = j.crop_w * j.cpp;
    o.group0 = col0 / 8;
    o.skip = col0 % 8;
    o.ngroups = (o.skip + o.ncols + 7) / 8;
    for (int i = 0; i < 4; ++i) {
      int v = i;
      if (j.crop_x & 1)
        v ^= 1;
      if (j.crop_y & 1)
        v ^= 2;
      o.mul[i] = (int)(16384.0F * 65535.0F / (float)(j.white_point - j.black_separate[v]));
      o.sub[i] = j.black_separate[v];
    }
  }
  *d = o;
  *mode = sse2 ? 0 : 1;
  return nullptr;
}

inline uint32_t scale_job_quads(const rsb200_scale_job& j) {
  return (j.crop_h + SCALE_ROWS - 1) / SCALE_ROWS;
}

} // namespace rsb200
