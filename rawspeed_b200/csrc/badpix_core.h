// badpix_core.h -- one bad pixel of K11 (bad-pixel interpolation), written so that the same
// source compiles as device code (badpix.cuh) and as plain C++ (tests/emu/badpix_emu.cpp).
//
// Reference: RawImageDataU16::fixBadPixel (common/RawImageDataU16.cpp:399-485) for cpp == 1:
// the nearest GOOD pixel (per the bad-pixel bitmap) to the left / right / above / below at step
// 2 (CFA) or 1, weighted by the opposite distance in 1/256ths, >> (7 + axes present).  Only
// good pixels are read and only bad pixels are written, so all bad pixels are independent.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define RSB_HD __host__ __device__ __forceinline__
#else
#define RSB_HD inline
#endif

namespace rsb200 {

struct BadPixJobDev {
  uint64_t offset;     // byte offset of row 0 of the uncropped image
  uint32_t pitch;      // bytes between rows
  uint32_t width, height;
  uint32_t step;       // 2 for CFA images, 1 otherwise
  uint64_t map_offset; // byte offset of this image's bitmap in the plan's bitmap buffer
  uint32_t map_pitch;  // roundUp(ceil(width / 8), 16) (RawImageData::createBadPixelMap)
  uint32_t first;      // this image's entries in the bad-pixel list: [first, first + count)
  uint32_t count;
  uint32_t pad;
};

RSB_HD bool badpix_is_bad(const uint8_t* map, uint32_t map_pitch, int y, int x) {
  return ((map[(size_t)map_pitch * (size_t)y + ((uint32_t)x >> 3)] >> (x & 7)) & 1) != 0;
}

// the interpolated value of bad pixel (x, y)
RSB_HD uint32_t badpix_value(const uint8_t* img, const BadPixJobDev& jb, const uint8_t* map, int x,
                             int y) {
  const int step = (int)jb.step, w = (int)jb.width, h = (int)jb.height;
  const uint8_t* const base = img + jb.offset;
  int values[4] = {-1, -1, -1, -1}, dist[4] = {0, 0, 0, 0};
  for (int xf = x - step; xf >= 0; xf -= step)
    if (!badpix_is_bad(map, jb.map_pitch, y, xf)) {
      values[0] = reinterpret_cast<const uint16_t*>(base + (size_t)y * jb.pitch)[xf];
      dist[0] = x - xf;
      break;
    }
  for (int xf = x + step; xf < w; xf += step)
    if (!badpix_is_bad(map, jb.map_pitch, y, xf)) {
      values[1] = reinterpret_cast<const uint16_t*>(base + (size_t)y * jb.pitch)[xf];
      dist[1] = xf - x;
      break;
    }
  for (int yf = y - step; yf >= 0; yf -= step)
    if (!badpix_is_bad(map, jb.map_pitch, yf, x)) {
      values[2] = reinterpret_cast<const uint16_t*>(base + (size_t)yf * jb.pitch)[x];
      dist[2] = y - yf;
      break;
    }
  for (int yf = y + step; yf < h; yf += step)
    if (!badpix_is_bad(map, jb.map_pitch, yf, x)) {
      values[3] = reinterpret_cast<const uint16_t*>(base + (size_t)yf * jb.pitch)[x];
      dist[3] = yf - y;
      break;
    }
  int weight[4] = {0, 0, 0, 0};
  int total_shifts = 7;
  if (const int tx = dist[0] + dist[1]; tx) {
    weight[0] = dist[0] ? (tx - dist[0]) * 256 / tx : 0;
    weight[1] = 256 - weight[0];
    total_shifts++;
  }
  if (const int ty = dist[2] + dist[3]; ty) {
    weight[2] = dist[2] ? (ty - dist[2]) * 256 / ty : 0;
    weight[3] = 256 - weight[2];
    total_shifts++;
  }
  int total = 0;
  for (int i = 0; i < 4; ++i)
    if (values[i] >= 0)
      total += values[i] * weight[i];
  total >>= total_shifts;
  return (uint32_t)(total < 0 ? 0 : (total > 65535 ? 65535 : total));
}

} // namespace rsb200
