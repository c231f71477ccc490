// badpix_host.h -- host side of K11: bad-pixel positions -> bitmap (RawImageData::
// transferBadPixelsToMap, common/RawImage.cpp:201-229) + the de-duplicated list of pixels
// fixBadPixelsThread visits (:297-323).  No CUDA here: shared with the CPU replay in tests/emu.
#pragma once

#include "../../include/rawspeed_b200.h"
#include "badpix_core.h"

#include <vector>

namespace rsb200 {

// Appends image `j`'s bitmap to `maps` and its pixels to `list` (row-major, each once);
// `prior_map` (may be null): an existing mBadPixelMap of the image to OR into.
inline const char* badpix_build(const rsb200_badpix_job& j, const uint32_t* positions,
                                uint32_t npositions, const uint8_t* prior_map, BadPixJobDev* d,
                                std::vector<uint8_t>* maps, std::vector<uint32_t>* list) {
  if ((j.offset & 1) || (j.pitch & 1) || !j.width || !j.height || j.width > 65535 ||
      j.height > 65535 || (uint64_t)j.width * 2 > j.pitch)
    return "malformed image descriptor";
  if ((uint64_t)j.first_position + j.num_positions > npositions)
    return "positions outside the array";
  BadPixJobDev o{};
  o.offset = j.offset;
  o.pitch = j.pitch;
  o.width = j.width;
  o.height = j.height;
  o.step = j.is_cfa ? 2u : 1u;
  o.map_pitch = ((j.width + 7) / 8 + 15) / 16 * 16;
  while (maps->size() % 16)
    maps->push_back(0);
  o.map_offset = maps->size();
  const size_t map_bytes = (size_t)o.map_pitch * j.height;
  maps->resize(maps->size() + map_bytes, 0);
  uint8_t* map = maps->data() + o.map_offset;
  if (prior_map)
    for (size_t k = 0; k < map_bytes; ++k)
      map[k] = prior_map[k];
  for (uint32_t k = 0; k < j.num_positions; ++k) {
    const uint32_t pos = positions[j.first_position + k];
    const uint32_t px = pos & 0xFFFFu, py = pos >> 16;
    if (px >= j.width || py >= j.height)
      return "bad pixel position outside the image";
    map[(size_t)o.map_pitch * py + (px >> 3)] |= (uint8_t)(1u << (px & 7u));
  }
  // fixBadPixelsThread: blocks of 32 pixels, (w + 15) / 32 of them per row
  o.first = (uint32_t)list->size();
  const uint32_t gw = (j.width + 15) / 32;
  for (uint32_t y = 0; y < j.height; ++y)
    for (uint32_t x = 0; x < gw; ++x) {
      const uint8_t* block = map + (size_t)o.map_pitch * y + (size_t)x * 4;
      if (!(block[0] | block[1] | block[2] | block[3]))
        continue;
      for (uint32_t bi = 0; bi < 4; ++bi)
        for (uint32_t bj = 0; bj < 8; ++bj)
          if (((block[bi] >> bj) & 1) && x * 32 + bi * 8 + bj < j.width)
            // (a bit of a caller-supplied prior_map behind the last column is not a pixel: the
            //  kernel stores to the column unconditionally)
            list->push_back((y << 16) | (x * 32 + bi * 8 + bj));
    }
  o.count = (uint32_t)list->size() - o.first;
  *d = o;
  return nullptr;
}

} // namespace rsb200
