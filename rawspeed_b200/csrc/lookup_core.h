// lookup_core.h -- the per-lane program of K12 (16-bit table lookup of a whole image), written
// so that the same source compiles as device code (lookup.cuh) and as plain C++ (the CPU
// replay in tests/emu/lookup_emu.cpp).
//
// Reference: RawImageDataU16::doLookup (common/RawImageDataU16.cpp:487-520), the APPLY_LOOKUP
// worker of RawImageData::sixteenBitLookup (common/RawImage.cpp:373-378): every sample of
// every row of the uncropped buffer goes through TableLookUp table 0; with dither the table
// holds {base, delta} pairs and a per-row multiply-with-carry generator
//   v' = 15700 * (v & 65535) + (v >> 16),  seed (width + 13 * y) ^ 0x45694584,
// advanced once per sample, picks the step inside the delta.  v' = 15700 v mod (15700 * 2^16
// - 1) for every state below that modulus -- the seeds lie ABOVE it (0x4569.... > 0x3D53....),
// so the first step (rarely two) is taken explicitly and the rest is jumped.
#pragma once

#include "scale_core.h" // ScaleVec, RS_HD

namespace rsb200 {

struct LookupJobDev {
  uint64_t offset;    // byte offset of row 0 of the uncropped image (multiple of 16)
  uint32_t pitch;     // bytes between rows (multiple of 16)
  uint32_t width;     // uncropped_dim.x (the seed uses the PIXEL width)
  uint32_t height;
  uint32_t ncols;     // width * cpp samples per row
  uint32_t ngroups;   // ceil(ncols / 8)
  uint32_t table;     // index of the image's table
  uint32_t quad_begin; // first global row quad of this job
  uint32_t pad;
};

constexpr uint32_t LUT_MWC_A = 15700u;
constexpr uint32_t LUT_MWC_M = 15700u * 65536u - 1u; // 1028915199

RS_HD uint32_t lut_mwc_step(uint32_t v) { return LUT_MWC_A * (v & 65535u) + (v >> 16); }
RS_HD uint32_t lut_mulmod(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % LUT_MWC_M); }
RS_HD uint32_t lut_powmod(uint32_t e) { // 15700^e mod M
  uint32_t r = 1, b = LUT_MWC_A;
  while (e) {
    if (e & 1)
      r = lut_mulmod(r, b);
    b = lut_mulmod(b, b);
    e >>= 1;
  }
  return r;
}
// state BEFORE sample x of row y (the seed after x steps)
RS_HD uint32_t lut_mwc_state(uint32_t width, uint32_t y, uint32_t x) {
  uint32_t v = (width + y * 13u) ^ 0x45694584u;
  while (x && v >= LUT_MWC_M) {
    v = lut_mwc_step(v);
    --x;
  }
  return x ? lut_mulmod(v, lut_powmod(x)) : v;
}
RS_HD uint32_t lut_mwc_jump(uint32_t v, uint32_t n, uint32_t an) {
  if (v < LUT_MWC_M)
    return lut_mulmod(v, an);
  while (n--)
    v = lut_mwc_step(v);
  return v;
}

// eight samples starting at sample x0 of a row; samples at or beyond ncols are returned
// unchanged and do not advance the generator.  DITHER: table = {base, delta} pairs, read as one
// 32-bit word per value; v = generator state before sample x0, updated.
template <bool DITHER>
RS_HD ScaleVec lut_group(const ScaleVec& in, const uint16_t* table, uint32_t ncols, uint32_t x0,
                         uint32_t& v) {
  ScaleVec o;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int q = 0; q < 4; ++q) {
    uint32_t out2[2];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int h = 0; h < 2; ++h) {
      const uint32_t x = x0 + 2u * (uint32_t)q + (uint32_t)h;
      const uint32_t p = h ? in.w[q] >> 16 : in.w[q] & 0xFFFFu;
      uint32_t res = p;
      if (x < ncols) {
        if (DITHER) {
          const uint32_t bd = reinterpret_cast<const uint32_t*>(table)[p]; // base | delta << 16
          v = lut_mwc_step(v);
          const uint32_t pix = (bd & 0xFFFFu) + (((bd >> 16) * (v & 2047u) + 1024u) >> 12);
          res = pix > 65535u ? 65535u : pix; // clampBits(pix, 16)
        } else {
          res = table[p];
        }
      }
      out2[h] = res;
    }
    o.w[q] = out2[0] | (out2[1] << 16);
  }
  return o;
}

} // namespace rsb200
