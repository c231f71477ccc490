// scale_core.h -- the per-lane program of K9 (black/white scaling), written so that the
// SAME source compiles as device code (scale.cuh) and as plain C++ (tests/emu/scale_emu.cpp
// replays the kernel's warp loop on the CPU against the oracle).
//
// Reference (paths relative to /root/reference/src/librawspeed):
//   RawImageDataU16::scaleValues_SSE2   common/RawImageDataU16.cpp:204-341
//   RawImageDataU16::scaleValues_plain  common/RawImageDataU16.cpp:343-399
//
// Work decomposition: one warp = four consecutive crop rows ("row quad"); per iteration a
// lane owns one 16-byte-aligned group of eight uint16 samples of each of the four rows.
//
//  SSE2 semantics: the dither source is eight 16-bit multiplicative generators per row
//  (state' = hi16(state * m) ^ lo16(state * m), signed product) advanced once per group of
//  eight columns -- sequential along the row, so it cannot be jumped.  Phase A: lane
//  (r = lane / 8, k = lane % 8) advances generator k of row r by 32 steps and leaves the low
//  byte of every step in shared memory; phase B: each lane scales its eight samples of each
//  row with the bytes of its own group.
//
//  Plain semantics: one multiply-with-carry generator per row, v' = 18000 * (v & 65535) +
//  (v >> 16) = 18000 * v mod (18000 * 2^16 - 1) once v is below the modulus, advanced once
//  per SAMPLE: a lane jumps to its group with one modular multiplication by a power of
//  18000 and steps through its eight samples.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define RS_HD __host__ __device__ __forceinline__
#else
#define RS_HD inline
#endif

namespace rsb200 {

struct ScaleJobDev {
  uint64_t offset;   // byte offset of row 0 of the uncropped image (multiple of 16)
  uint32_t pitch;    // bytes between rows (multiple of 16)
  uint32_t off_y;    // mOffset.y
  uint32_t crop_w;   // dim.x (pixels): seeds of the dither generators
  uint32_t crop_h;   // dim.y
  uint32_t group0;   // first 8-sample group of a row that holds work
  uint32_t ngroups;  // groups per row that hold work
  uint32_t skip;     // plain: samples of group0 left of the crop (col0 % 8); SSE2: 0
  uint32_t ncols;    // plain: crop_w * cpp samples; SSE2: 8 * ngroups
  int32_t mul[4];    // SSE2: [2*(buffer row & 1) + (buffer column & 1)], 16-bit values
  int32_t sub[4];    // plain: [2*(crop row & 1) + (crop sample & 1)]
  int32_t full_fp, half_fp; // full_scale_fp, half_scale_fp
  uint32_t dither;   // mDitherScale
  uint32_t quad_begin; // first global row quad of this job
};

constexpr int SCALE_ROWS = 4;           // rows per warp
constexpr int SCALE_RND_STRIDE = 264;   // bytes of dither per row and iteration (256 + 8 pad:
                                        // the four rows land in different banks)
constexpr uint32_t SCALE_MWC_A = 18000u;
constexpr uint32_t SCALE_MWC_M = 18000u * 65536u - 1u; // 1179647999

struct ScaleVec { // eight uint16 samples
  uint32_t w[4];
};

// ---------------------------------------------------------------- SSE2 semantics
// seed of generator k of crop row y (_mm_set_epi32 at :300-304, 16-bit lanes)
RS_HD int32_t scale_sse2_seed(uint32_t crop_w, uint32_t y, int k) {
  const uint32_t q = (uint32_t)k >> 1;
  const uint32_t a = q == 0 ? 1234u : (q == 1 ? 4272u : (q == 2 ? 2342u : 1676u));
  const uint32_t b = q == 0 ? 23464u : (q == 1 ? 12123u : (q == 2 ? 34311u : 18000u));
  const uint32_t l = crop_w * a + y * b;
  return (int32_t)(int16_t)((k & 1) ? (l >> 16) : l);
}

// one step: sserandom = mulhi_epi16(r, m) ^ mullo_epi16(r, m); state kept sign-extended
RS_HD int32_t scale_sse2_step(int32_t s, int k) {
  const int32_t m = (k & 1) ? 0x4d9f : 0x1d32;
  const uint32_t prod = (uint32_t)(s * m);
  return (int32_t)(int16_t)((prod >> 16) ^ prod);
}

// phase A: 32 steps of generator (lane & 7) of row (lane >> 3); byte of step t -> rnd[r][t][k]
RS_HD void scale_sse2_advance(int32_t& state, int lane, uint8_t* rnd) {
  const int k = lane & 7;
  uint8_t* dst = rnd + (lane >> 3) * SCALE_RND_STRIDE + k;
#if defined(__CUDACC__)
#pragma unroll 8
#endif
  for (int t = 0; t < 32; ++t) {
    state = scale_sse2_step(state, k);
    dst[t * 8] = (uint8_t)state;
  }
}

// one sample: subs_epu16, 16x16 -> 32 multiply, + round + dither, >> 10, saturate to 16 bit
RS_HD uint32_t scale_sse2_sample(uint32_t p, uint32_t sub16, uint32_t mul16, uint32_t rbyte,
                                 uint32_t full16, uint32_t kround) {
  const uint32_t pix = p > sub16 ? p - sub16 : 0u;
  const uint32_t r16 = (rbyte * full16) & 0xFFFFu; // mullo_epi16(rand & 0xff, full_scale_fp)
  int32_t v = (int32_t)(pix * mul16 + kround - r16); // epi32 adds wrap
  v >>= 10;                                          // srai
  return (uint32_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v)); // -32768, packs_epi32, ^0x8000
}

// eight samples of a row of buffer row parity rp; rb_lo / rb_hi = the eight dither bytes of
// this group.  (No dynamic indexing of the job: it lives in registers.)
RS_HD ScaleVec scale_sse2_group(const ScaleVec& in, const ScaleJobDev& j, uint32_t rp,
                                uint32_t rb_lo, uint32_t rb_hi) {
  const uint32_t sub0 = (uint32_t)(rp ? j.sub[2] : j.sub[0]) & 0xFFFFu;
  const uint32_t sub1 = (uint32_t)(rp ? j.sub[3] : j.sub[1]) & 0xFFFFu;
  const uint32_t mul0 = (uint32_t)(rp ? j.mul[2] : j.mul[0]) & 0xFFFFu;
  const uint32_t mul1 = (uint32_t)(rp ? j.mul[3] : j.mul[1]) & 0xFFFFu;
  const uint32_t full16 = (uint32_t)j.full_fp & 0xFFFFu;
  const uint32_t kround = 512u + (uint32_t)(j.half_fp >> 4);
  ScaleVec o;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int q = 0; q < 4; ++q) {
    const uint32_t rb = q < 2 ? rb_lo : rb_hi;
    const uint32_t b0 = (rb >> (16 * (q & 1))) & 0xFFu, b1 = (rb >> (16 * (q & 1) + 8)) & 0xFFu;
    const uint32_t lo = scale_sse2_sample(in.w[q] & 0xFFFFu, sub0, mul0, b0, full16, kround);
    const uint32_t hi = scale_sse2_sample(in.w[q] >> 16, sub1, mul1, b1, full16, kround);
    o.w[q] = lo | (hi << 16);
  }
  return o;
}

// ---------------------------------------------------------------- plain semantics
RS_HD uint32_t scale_mwc_step(uint32_t v) { // v = 18000 * (v & 65535) + (v >> 16)
  return SCALE_MWC_A * (v & 65535u) + (v >> 16);
}
RS_HD uint32_t scale_mulmod(uint32_t a, uint32_t b) {
  return (uint32_t)(((uint64_t)a * b) % SCALE_MWC_M);
}
RS_HD uint32_t scale_powmod(uint32_t e) { // 18000^e mod M
  uint32_t r = 1, b = SCALE_MWC_A;
  while (e) {
    if (e & 1)
      r = scale_mulmod(r, b);
    b = scale_mulmod(b, b);
    e >>= 1;
  }
  return r;
}
// state BEFORE crop sample x of crop row y (v after x steps from the seed)
RS_HD uint32_t scale_mwc_state(uint32_t crop_w, uint32_t y, uint32_t x) {
  uint32_t v = crop_w + y * 36969u;
  // v' = 18000 v mod M holds from the first state below M on; the seed (and, for one seed
  // in 2^32, a few of its successors) can lie above
  while (x && v >= SCALE_MWC_M) {
    v = scale_mwc_step(v);
    --x;
  }
  return x ? scale_mulmod(v, scale_powmod(x)) : v;
}

// n steps on from v; an = 18000^n mod M
RS_HD uint32_t scale_mwc_jump(uint32_t v, uint32_t n, uint32_t an) {
  if (v < SCALE_MWC_M)
    return scale_mulmod(v, an);
  while (n--)
    v = scale_mwc_step(v);
  return v;
}

RS_HD uint32_t scale_plain_sample(uint32_t p, int32_t sub, int32_t mul, int32_t rnd) {
  const int32_t v = (int32_t)((uint32_t)((int32_t)p - sub) * (uint32_t)mul + 8192u + (uint32_t)rnd) >> 14;
  return (uint32_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v)); // clampBits(.., 16)
}

// eight samples starting at crop sample x0 (may be negative / run past ncols: those samples
// are returned unchanged and do not advance the generator); v = state before sample
// max(x0, 0), updated to the state after the last sample processed.  x0 has the parity of
// j.skip (groups start at multiples of 8), so the table entry of in-group position k is
// [2*(y & 1) + ((skip ^ k) & 1)].
RS_HD ScaleVec scale_plain_group(const ScaleVec& in, const ScaleJobDev& j, uint32_t y, int32_t x0,
                                 uint32_t& v) {
  const bool yo = (y & 1u) != 0, so = (j.skip & 1u) != 0;
  const int32_t sub_a = yo ? j.sub[2] : j.sub[0], sub_b = yo ? j.sub[3] : j.sub[1];
  const int32_t mul_a = yo ? j.mul[2] : j.mul[0], mul_b = yo ? j.mul[3] : j.mul[1];
  const int32_t sub_k0 = so ? sub_b : sub_a, sub_k1 = so ? sub_a : sub_b; // k even / k odd
  const int32_t mul_k0 = so ? mul_b : mul_a, mul_k1 = so ? mul_a : mul_b;
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
      const int32_t x = x0 + 2 * q + h;
      const uint32_t p = h ? in.w[q] >> 16 : in.w[q] & 0xFFFFu;
      uint32_t res = p;
      if (x >= 0 && (uint32_t)x < j.ncols) {
        int32_t rnd = 0;
        if (j.dither) {
          v = scale_mwc_step(v);
          rnd = j.half_fp - j.full_fp * (int32_t)(v & 2047u);
        }
        res = scale_plain_sample(p, h ? sub_k1 : sub_k0, h ? mul_k1 : mul_k0, rnd);
      }
      out2[h] = res;
    }
    o.w[q] = out2[0] | (out2[1] << 16);
  }
  return o;
}

} // namespace rsb200
