// phaseone.cuh -- K8: Phase One IIQ row codec, sm_100a.
//
// Replaces the body of PhaseOneDecompressor::decompressStrip
// (decompressors/PhaseOneDecompressor.cpp:85-135) and the OpenMP loop over strips
// (:137-150).  One strip = one image row = one MSB32 bit stream (32-bit
// little-endian chunks consumed MSB first, BitStreamMSB32.h): every 8 pixels two
// code lengths are read (up to 5 zero bits, then one bit, into
// {8,7,6,9,11,10,5,12,14,13}; at column 0 any 1 bit in the prefix is an error), then
// each pixel is either a raw 16-bit value (length 14; always for the last width % 8
// pixels) or a difference to its same-parity predecessor.  Rows are independent,
// pixels of a row are not: one thread per row, 64-bit cache in registers,
// fill(32) before every pixel exactly like the reference (the over-read rule of
// BitStreamer.h:100-131 -- zero padding up to 8 bytes past the end, error beyond --
// is part of the result).
#pragma once

#include "common.cuh"

namespace rsb200 {

constexpr int P1_NT = 64;

struct P1StripDev {
  uint64_t in_offset;
  uint32_t in_size;
  uint32_t row;
  uint32_t job;
  uint32_t pad;
};
struct P1JobDev {
  uint64_t out_offset;
  uint32_t out_pitch;
  uint32_t width;
};

// 4 bytes at byte position pos of a strip, little endian, zero padded past its end
__device__ __forceinline__ uint32_t p1_chunk(const uint8_t* __restrict__ base, uint32_t size,
                                             uint32_t pos) {
  if (pos + 4u <= size) {
    const uint8_t* p = base + pos;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(p - mis);
    const uint32_t a = __ldg(pw), b = mis ? __ldg(pw + 1) : 0u;
    return __funnelshift_r(a, b, 8u * mis);
  }
  uint32_t v = 0;
  for (uint32_t k = 0; k < 4; ++k)
    if (pos + k < size)
      v |= (uint32_t)__ldg(base + pos + k) << (8 * k);
  return v;
}

__global__ void __launch_bounds__(P1_NT)
    p1_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
              const P1StripDev* __restrict__ strips, uint32_t nstrips,
              const P1JobDev* __restrict__ jobs, uint32_t* __restrict__ bad_jobs) {
  const uint32_t s = blockIdx.x * P1_NT + threadIdx.x;
  if (s >= nstrips)
    return;
  const P1StripDev st = strips[s];
  const P1JobDev jb = jobs[st.job];
  const uint8_t* base = in + st.in_offset;
  const uint32_t size = st.in_size, w = jb.width;
  uint32_t* o32 = reinterpret_cast<uint32_t*>(out + jb.out_offset + (uint64_t)st.row * jb.out_pitch);
  if (size < 4u) { // BitStreamer ctor: "Bit stream size is smaller than MaxProcessBytes"
    atomicOr(bad_jobs + st.job, 1u);
    return;
  }
  uint32_t hi = 0, lo = 0, pos = 0;
  int nbits = 0;
  int32_t pred0 = 0, pred1 = 0;
  uint32_t len0 = 0, len1 = 0, even = 0;
  bool bad = false;
  auto getbits = [&](uint32_t n) { // 1 <= n <= 16, n <= nbits
    const uint32_t v = hi >> (32u - n);
    hi = __funnelshift_l(lo, hi, n);
    lo <<= n;
    nbits -= (int)n;
    return v;
  };
  const uint32_t lim = w & ~7u;
  for (uint32_t col = 0; col < w && !bad; ++col) {
    // pump.fill(32)
    if (nbits < 32) {
      if (pos > size + 8u) { // "Buffer overflow read in BitStreamer"
        bad = true;
        break;
      }
      const uint32_t ch = p1_chunk(base, size, pos);
      hi |= __funnelshift_rc(ch, 0u, (uint32_t)nbits);
      lo = __funnelshift_lc(0u, ch, 32u - (uint32_t)nbits);
      nbits += 32;
      pos += 4;
    }
    if (col >= lim) {
      len0 = len1 = 14;
    } else if ((col & 7u) == 0u) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        uint32_t j = 0;
        for (; j < 5; ++j) {
          if (getbits(1) != 0u) {
            if (col == 0)
              bad = true; // "Can not initialize lengths. Data is corrupt."
            break;
          }
        }
        if (j > 0) {
          // length[2 * (j - 1) + bit] of {8, 7, 6, 9, 11, 10, 5, 12, 14, 13}, packed in nibbles
          const uint32_t idx = 2u * (j - 1u) + getbits(1);
          const uint32_t l = (uint32_t)((0xDEC5AB9678ull >> (4u * idx)) & 15u);
          if (t == 0)
            len0 = l;
          else
            len1 = l;
        }
      }
      if (bad)
        break;
    }
    const uint32_t i = (col & 1u) ? len1 : len0;
    int32_t& pred = (col & 1u) ? pred1 : pred0;
    if (i == 14u)
      pred = (int32_t)getbits(16);
    else
      pred += (int32_t)getbits(i) + 1 - (1 << (i - 1u));
    if (col & 1u)
      o32[col >> 1] = even | ((uint32_t)pred << 16);
    else
      even = (uint32_t)pred & 0xFFFFu;
  }
  if (bad)
    atomicOr(bad_jobs + st.job, 1u);
}

// ---- second version: no loads inside a group of 8 pixels ----
// All rows of a warp are at the same column at the same time; only their bit positions
// differ.  A per-lane "refill when my cache runs low" therefore issues a load at almost
// every pixel for SOME lane, and the warp-wide scoreboard of the destination register makes
// every lane wait for it (the lesson of K2T).  A group of 8 pixels consumes at most 14 + 8*16
// = 142 bits, so the 7 chunks from the one that holds the group's first bit are fetched at
// the group boundary -- the same instructions for all lanes -- and the group is decoded from
// registers: window (cur, nxt) + a 5-deep register queue shifted at chunk crossings.
// The reference's over-read rule (BitStreamer.h:100-131) in closed form: refill k reads at
// byte 4k and fails when 4k > size + 8; the pixel that starts at bit T needs refills
// 0 .. T/32 + (T % 32 ? 1 : 0), so the row fails iff some pixel starts at a bit
// T > 32 * (K - 1) with K = (size + 8) / 4 + 1.
__global__ void __launch_bounds__(P1_NT)
    p1_kernel_v2(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                 const P1StripDev* __restrict__ strips, uint32_t nstrips,
                 const P1JobDev* __restrict__ jobs, uint32_t* __restrict__ bad_jobs) {
  const uint32_t s_raw = blockIdx.x * P1_NT + threadIdx.x;
  const bool live = s_raw < nstrips;
  const P1StripDev st = strips[live ? s_raw : nstrips - 1u];
  const P1JobDev jb = jobs[st.job];
  const uint8_t* base = in + st.in_offset;
  const uint32_t size = st.in_size, w = jb.width;
  uint32_t* o32 = reinterpret_cast<uint32_t*>(out + jb.out_offset + (uint64_t)st.row * jb.out_pitch);
  bool bad = size < 4u; // BitStreamer ctor: "Bit stream size is smaller than MaxProcessBytes"
  const uint32_t tmax = 32u * ((size + 8u) / 4u); // a pixel may start at bit T <= tmax
  uint32_t p = 0; // bit position in the row's stream
  uint32_t cur = 0, nxt = 0, q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4 = 0;
  int32_t pred0 = 0, pred1 = 0;
  uint32_t len0 = 0, len1 = 0, even = 0;
  auto getbits = [&](uint32_t n) { // 1 <= n <= 16
    const uint32_t v = __funnelshift_l(nxt, cur, p) >> (32u - n);
    const uint32_t pn = p + n;
    if ((pn ^ p) & 32u) { // into the next chunk
      cur = nxt;
      nxt = q0;
      q0 = q1;
      q1 = q2;
      q2 = q3;
      q3 = q4;
    }
    p = pn;
    return v;
  };
  const uint32_t lim = w & ~7u;
  for (uint32_t col = 0; col < w; ++col) {
    if ((col & 7u) == 0u) { // group boundary: chunks p/32 .. p/32 + 6, for every lane
      const uint32_t c0 = p >> 5;
      cur = p1_chunk(base, size, 4u * c0);
      nxt = p1_chunk(base, size, 4u * c0 + 4u);
      q0 = p1_chunk(base, size, 4u * c0 + 8u);
      q1 = p1_chunk(base, size, 4u * c0 + 12u);
      q2 = p1_chunk(base, size, 4u * c0 + 16u);
      q3 = p1_chunk(base, size, 4u * c0 + 20u);
      q4 = p1_chunk(base, size, 4u * c0 + 24u);
    }
    if (p > tmax) // pump.fill(32) of this pixel would read past size + 8
      bad = true;
    if (col >= lim) {
      len0 = len1 = 14;
    } else if ((col & 7u) == 0u) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        uint32_t j = 0;
        for (; j < 5; ++j) {
          if (getbits(1) != 0u) {
            if (col == 0)
              bad = true; // "Can not initialize lengths. Data is corrupt."
            break;
          }
        }
        if (j > 0) {
          const uint32_t idx = 2u * (j - 1u) + getbits(1);
          const uint32_t l = (uint32_t)((0xDEC5AB9678ull >> (4u * idx)) & 15u);
          if (t == 0)
            len0 = l;
          else
            len1 = l;
        }
      }
    }
    uint32_t i = (col & 1u) ? len1 : len0;
    if (i == 0u) // (only after an error at column 0: keep the arithmetic defined)
      i = 14u;
    int32_t& pred = (col & 1u) ? pred1 : pred0;
    if (i == 14u)
      pred = (int32_t)getbits(16);
    else
      pred += (int32_t)getbits(i) + 1 - (1 << (i - 1u));
    if (col & 1u) {
      if (live && !bad)
        o32[col >> 1] = even | ((uint32_t)pred << 16);
    } else {
      even = (uint32_t)pred & 0xFFFFu;
    }
  }
  if (bad && live)
    atomicOr(bad_jobs + st.job, 1u);
}

} // namespace rsb200
