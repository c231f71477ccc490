// arw2.cuh -- K6: Sony ARW2 block codec, sm_100a.
//
// Replaces the body of SonyArw2Decompressor::decompressRow
// (decompressors/SonyArw2Decompressor.cpp:58-112) and the per-row OpenMP loop
// (:114-133): one byte per pixel; a row is an LSB-first bit stream of 128-bit
// blocks, each block = max(11) min(11) imax(4) imin(4) + 14 x 7-bit deltas for 16
// pixels of the same column parity (the blocks of the even and of the odd pixels
// of 32 columns follow each other).  Every value goes through
// RawImageDataU16::setWithLookUp (common/RawImage.h:335-353).
//
// One thread = 32 columns = two blocks = 32 input bytes -> 64 output bytes (four
// 128-bit stores).  The only serial dependency of the reference, the dither
// state of setWithLookUp (r' = 15700*(r & 65535) + (r >> 16), seeded per row with
// the row's first 24 bits), is a multiply-with-carry generator: for r below
// m = 15700*2^16 - 1 it equals r' = 15700*r mod m, so the state at call n of the
// row is r0 * 15700^n mod m and a thread jumps to its 32 calls with one modular
// multiplication by a constant from a 300-entry table.
#pragma once

#include "common.cuh"

namespace rsb200 {

constexpr int ARW2_NT = 256;
constexpr uint32_t ARW2_M = 15700u * 65536u - 1u; // modulus of the dither generator
constexpr int ARW2_MAX_GROUPS = 300;              // 9600 / 32 columns

struct Arw2JobDev {
  uint64_t in_offset;
  uint64_t out_offset;
  uint32_t out_pitch;
  uint32_t width;
  uint32_t height;
  uint32_t groups_per_row; // width / 32
  uint32_t group_begin;    // first group of this job in the plan
  int32_t table;           // plan table index (-1: no table)
};

__constant__ uint32_t c_arw2_jump[ARW2_MAX_GROUPS]; // 15700^(32 g) mod ARW2_M

// MODE 0: no table, 1: plain table (4096 x u16), 2: dithered table (4096 x {base, delta})
template <int MODE>
__device__ __forceinline__ uint32_t arw2_lookup(const uint16_t* __restrict__ tab, uint32_t value,
                                                uint32_t& r) {
  if (MODE == 0)
    return value;
  if (MODE == 1)
    return __ldg(tab + value);
  const uint32_t e = __ldg(reinterpret_cast<const uint32_t*>(tab) + value);
  const uint32_t base = e & 0xFFFFu, delta = e >> 16;
  const uint32_t pix = base + ((delta * (r & 2047u) + 1024u) >> 12);
  r = 15700u * (r & 65535u) + (r >> 16);
  return pix & 0xFFFFu;
}

// 16 pixels of one block (w0..w3 = the block as a little-endian 128-bit number)
template <int MODE>
__device__ __forceinline__ bool arw2_block(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3,
                                           const uint16_t* __restrict__ tab, uint32_t& r,
                                           uint32_t (&px)[16]) {
  const int vmax = (int)(w0 & 0x7FFu), vmin = (int)((w0 >> 11) & 0x7FFu);
  const uint32_t imax = (w0 >> 22) & 15u, imin = (w0 >> 26) & 15u;
  int sh = 0;
  while (sh < 4 && (0x80 << sh) <= (vmax - vmin))
    ++sh;
  // the 98 delta bits start at bit 30
  w0 = __funnelshift_r(w0, w1, 30);
  w1 = __funnelshift_r(w1, w2, 30);
  w2 = __funnelshift_r(w2, w3, 30);
  w3 >>= 30;
#pragma unroll
  for (uint32_t i = 0; i < 16; ++i) {
    int p;
    if (i == imax)
      p = vmax;
    else if (i == imin)
      p = vmin;
    else {
      p = (int)((w0 & 127u) << sh) + vmin;
      p = min(p, 0x7ff);
      w0 = __funnelshift_r(w0, w1, 7);
      w1 = __funnelshift_r(w1, w2, 7);
      w2 = __funnelshift_r(w2, w3, 7);
      w3 >>= 7;
    }
    px[i] = arw2_lookup<MODE>(tab, (uint32_t)p << 1, r);
  }
  return imax != imin;
}

template <int MODE>
__global__ void __launch_bounds__(ARW2_NT)
    arw2_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                const Arw2JobDev* __restrict__ jobs, int njobs, uint32_t total_groups,
                const uint16_t* __restrict__ tables, uint32_t* __restrict__ bad_jobs) {
  const uint32_t g = blockIdx.x * ARW2_NT + threadIdx.x;
  if (g >= total_groups)
    return;
  int lo = 0, hi = njobs - 1; // job of this group
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].group_begin <= g)
      lo = mid;
    else
      hi = mid - 1;
  }
  const Arw2JobDev jb = jobs[lo];
  const uint32_t gl = g - jb.group_begin;
  const uint32_t row = gl / jb.groups_per_row, gx = gl - row * jb.groups_per_row;
  const uint8_t* rowp = in + jb.in_offset + (uint64_t)row * jb.width;
  // 32 input bytes at any alignment: nine aligned words, funnel-shifted
  uint32_t w[8];
  {
    const uint8_t* p = rowp + gx * 32u;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(p - mis);
    uint32_t a[9];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      a[k] = __ldg(pw + k);
    a[8] = mis ? __ldg(pw + 8) : 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      w[k] = __funnelshift_r(a[k], a[k + 1], 8u * mis);
  }
  const uint16_t* tab = nullptr;
  uint32_t r = 0;
  if (MODE != 0)
    tab = tables + (size_t)jb.table * (MODE == 2 ? 8192u : 4096u);
  if (MODE == 2) {
    // dither state of the row before call 32*gx
    const uint32_t r0 = (uint32_t)rowp[0] | ((uint32_t)rowp[1] << 8) | ((uint32_t)rowp[2] << 16);
    r = (uint32_t)(((uint64_t)r0 * c_arw2_jump[gx]) % ARW2_M);
  }
  uint32_t ev[16], od[16];
  bool ok = arw2_block<MODE>(w[0], w[1], w[2], w[3], tab, r, ev);
  ok = arw2_block<MODE>(w[4], w[5], w[6], w[7], tab, r, od) && ok;
  if (!ok) // "ARW2 invariant failed, same pixel is both min and max"
    atomicOr(bad_jobs + lo, 1u);
  uint8_t* op = out + jb.out_offset + (uint64_t)row * jb.out_pitch + gx * 64u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    o.x = ev[4 * q + 0] | (od[4 * q + 0] << 16);
    o.y = ev[4 * q + 1] | (od[4 * q + 1] << 16);
    o.z = ev[4 * q + 2] | (od[4 * q + 2] << 16);
    o.w = ev[4 * q + 3] | (od[4 * q + 3] << 16);
    stg_cs_v4(op + 16 * q, o);
  }
}

} // namespace rsb200
