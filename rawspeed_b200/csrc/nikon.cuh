// nikon.cuh -- K3N: reconstruction step of NikonDecompressor::decompress
// (decompressors/NikonDecompressor.cpp:513-560) for streams without a split.
// The entropy decode is the shared multi-CTA path (ljpeg_ranges.cuh) with the plain
// MSB bit source and the nikon_tree table; it leaves the differences, in stream
// order, in the linear int16 scratch buffer.  Here
//
//   value(r, c) = pUp[r & 1][c] + sum of d(r', c) over rows r' <= r of r's parity   (c = 0, 1)
//   value(r, k) = value(r, k-2) + d(r, k)                                           (k >= 2)
//
// in plain int arithmetic, then clampBits(value, 15) and
// RawImageDataU16::setWithLookUp (common/RawImage.h:335-353) with the curve as a
// dithered table.  The dither state is seeded ONCE with the first 24 bits of the
// stream and stepped once per pixel in raster order, r' = 15700*(r & 65535) + (r >> 16);
// as in arw2.cuh this equals r' = 15700*r mod m (m = 15700*2^16 - 1), so pixel n sees
// r0 * 15700^n mod m: a warp computes 15700^(row*width) by square-and-multiply, a lane
// adds its column offset with one more modular multiplication.
//
//   k3n_column_kernel  one warp per (segment, row parity, column 0/1): int32 scan
//   k3n_row_kernel     one warp per row: per-parity int32 prefix sums, clamp, curve
#pragma once

#include "arw2.cuh"
#include "ljpeg.cuh"

namespace rsb200 {

__device__ __forceinline__ uint32_t n_mulmod(uint32_t a, uint32_t b) {
  return (uint32_t)(((uint64_t)a * b) % ARW2_M);
}
// 15700^e mod m
__device__ __forceinline__ uint32_t n_powmod(uint32_t e) {
  uint32_t base = 15700u, acc = 1u;
  while (e) {
    if (e & 1u)
      acc = n_mulmod(acc, base);
    base = n_mulmod(base, base);
    e >>= 1;
  }
  return acc;
}

__global__ void k3n_column_kernel(const DevScan* __restrict__ scans,
                                  const uint32_t* __restrict__ scan_ids, int nscans,
                                  const uint16_t* __restrict__ diffs,
                                  uint16_t* __restrict__ colvals) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int si = warp >> 2;
  if (si >= nscans)
    return;
  const uint32_t scan = scan_ids ? scan_ids[si] : (uint32_t)si;
  const DevScan& sc = scans[scan];
  if (sc.kind != 3)
    return;
  const uint32_t q = (warp >> 1) & 1u, c = warp & 1u; // row parity, column
  const int16_t* d = reinterpret_cast<const int16_t*>(diffs + sc.diff_offset) + c;
  // two int32 per row in the column scratch (4 uint16 slots)
  int32_t* cv = reinterpret_cast<int32_t*>(colvals + sc.col_offset) + c;
  const uint32_t nj = (sc.rows > q) ? (sc.rows - q + 1) / 2 : 0;
  int run = (int)sc.init_pred[q * 2 + c]; // pUp[q][c]
  for (uint32_t j0 = 0; j0 < nj; j0 += 32) {
    const uint32_t j = j0 + lane, r = q + 2 * j;
    int v = (j < nj) ? (int)d[(uint64_t)r * sc.row_samples] : 0;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
      const int n = __shfl_up_sync(0xFFFFFFFFu, v, k);
      if (lane >= k)
        v += n;
    }
    v += run;
    if (j < nj)
      cv[(uint64_t)r * 2] = v;
    run = __shfl_sync(0xFFFFFFFFu, v, 31);
  }
}

__device__ __forceinline__ uint32_t n_store_value(int v, const uint32_t* __restrict__ lut,
                                                  uint32_t& r) {
  const uint32_t value = (uint32_t)min(max(v, 0), 32767); // clampBits(v, 15)
  if (!lut)
    return value;
  const uint32_t e = __ldg(lut + value);
  const uint32_t pix = (e & 0xFFFFu) + (((e >> 16) * (r & 2047u) + 1024u) >> 12);
  r = 15700u * (r & 65535u) + (r >> 16);
  return pix & 0xFFFFu;
}

__global__ void __launch_bounds__(K3_THREADS)
    k3n_row_kernel(const uint8_t* __restrict__ in, const DevScan* __restrict__ scans,
                   const K3RowRef* __restrict__ rows, uint32_t nrows,
                   const uint16_t* __restrict__ diffs, const uint16_t* __restrict__ colvals,
                   const uint16_t* __restrict__ luts, uint8_t* __restrict__ out) {
  const uint32_t wrow = (blockIdx.x * K3_THREADS + threadIdx.x) >> 5;
  if (wrow >= nrows)
    return;
  const K3RowRef ref = rows[wrow];
  const DevScan& sc = scans[ref.scan];
  if (sc.kind != 3)
    return;
  const int lane = threadIdx.x & 31;
  const uint32_t r = ref.row, npairs = sc.row_samples / 2;
  const uint32_t* d = reinterpret_cast<const uint32_t*>(diffs + sc.diff_offset +
                                                        (uint64_t)r * sc.row_samples);
  uint32_t* o = reinterpret_cast<uint32_t*>(out + sc.out_offset + (uint64_t)r * sc.out_pitch);
  const int32_t* cv = reinterpret_cast<const int32_t*>(colvals + sc.col_offset) + (uint64_t)r * 2;
  // pad0[0] = 1 + index of the dithered table of this image (0: no table)
  const uint32_t* lut = sc.pad0[0]
                            ? reinterpret_cast<const uint32_t*>(luts) + (size_t)(sc.pad0[0] - 1) * 65536u
                            : nullptr;
  constexpr uint32_t PER = 4; // pairs per lane and step
  uint32_t rnd = 0;
  if (lut) {
    const uint8_t* s = in + sc.in_offset; // random = bits.peekBits(24)
    const uint32_t r0 = ((uint32_t)s[0] << 16) | ((uint32_t)s[1] << 8) | (uint32_t)s[2];
    // state before pixel (r, 2*PER*lane)
    rnd = n_mulmod(n_mulmod(r0, n_powmod(r * sc.row_samples)), n_powmod(2u * PER * (uint32_t)lane));
  }
  const uint32_t jump = lut ? n_powmod(2u * PER * 31u) : 0u; // to my pixels of the next step
  int run0 = 0, run1 = 0;
  for (uint32_t p0 = 0; p0 < npairs; p0 += 32 * PER) {
    const uint32_t pb = p0 + lane * PER;
    int e[PER], f[PER];
    int s0 = 0, s1 = 0;
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) {
      const uint32_t p = pb + k;
      int a = 0, b = 0;
      if (p < npairs) {
        if (p == 0) { // the column kernel already holds the first two values of the row
          a = cv[0];
          b = cv[1];
        } else {
          const uint32_t w = __ldg(d + p);
          a = (int)(int16_t)(w & 0xFFFFu);
          b = (int)(int16_t)(w >> 16);
        }
      }
      s0 += a;
      s1 += b;
      e[k] = s0;
      f[k] = s1;
    }
    int i0 = s0, i1 = s1;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
      const int x = __shfl_up_sync(0xFFFFFFFFu, i0, k);
      const int y = __shfl_up_sync(0xFFFFFFFFu, i1, k);
      if (lane >= k) {
        i0 += x;
        i1 += y;
      }
    }
    const int b0 = run0 + i0 - s0, b1 = run1 + i1 - s1; // sums before my first pair
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) {
      const uint32_t p = pb + k;
      // (the dither state steps for every pixel position of my 8, also past the row end,
      //  so that `jump` lands on my pixels of the next step; nothing is stored there)
      const uint32_t pa = n_store_value(b0 + e[k], lut, rnd);
      const uint32_t pv = n_store_value(b1 + f[k], lut, rnd);
      if (p < npairs)
        o[p] = pa | (pv << 16);
    }
    if (lut)
      rnd = n_mulmod(rnd, jump);
    run0 += __shfl_sync(0xFFFFFFFFu, i0, 31);
    run1 += __shfl_sync(0xFFFFFFFFu, i1, 31);
  }
}

} // namespace rsb200
