// pentax.cuh -- K3P: reconstruction step of PentaxDecompressor::decompress
// (decompressors/PentaxDecompressor.cpp:158-176, paths relative to
// /root/reference/src/librawspeed).  The entropy decode is the shared multi-CTA
// path (ljpeg_ranges.cuh) run with the plain MSB bit source (DevScan::pump = 1,
// BitStreamerMSB: no stuffing, no markers); it leaves the differences, in stream
// order, in the linear int16 scratch buffer.  Here:
//
//   value(r, c)  = value(r-2, c) + d(r, c)      for c = 0, 1   (0 above row 0/1)
//   value(r, k)  = value(r, k-2) + d(r, k)      for k >= 2
//
// in plain int arithmetic, and every value must satisfy isIntN(value, 16)
// (adt/Bit.h:83-90: 0..65535) or the reference throws "decoded value out of
// bounds at col:row" at the FIRST such pixel in stream order.  Values before the
// first violation are exact, so the minimum (row, col) over all violations found
// here is that pixel; it is recorded in `oob[scan]` (0xFFFFFFFF = none).
//
//   k3p_column_kernel  one warp per (segment, row parity, column 0/1): int32 warp
//                      scan down every second row -> first two values of each row
//   k3p_row_kernel     one warp per row: per-parity int32 prefix sums along the row,
//                      range check, 32-bit stores (two pixels)
#pragma once

#include "ljpeg.cuh"

namespace rsb200 {

__device__ __forceinline__ uint32_t k3p_key(uint32_t row, uint32_t col) {
  return (row << 14) | col; // col < 8384 < 2^14
}

__global__ void k3p_column_kernel(const DevScan* __restrict__ scans,
                                  const uint32_t* __restrict__ scan_ids, int nscans,
                                  const uint16_t* __restrict__ diffs,
                                  uint16_t* __restrict__ colvals, uint32_t* __restrict__ oob) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int si = warp >> 2;
  if (si >= nscans)
    return;
  const uint32_t scan = scan_ids ? scan_ids[si] : (uint32_t)si;
  const DevScan& sc = scans[scan];
  if (sc.kind != 2)
    return;
  const uint32_t q = (warp >> 1) & 1u, c = warp & 1u; // row parity, column
  const int16_t* d = reinterpret_cast<const int16_t*>(diffs + sc.diff_offset) + c;
  uint16_t* cv = colvals + sc.col_offset + c;
  const uint32_t nj = (sc.rows > q) ? (sc.rows - q + 1) / 2 : 0;
  int run = 0;
  uint32_t first_bad = 0xFFFFFFFFu;
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
    if (j < nj) {
      cv[(uint64_t)r * 4] = (uint16_t)v;
      if (((uint32_t)v >> 16) != 0)
        first_bad = min(first_bad, k3p_key(r, c));
    }
    run = __shfl_sync(0xFFFFFFFFu, v, 31);
  }
  if (first_bad != 0xFFFFFFFFu)
    atomicMin(&oob[scan], first_bad);
}

__global__ void __launch_bounds__(K3_THREADS)
    k3p_row_kernel(const DevScan* __restrict__ scans, const K3RowRef* __restrict__ rows,
                   uint32_t nrows, const uint16_t* __restrict__ diffs,
                   const uint16_t* __restrict__ colvals, uint8_t* __restrict__ out,
                   uint32_t* __restrict__ oob) {
  const uint32_t wrow = (blockIdx.x * K3_THREADS + threadIdx.x) >> 5;
  if (wrow >= nrows)
    return;
  const K3RowRef ref = rows[wrow];
  const DevScan& sc = scans[ref.scan];
  if (sc.kind != 2)
    return;
  const int lane = threadIdx.x & 31;
  const uint32_t r = ref.row, npairs = sc.row_samples / 2;
  // pairs (even, odd sample) as 32-bit words; row starts are 4-byte aligned (even width)
  const uint32_t* d = reinterpret_cast<const uint32_t*>(diffs + sc.diff_offset +
                                                        (uint64_t)r * sc.row_samples);
  uint32_t* o = reinterpret_cast<uint32_t*>(out + sc.out_offset + (uint64_t)r * sc.out_pitch);
  const uint16_t* cv = colvals + sc.col_offset + (uint64_t)r * 4;
  int run0 = 0, run1 = 0;
  uint32_t first_bad = 0xFFFFFFFFu;
  constexpr uint32_t PER = 4; // pairs per lane and step
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
          a = (int)cv[0];
          b = (int)cv[1];
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
      if (p < npairs) {
        const int va = b0 + e[k], vb = b1 + f[k];
        if (((uint32_t)va >> 16) != 0)
          first_bad = min(first_bad, k3p_key(r, 2 * p));
        if (((uint32_t)vb >> 16) != 0)
          first_bad = min(first_bad, k3p_key(r, 2 * p + 1));
        o[p] = ((uint32_t)va & 0xFFFFu) | ((uint32_t)vb << 16);
      }
    }
    run0 += __shfl_sync(0xFFFFFFFFu, i0, 31);
    run1 += __shfl_sync(0xFFFFFFFFu, i1, 31);
  }
  if (first_bad != 0xFFFFFFFFu)
    atomicMin(&oob[ref.scan], first_bad);
}

} // namespace rsb200
