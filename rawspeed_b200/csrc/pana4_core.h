// pana4_core.h -- one PanasonicV4 pixel packet (16 bytes -> 14 pixels), written so that the
// same source compiles as device code (pana.cuh, VER == 4) and as plain C++ (the CPU replay
// in tests/emu/pana4_emu.cpp).
//
// Reference: PanasonicV4Decompressor::processPixelPacket
//   decompressors/PanasonicV4Decompressor.cpp:171-214, ProxyStream::getBits :164-168.
//
// ProxyStream reads a block backwards in 16-byte steps: with c bits of a packet consumed,
// getBits(n) returns bits [128-c-n, 128-c) of the packet taken as a 128-bit little-endian
// number (byte = (vbits >> 3) ^ 0x3ff0 walks the packet's bytes from the last to the first).
// A packet always consumes exactly 128 bits: 14 x 8, four 2-bit shifts (before pixels 2, 5, 8,
// 11) and exactly one 4-bit field per colour (at the colour's first non-zero byte, at pixel
// 12 / 13 at the latest), so packets are independent units; what varies inside a packet is
// only WHERE the two 4-bit fields sit.  The packet is kept top-aligned in two 64-bit
// registers and shifted left by the (compile-time) width of every field read.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define RS4_HD __host__ __device__ __forceinline__
#else
#define RS4_HD inline
#endif

namespace rsb200 {

struct Pana4Bits {
  uint64_t hi, lo; // the unread bits, top-aligned
};

template <int N> RS4_HD uint32_t pana4_get(Pana4Bits& b) {
  const uint32_t v = (uint32_t)(b.hi >> (64 - N));
  b.hi = (b.hi << N) | (b.lo >> (64 - N));
  b.lo <<= N;
  return v;
}

template <int P>
RS4_HD void pana4_pixel(Pana4Bits& b, int& sh, int (&pred)[2], int (&nonz)[2], uint32_t (&px)[14],
                        uint32_t& zeros) {
  constexpr int cc = P & 1;
  if (P % 3 == 2) // u == 2 (:183-186): extractHighBits(4U, getBits(2), 3) = 4 >> (3 - n)
    sh = 4 >> (3 - (int)pana4_get<2>(b));
  if (nonz[cc]) {
    const int j = (int)pana4_get<8>(b);
    if (j) {
      pred[cc] -= 0x80 << sh;
      if (pred[cc] < 0 || sh == 4)
        pred[cc] &= (1 << sh) - 1; // ~(-(1 << sh))
      pred[cc] += j << sh;
    }
  } else {
    nonz[cc] = (int)pana4_get<8>(b);
    if (nonz[cc] || P > 11)
      pred[cc] = nonz[cc] << 4 | (int)pana4_get<4>(b);
  }
  px[P] = (uint32_t)pred[cc] & 0xFFFFu;
  if (pred[cc] == 0)
    zeros |= 1u << P;
}

template <int P>
RS4_HD void pana4_pixels(Pana4Bits& b, int& sh, int (&pred)[2], int (&nonz)[2], uint32_t (&px)[14],
                         uint32_t& zeros) {
  if constexpr (P < 14) {
    pana4_pixel<P>(b, sh, pred, nonz, px, zeros);
    pana4_pixels<P + 1>(b, sh, pred, nonz, px, zeros);
  }
}

// w[0..3]: the packet's 16 bytes as little-endian words; returns the mask of pixels that are 0
RS4_HD uint32_t pana4_packet(const uint32_t (&w)[4], uint32_t (&px)[14]) {
  Pana4Bits b;
  b.hi = ((uint64_t)w[3] << 32) | w[2];
  b.lo = ((uint64_t)w[1] << 32) | w[0];
  int sh = 0, pred[2] = {0, 0}, nonz[2] = {0, 0};
  uint32_t zeros = 0;
  pana4_pixels<0>(b, sh, pred, nonz, px, zeros);
  return zeros;
}

// byte offset, inside the image's data, of rearranged byte `o` of block `blk`: the two
// sections of a block are swapped (ProxyStream::parseBlock :136-159), split == 0 = no swap
RS4_HD uint64_t pana4_src(uint32_t blk, uint32_t o, uint32_t split) {
  return (uint64_t)blk * 0x4000u + (split ? ((o + split) & 0x3FFFu) : o);
}

} // namespace rsb200
