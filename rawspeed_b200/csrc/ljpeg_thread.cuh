// ljpeg_thread.cuh -- K2T: LJPEG tile decode for LARGE batches, one THREAD per
// entropy-coded segment (DNG tile / restart interval), sm_100a.
//
// Same semantics as k2_fused_kernel (see ljpeg.cuh / ljpeg_fused.cuh for the
// reference citations: BitStreamerJPEG.h:106-183, PrefixCodeLUTDecoder.h:172-216,
// AbstractPrefixCodeDecoder.h:43-76, LJpegDecompressor.cpp:184-339).
//
// Why a second kernel: a batch of frames holds 10^4..10^5 independent segments
// (726 tiles per 45 MP frame).  With that many streams the serial dependency of
// a Huffman stream is no longer a problem -- each thread simply IS the
// reference's sequential decoder (64-bit bit cache, fill(32) before every
// symbol, 4 clean bytes per refill, FF00 unstuffing, first FFxx ends the data)
// and the machine is kept busy by the number of streams.  No synchronisation
// rounds, no second decode pass, no shared-memory staging: ~25 instructions
// per sample instead of ~160 issue slots in the block-per-segment kernel, at
// the price of a fixed latency (one tile's serial decode, ~2 ms), which is why
// the plan only takes this path when the launch holds enough segments.
//
//   * input: every thread streams its segment with 128-bit loads, one 16-byte
//     block ahead (double buffered in registers), raw words kept in a small
//     register FIFO so the 4-byte refill is one funnel shift + byte swap;
//   * Huffman LUTs of the plan (<= 4 tables) in shared memory;
//   * predictor 1 in registers (mod 2^16), first MCU of a row predicted from
//     the first MCU of the previous row; 8 samples are packed and written with
//     one 128-bit store (rows of a tile are 16-byte aligned);
//   * `consumed` falls out of the refill cadence (it IS the reference's cadence).
#pragma once

#include "ljpeg.cuh"

namespace rsb200 {

constexpr int T_NT = 64;       // threads (= segments) per CTA
constexpr int T_MAXTAB = 4;    // plan tables staged in shared memory

struct ThreadShared {
  DevTable tab[T_MAXTAB];
};

__host__ __device__ inline size_t thread_smem_bytes(int ntab) {
  return sizeof(DevTable) * (size_t)ntab;
}

__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---- refill of the bit cache, everything that is not "4 plain bytes" ----
// One aligned raw word `w` (bytes k0..kend-1 of it belong to the segment), `wn` =
// the word after it (look-ahead for an FF in the last byte; bytes past the limit
// already zeroed).  flags: bit 0 = the first byte is the stuffing 00 of an FF that
// ended the previous word.  Returns .x = the clean bytes (right aligned), .y =
// number of clean bytes | flags' << 8 (bit 0 as above, bit 1 = end marker found) |
// bytes not delivered << 16 (stuffing, bytes outside the segment, marker and after).
__device__ __noinline__ uint2 t_refill_word(uint32_t w, uint32_t wn, uint32_t k0, uint32_t kend,
                                            uint32_t flags) {
  uint32_t acc = 0, n = 0, out_flags = 0;
  uint64_t v = ((uint64_t)wn << 32) | w;
  v >>= 8 * k0;
  uint32_t k = k0;
  if ((flags & 1u) && k < kend) { // stuffing byte of the previous word's FF
    v >>= 8;
    ++k;
  }
  while (k < kend) {
    const uint32_t c0 = (uint32_t)v & 0xFFu;
    if (c0 != 0xFFu) {
      acc = (acc << 8) | c0;
      ++n;
      v >>= 8;
      ++k;
      continue;
    }
    const uint32_t c1 = ((uint32_t)v >> 8) & 0xFFu; // (past the limit: 0 -> "FF 00")
    if (c1 != 0u) {
      out_flags |= 2u; // FF xx: end of the data, the FF is not data
      break;
    }
    acc = (acc << 8) | 0xFFu;
    ++n;
    if (k + 1 < 4u) {
      v >>= 16;
      k += 2;
    } else {
      out_flags |= 1u; // its 00 is the first byte of the next word
      ++k;
    }
  }
  return make_uint2(acc, n | (out_flags << 8) | ((4u - n) << 16));
}

// The JPEG bit source (BitStreamerJPEG.h:106-183), one per thread: 64-bit cache
// refilled by whole aligned raw words whenever fewer than 32 bits are left.
// (The reference refills 4 DATA bytes at a time; the schedule differs but the bit
// sequence is the same, and `consumed` is rebuilt from the bit offset of the
// last symbol -- see t_stream_position.)
struct TSrc {
  uint32_t hi, lo;   // unread bits, MSB aligned in hi:lo
  int nbits;         // number of unread bits (real + fake)
  uint32_t fake;     // zero bits supplied after the end of the data
  uint32_t rp;       // raw offset (from gbase, multiple of 4) of the next unread word
  uint32_t dropped;  // raw bytes before rp that are not data (lead-in, stuffing)
  uint32_t w0;              // raw word at rp
  uint32_t n0, n1, n2, n3;  // the following words, nleft of them valid
  uint32_t m0, m1, m2, m3;  // the 16-byte block after those
  int nleft;
  uint32_t next_blk;  // index of the block to fetch after m
  uint32_t nblk;      // blocks that may be read (16-byte padded input buffer)
  const uint4* blocks;
  uint32_t limit;     // valid raw bytes from gbase
  uint32_t fast_end;  // rp below this: the word at rp lies inside the segment
  uint32_t flags;     // bit 0: byte at rp is a pending stuffing 00; bit 1: data ended
  uint32_t k0;        // bytes to skip in the next word (segment start inside a word)
  // two (rp, dropped, flags) snapshots, >= 64 raw bytes apart, for t_stream_position
  uint32_t a0_rp, a0_dr, a1_rp, a1_dr;

  __device__ __forceinline__ uint4 fetch(uint32_t b) const {
    if (b < nblk)
      return ldg_stream_v4(blocks + b);
    return make_uint4(0u, 0u, 0u, 0u);
  }
  __device__ __forceinline__ void pop() {
    w0 = n0;
    n0 = n1;
    n1 = n2;
    n2 = n3;
    if (--nleft == 0) {
      n0 = m0; n1 = m1; n2 = m2; n3 = m3;
      nleft = 4;
      const uint4 q = fetch(next_blk++);
      m0 = q.x; m1 = q.y; m2 = q.z; m3 = q.w;
    }
  }
  __device__ __forceinline__ void init(const uint8_t* gb, uint32_t skew, uint32_t lim,
                                       uint64_t readable) {
    blocks = reinterpret_cast<const uint4*>(gb);
    limit = lim;
    nblk = (uint32_t)(readable >> 4);
    hi = lo = 0;
    nbits = 0;
    fake = 0;
    flags = 0;
    const uint4 a = fetch(0), b = fetch(1);
    n0 = a.x; n1 = a.y; n2 = a.z; n3 = a.w;
    m0 = b.x; m1 = b.y; m2 = b.z; m3 = b.w;
    nleft = 4;
    next_blk = 2;
    w0 = 0;
#pragma unroll 1
    for (uint32_t k = 0; k < 1u + (skew >> 2); ++k)
      pop();
    rp = skew & ~3u;
    k0 = skew & 3u;
    dropped = 0;
    fast_end = lim & ~3u; // rp < fast_end: the whole word lies inside the segment
    a0_rp = a1_rp = 0xFFFFFFFFu; // (no snapshot: walk from the segment start)
    a0_dr = a1_dr = 0;
  }
  // one raw word -> cache
  __device__ __forceinline__ void refill() {
    const uint32_t w = w0;
    uint32_t be = __byte_perm(w, 0, 0x0123), nb = 32;
    if ((rp & 63u) == 0u && (flags | k0) == 0u) { // snapshot: rp is a data byte boundary
      a1_rp = a0_rp;
      a1_dr = a0_dr;
      a0_rp = rp;
      a0_dr = dropped;
    }
    // any byte == FF  <=>  any byte of ~w == 0
    const uint32_t ff = (~w - 0x01010101u) & w & 0x80808080u;
    if (rp < fast_end && (ff | flags | k0) == 0u) {
      pop();
      rp += 4;
    } else if (flags & 2u) {
      be = 0; // data ended: zero bits from here on
      fake += 32;
    } else {
      const uint32_t kend = rp >= limit ? 0u : min(4u, limit - rp);
      uint32_t wn = n0;
      if (rp + 4u >= limit)
        wn = 0;
      else if (rp + 8u > limit)
        wn &= 0xFFFFFFFFu >> (8u * (rp + 8u - limit));
      const uint2 r = t_refill_word(w, wn, k0, kend, flags);
      const uint32_t n = r.y & 0xFFu;
      flags = (r.y >> 8) & 3u;
      if (kend < 4u)
        flags |= 2u; // ran off the end of the segment
      dropped += 4u - n; // (rp - dropped keeps counting the data bytes delivered)
      rp += 4;
      k0 = 0;
      be = n ? r.x << (32u - 8u * n) : 0u;
      if (flags & 2u) {
        fake += 32u - 8u * n; // zero bits complete this refill; they are missing data
      } else {
        nb = 8u * n;
        pop();
      }
    }
    hi |= __funnelshift_rc(be, 0u, (uint32_t)nbits); // be >> nbits (nbits <= 32)
    lo = __funnelshift_lc(0u, be, 32u - (uint32_t)nbits); // (lo holds no unread bits while nbits <= 32)
    nbits += (int)nb;
  }
  // at least 32 bits in the cache (one symbol is at most 32 bits long)
  __device__ __forceinline__ void fill() {
#pragma unroll 1
    while (nbits < 32)
      refill();
  }
  __device__ __forceinline__ void skip(uint32_t n) { // n <= 32
    hi = __funnelshift_lc(lo, hi, n);
    asm("shl.b32 %0, %0, %1;" : "+r"(lo) : "r"(n)); // (PTX shifts clamp: n = 32 -> 0)
    nbits -= (int)n;
  }
};

// BitStreamerJPEG::getStreamPosition() of the reference after the last symbol, whose
// first bit is data bit T of the segment: the reference has done R = T/32 + 1 (+1 if
// T % 32 != 0) refills of 4 data bytes by then (BitStreamer::fill(32) before every
// symbol, BitStreamer.h:216-229), so its position is the raw offset after 4R data
// bytes, or the end marker if that comes first.  Walks forward from a snapshot.
__device__ __noinline__ uint32_t t_stream_position(const uint8_t* gbase, uint32_t limit,
                                                   uint32_t skew, uint32_t T, uint32_t a_rp,
                                                   uint32_t a_dr) {
  const uint32_t R = (T >> 5) + 1u + ((T & 31u) ? 1u : 0u);
  const uint32_t need = 4u * R;
  uint32_t rawp = skew, c = 0;
  if (a_rp != 0xFFFFFFFFu) {
    rawp = a_rp;
    c = a_rp - (skew & ~3u) - a_dr;
  }
  auto byte_at = [&](uint32_t q) { return q < limit ? (uint32_t)__ldg(gbase + q) : 0u; };
  while (c < need) {
    if (byte_at(rawp) == 0xFFu) {
      if (byte_at(rawp + 1) != 0u)
        break; // marker: the position stays on it
      rawp += 2;
    } else {
      rawp += 1;
    }
    ++c;
  }
  return rawp - skew;
}

// symbols the LUT does not resolve (T.81 F.16 walk); .x = difference, .y = bits
// consumed | bad-code flag << 31
__device__ __noinline__ uint2 t_slow_symbol(const DevTable* t, uint32_t x) {
  const SymLen s = decode_sym(t, x);
  return make_uint2((uint32_t)sym_diff(s, x),
                    (uint32_t)s.total | (s.codelen == 0 ? 0x80000000u : 0u));
}

// One difference at the top of window x with the LUT at shared address lutb
// (slow path: table t).  Same arithmetic as f_decode_diff.
__device__ __forceinline__ uint32_t t_decode_diff(const DevTable* t, uint32_t lutb, uint32_t x,
                                                  uint32_t& tl, uint32_t& bad) {
  const uint32_t e =
      lds_u16<0>(mad_hi(x & ~((1u << (32 - LUT_BITS)) - 1u), 1u << (LUT_BITS + 1), lutb));
  tl = e >> 10;
  if (e == 0) { // code longer than the LUT, SSSS = 16, or corrupt
    const uint2 r = t_slow_symbol(t, x);
    tl = r.y & 0xFFu;
    bad |= r.y >> 31;
    return r.x;
  }
  const uint32_t tt = __funnelshift_l(0u, x, e);
  const uint32_t f = (uint32_t)((int32_t)~tt >> 31);
  return __funnelshift_l(tt, f, e >> 5) - f;
}

// two samples (components ca, cb of the MCU) -> one output word
#define T_PAIR(ca, cb, word)                                                    \
  do {                                                                          \
    uint32_t& tl_ = last_tl;                                                    \
    bs.fill();                                                                  \
    const uint32_t da_ = t_decode_diff(tabp[ca], lutb[ca], bs.hi, tl_, bad);    \
    bs.skip(tl_);                                                               \
    pred[ca] += da_;                                                            \
    const uint32_t va_ = pred[ca];                                              \
    bs.fill();                                                                  \
    const uint32_t db_ = t_decode_diff(tabp[cb], lutb[cb], bs.hi, tl_, bad);    \
    bs.skip(tl_);                                                               \
    pred[cb] += db_;                                                            \
    word = __byte_perm(va_, pred[cb], 0x5410);                                  \
  } while (0)

template <int G>
__device__ __forceinline__ void
thread_body(const ThreadShared& sh, const DevScan* __restrict__ scp,
            const uint8_t* __restrict__ in, uint64_t in_total, uint8_t* __restrict__ out,
            DevResult* __restrict__ res) {
  const uint64_t in_offset = scp->in_offset;
  const uint64_t abase = in_offset & ~15ull;
  const uint32_t skew = (uint32_t)(in_offset - abase);
  const uint32_t limit = skew + scp->in_size;
  const uint64_t readable = ((in_total + 15) & ~15ull) - abase;
  TSrc bs;
  bs.init(in + abase, skew, limit, readable);

  uint32_t lutb[G];
  const DevTable* tabp[G];
  uint32_t rowstart[G], pred[G];
#pragma unroll
  for (int c = 0; c < G; ++c) {
    tabp[c] = &sh.tab[scp->table_idx[scp->table_of[c]]];
    lutb[c] = smem_u32(tabp[c]->lut);
    rowstart[c] = scp->init_pred[c];
  }
  const uint32_t rows = scp->rows;
  const uint32_t units = scp->row_samples >> 3; // row_samples is a multiple of 8
  const uint32_t store_w = scp->store_w;
  const uint32_t out_pitch = scp->out_pitch;
  uint8_t* orow = out + scp->out_offset + (uint64_t)scp->out_y * out_pitch + 2ull * scp->out_x;
  uint32_t bad = 0, last_tl = 0;

  for (uint32_t r = 0; r < rows; ++r) {
#pragma unroll
    for (int c = 0; c < G; ++c)
      pred[c] = rowstart[c];
    for (uint32_t u = 0; u < units; ++u) {
      // 8 samples = 4 words, collected in a rotating 128-bit register
      uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
      if (G == 4) {
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
          uint32_t wa, wb;
          T_PAIR(0, 1, wa);
          T_PAIR(2, 3, wb);
          o0 = o2;
          o1 = o3;
          o2 = wa;
          o3 = wb;
        }
      } else {
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          uint32_t w;
          T_PAIR(0, G - 1, w);
          o0 = o1;
          o1 = o2;
          o2 = o3;
          o3 = w;
        }
      }
      if (u == 0) { // the first MCU of the row predicts the first MCU of the next row
        rowstart[0] = o0 & 0xFFFFu;
        if (G >= 2)
          rowstart[1] = o0 >> 16;
        if (G == 4) {
          rowstart[2] = o1 & 0xFFFFu;
          rowstart[3] = o1 >> 16;
        }
      }
      const uint32_t s = u << 3;
      if (s + 8 <= store_w) {
        stg_cs_v4(orow + 16ull * u, make_uint4(o0, o1, o2, o3));
      } else if (s < store_w) {
        uint16_t* o16 = reinterpret_cast<uint16_t*>(orow) + s;
        const uint32_t ow[4] = {o0, o1, o2, o3};
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (s + k < store_w)
            o16[k] = (uint16_t)(ow[k >> 1] >> (16 * (k & 1)));
      }
    }
    orow += out_pitch;
  }
  // status: a needed symbol used bits that are not there (DESIGN.md "known deviations");
  // consumed: no refill follows the last symbol, so rp is the reference's stream position
  const bool over = (uint32_t)bs.nbits < bs.fake;
  res->status = bad ? 1u : (over ? 2u : 0u);
  {
    // data bits consumed in total, minus the last symbol = bit offset of the last symbol
    const uint32_t clean = bs.rp - (skew & ~3u) - bs.dropped;
    const uint32_t T = 8u * clean + bs.fake - (uint32_t)bs.nbits - last_tl;
    // the older snapshot is guaranteed to lie before the 4R-th data byte
    res->consumed = t_stream_position(in + abase, limit, skew, T, bs.a1_rp, bs.a1_dr);
  }
}
#undef T_PAIR

__global__ void __launch_bounds__(T_NT)
    k2_thread_kernel(const uint8_t* __restrict__ in, uint64_t in_total,
                     const DevScan* __restrict__ scans, const DevTable* __restrict__ tables,
                     int ntab, uint8_t* __restrict__ out, DevResult* __restrict__ results,
                     const uint32_t* __restrict__ scan_ids, uint32_t nids) {
  extern __shared__ __align__(16) uint8_t t_smem_raw[];
  ThreadShared& sh = *reinterpret_cast<ThreadShared*>(t_smem_raw);
  const int tid = threadIdx.x;
  {
    const uint4* src = reinterpret_cast<const uint4*>(tables);
    uint4* dst = reinterpret_cast<uint4*>(sh.tab);
    const int n = ntab * (int)(sizeof(DevTable) / 16);
    for (int i = tid; i < n; i += T_NT)
      dst[i] = src[i];
  }
  __syncthreads();
  const uint32_t id = blockIdx.x * T_NT + tid;
  if (id >= nids)
    return;
  const uint32_t scan_idx = scan_ids[id];
  const DevScan* scp = scans + scan_idx;
  DevResult* res = results + scan_idx;
  const uint32_t G = scp->group;
  if (G == 1)
    thread_body<1>(sh, scp, in, in_total, out, res);
  else if (G == 2)
    thread_body<2>(sh, scp, in, in_total, out, res);
  else
    thread_body<4>(sh, scp, in, in_total, out, res);
}

} // namespace rsb200
