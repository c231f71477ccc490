// ljpeg_clean.cuh -- K2C: unstuffing pre-pass of the one-thread-per-segment LJPEG
// path (K2T, ljpeg_thread.cuh), sm_100a.
//
// One CTA per entropy-coded segment streams its raw bytes through the same
// TMA-staged, cooperative unstuffer as the fused kernel (f_unstuff:
// BitStreamerJPEG.h:106-183 -- FF00 -> FF, the first FFxx ends the data, bytes past
// the buffer do not exist) and writes
//   * the clean data as big-endian 32-bit words (the serial decoders then need
//     no stuffing / marker / bounds logic at all), zero padded behind the end;
//   * one "anchor" per 256 raw bytes: the number of clean bytes that precede that
//     raw offset (lets a decoder map a clean offset back to a raw position for
//     `consumed`, the reference's BitStreamerJPEG::getStreamPosition());
//   * the clean length.
#pragma once

#include "ljpeg_fused.cuh"

namespace rsb200 {

// per thread-path segment, written by the host at plan creation
struct DevTScan {
  uint64_t clean_off;   // first word of this segment's clean data (in words)
  uint32_t cap_words;   // words that may be read (data + zero padding)
  uint32_t anchor_off;  // first anchor of this segment
  uint32_t n_anchor;    // anchors (one per 256 raw bytes from the 16-byte aligned base)
  uint32_t pad;
};
// ... and by K2C
struct DevTInfo {
  uint32_t clean_len; // data bytes of the segment
  uint32_t marker;    // 1: an end marker was found; 0: the data ran to the end of the buffer
};

constexpr uint32_t T_ANCHOR_SHIFT = 8; // one anchor per 256 raw bytes
constexpr uint32_t T_PAD_WORDS = 8;    // zero words behind the data

__global__ void __launch_bounds__(F_NT, 5)
    k2_clean_kernel(const uint8_t* __restrict__ in, uint64_t in_total,
                    const DevScan* __restrict__ scans, const uint32_t* __restrict__ scan_ids,
                    const DevTScan* __restrict__ tscans, uint32_t* __restrict__ clean,
                    uint32_t* __restrict__ anchors, DevTInfo* __restrict__ infos) {
  extern __shared__ __align__(128) uint8_t c_smem_raw[];
  FusedShared& sh = *reinterpret_cast<FusedShared*>(c_smem_raw);
  const int tid = threadIdx.x;
  const uint32_t scan_idx = scan_ids[blockIdx.x];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&scans[scan_idx]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sh.sc);
    for (int i = tid; i < (int)(sizeof(DevScan) / 4); i += F_NT)
      dst[i] = src[i];
  }
  if (tid == 0) {
    mbar_init(&sh.bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const DevScan& sc = sh.sc;
  const DevTScan ts = tscans[blockIdx.x];
  const uint64_t abase = sc.in_offset & ~15ull;
  const uint32_t skew = (uint32_t)(sc.in_offset - abase);
  const uint32_t limit = skew + sc.in_size;
  FStream st;
  st.gbase = in + abase;
  st.limit = limit;
  st.skew = skew;
  st.readable = ((in_total + 15) & ~15ull) - abase;
  st.chunk_begin = 0;
  st.chunk_end = (limit + F_RAW - 1) / F_RAW;
  st.plain = sc.pump != 0;
  st.pending = true;
  st.pending_par = 0;
  if (tid == 0) {
    FusedCarry c;
    c.pos = c.sym = 0;
    c.tail_len = 0;
    c.tail_raw = skew;
    c.ubytes = 0;
    c.prev_ff = 0;
    c.ended = 0;
    c.leftover = c.proc = c.status = 0;
    c.pc01 = c.pc23 = c.col01 = c.col23 = c.rb01 = c.rb23 = 0;
    sh.cy = c;
    f_issue_chunk(sh, st, 0);
  }
  __syncthreads();
  uint32_t* cw = clean + ts.clean_off;
  uint32_t* anc = anchors + ts.anchor_off;
  uint32_t nchunks = 0;
  for (uint32_t chunk = 0;; ++chunk) {
    const FusedCarry cy = sh.cy;
    nchunks = chunk;
    if (cy.ended)
      break;
    mbar_wait(&sh.bar, chunk & 1);
    st.pending = false;
    const FChunk co = f_unstuff(sh, st, cy, chunk); // clean bytes [0, co.len) in sh.ub
    // anchors of this chunk's 256-byte raw blocks (sh.anchor = index in ub of the
    // first clean byte each 32-byte subsequence produced)
    if ((tid & 7) == 0) {
      const uint32_t a = ((chunk * (uint32_t)F_RAW) >> T_ANCHOR_SHIFT) + (tid >> 3);
      if (a < ts.n_anchor)
        anc[a] = cy.ubytes + sh.anchor[tid];
    }
    // whole words out (cy.ubytes is a multiple of 4); the last chunk also writes
    // its partial word (zero filled by f_unstuff) and the padding
    const uint32_t tail = co.final_chunk ? 0u : (co.len & 3u);
    const uint32_t keep = co.len - tail;
    uint32_t nwords = keep >> 2, ndata = keep >> 2;
    if (co.final_chunk) {
      ndata = (co.len + 3u) >> 2; // (bytes behind the data in the last word are zero)
      nwords = ndata + T_PAD_WORDS;
    }
    const uint32_t w0 = cy.ubytes >> 2;
    for (uint32_t i = tid; i < nwords; i += F_NT)
      if (w0 + i < ts.cap_words)
        cw[w0 + i] = i < ndata ? sh.ub[i] : 0u;
    uint32_t tailbyte = 0;
    if ((uint32_t)tid < tail)
      tailbyte = reinterpret_cast<uint8_t*>(sh.ub)[(keep + tid) ^ 3u];
    __syncthreads();
    if ((uint32_t)tid < tail)
      reinterpret_cast<uint8_t*>(sh.ub)[tid ^ 3u] = (uint8_t)tailbyte;
    if (tid == 0) {
      FusedCarry& c2 = sh.cy;
      c2.tail_len = tail;
      c2.ubytes = cy.ubytes + keep;
      c2.prev_ff = !st.plain && (sh.last_raw_byte == 0xFFu) &&
                   ((chunk + 1) * (uint32_t)F_RAW - 1 < limit) &&
                   ((chunk + 1) * (uint32_t)F_RAW - 1 >= skew);
      c2.ended = co.final_chunk ? 1u : 0u;
      if (co.final_chunk) {
        infos[blockIdx.x].clean_len = cy.ubytes + co.len;
        infos[blockIdx.x].marker = co.mpos != 0xFFFFFFFFu ? 1u : 0u;
      }
    }
    __syncthreads();
  }
  // anchors of raw blocks behind the last processed chunk (the data ended at a
  // marker before them): no clean offset maps there
  for (uint32_t a = ((nchunks * (uint32_t)F_RAW) >> T_ANCHOR_SHIFT) + tid; a < ts.n_anchor; a += F_NT)
    anc[a] = 0xFFFFFFFFu;
  if (st.pending)
    mbar_wait(&sh.bar, st.pending_par);
}

} // namespace rsb200
