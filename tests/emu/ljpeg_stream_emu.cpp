// ljpeg_stream_emu.cpp -- CPU replay of k2_stream_kernel (rawspeed_b200/csrc/ljpeg_stream.cuh): the
// kernel body itself, compiled by g++ against tests/emu/cuda_emu.h (one fiber per CUDA thread),
// driven with the descriptors the plan builder (ljpeg_host.h) produces.  Test infrastructure: it
// checks the thread-local unstuffer, the ring bookkeeping and the end-of-stream logic where there
// is no GPU; parity of the real kernel is the GPU tests' job.
#include "cuda_emu.h"

#ifndef RSB200_EMU_WIDE
#define RSB200_EMU_WIDE false
#endif

#include "../../rawspeed_b200/csrc/ljpeg_stream.cuh"
#include "../../rawspeed_b200/csrc/ljpeg_host.h"

#include <vector>

using namespace rsb200;

// any_mode 0: a lane fills when it is low itself; 1: every lane with room fills at every unit
extern "C" int stream_emu_run(const uint8_t* in, uint64_t in_total, const rsb200_huff_table* tables,
                              int ntables, const rsb200_ljpeg_scan* scans, int nscans, uint8_t* out,
                              rsb200_scan_result* results, uint32_t* redo_out, int reverse, int any_mode) {
  std::vector<DevTable> ht((size_t)ntables);
  for (int i = 0; i < ntables; ++i)
    if (!build_dev_table(tables[i], ht[(size_t)i]))
      return -2;
  std::vector<DevScan> ds((size_t)nscans);
  std::vector<uint32_t> ids((size_t)nscans);
  for (int i = 0; i < nscans; ++i) {
    if (!ljpeg_scan_to_dev(scans[i], ntables, ds[(size_t)i]))
      return -3;
    const DevScan& d = ds[(size_t)i];
    // the plan's thread_eligible(): whole 8-sample units, aligned 128-bit stores
    if (!(d.kind == 0 && d.pump == 0 && d.mcu_h == 1 && (d.group == 1 || d.group == 2 || d.group == 4) &&
          (d.row_samples & 7u) == 0 && ((d.out_offset | d.out_pitch) & 15u) == 0 && (d.out_x & 7u) == 0))
      return -1;
    ids[(size_t)i] = (uint32_t)i | 0x80000000u;
  }
  // the product's input contract: base 16-byte aligned, readable up to the next 16-byte boundary
  const uint64_t padded = (in_total + 15) & ~15ull;
  std::vector<uint8_t> buf(padded + 64, 0xA5);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(buf.data()) + 15) & ~(uintptr_t)15);
  memcpy(base, in, in_total);
  std::vector<DevResult> res((size_t)nscans);
  std::vector<uint32_t> redo((size_t)nscans, 7u);
  g_emu_any_mode = any_mode;
  const unsigned nblocks = (unsigned)((nscans + T_NT - 1) / T_NT);
  for (unsigned b = 0; b < nblocks; ++b)
    cuemu::run_cta(b, nblocks, T_NT, std::max(sizeof(StreamShared), stream_smem_bytes(T_MAXTAB)), reverse != 0, [&](uint8_t* smem) {
      StreamShared& sh = *reinterpret_cast<StreamShared*>(smem);
      stream_entry<RSB200_EMU_WIDE>(sh, base, in_total, ds.data(), ht.data(), ntables, out, res.data(), ids.data(),
                   (uint32_t)nscans, redo.data(), (any_mode & 1) != 0);
    });
  for (int i = 0; i < nscans; ++i) {
    results[i].status = res[(size_t)i].status;
    results[i].consumed = res[(size_t)i].consumed;
    if (redo_out)
      redo_out[i] = redo[(size_t)i];
  }
  return 0;
}
