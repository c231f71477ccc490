// ljpeg_tile_emu.cpp -- CPU replay of k2_tile_kernel<R> (rawspeed_b200/csrc/ljpeg_tile.cuh): the
// kernel body itself, compiled by g++ against tests/emu/cuda_emu.h (one fiber per CUDA thread,
// barriers / shuffles / mbarrier waits are yield points), driven with the descriptors the plan
// builder (ljpeg_host.h) produces.  Test infrastructure: it checks the kernel's arithmetic,
// indexing and barrier structure where there is no GPU (forward and reverse thread order);
// parity of the real kernel is the GPU tests' job.
#include "cuda_emu.h"

#include "../../rawspeed_b200/csrc/ljpeg_tile.cuh"
#include "../../rawspeed_b200/csrc/ljpeg_host.h"

#include <vector>

using namespace rsb200;

namespace {
template <int R>
int run(const uint8_t* in, uint64_t in_total, const rsb200_huff_table* tables, int ntables,
        const rsb200_ljpeg_scan* scans, int nscans, uint8_t* out, rsb200_scan_result* results,
        int reverse, int preroll_override, int npieces_override) {
  using G = TileGeom<R>;
  std::vector<DevTable> ht((size_t)ntables);
  for (int i = 0; i < ntables; ++i)
    if (!build_dev_table(tables[i], ht[(size_t)i]))
      return -2;
  std::vector<DevScan> ds((size_t)nscans);
  std::vector<DevTileParam> pr((size_t)nscans);
  for (int i = 0; i < nscans; ++i) {
    if (!ljpeg_scan_to_dev(scans[i], ntables, ds[(size_t)i]))
      return -3;
    if (!tile_eligible(ds[(size_t)i], G::MIN_RS))
      return -1;
    tile_params(ds[(size_t)i], G::NPIECE, G::DCAP, preroll_override, pr[(size_t)i].npieces,
                pr[(size_t)i].preroll);
    if (npieces_override > 0)
      pr[(size_t)i].npieces = (uint32_t)std::min(npieces_override, G::NPIECE);
  }
  // the product's input contract: base 16-byte aligned, readable up to the next 16-byte boundary
  const uint64_t padded = (in_total + 15) & ~15ull;
  std::vector<uint8_t> buf(padded + 64, 0xA5);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(buf.data()) + 15) & ~(uintptr_t)15);
  memcpy(base, in, in_total);
  std::vector<DevResult> res((size_t)nscans);
  for (int i = 0; i < nscans; ++i) {
    const DevTileParam* prp = pr.data() + i - i; // params are indexed by blockIdx.x
    (void)prp;
    cuemu::run_cta((unsigned)i, (unsigned)nscans, TL_NT, sizeof(TileShared<R>), reverse != 0,
                   [&](uint8_t* smem) {
                     TileShared<R>& sh = *reinterpret_cast<TileShared<R>*>(smem);
                     tile_entry<R>(sh, base, in_total, ds.data(), ht.data(), out, res.data(), nullptr,
                                   pr.data());
                   });
  }
  for (int i = 0; i < nscans; ++i) {
    results[i].status = res[(size_t)i].status;
    results[i].consumed = res[(size_t)i].consumed;
  }
  return 0;
}
} // namespace

extern "C" int tile_emu_run(const uint8_t* in, uint64_t in_total, const rsb200_huff_table* tables,
                            int ntables, const rsb200_ljpeg_scan* scans, int nscans, uint8_t* out,
                            rsb200_scan_result* results, int R, int reverse, int preroll_override,
                            int npieces_override) {
  if (R == 1)
    return run<1>(in, in_total, tables, ntables, scans, nscans, out, results, reverse,
                  preroll_override, npieces_override);
  if (R == 2)
    return run<2>(in, in_total, tables, ntables, scans, nscans, out, results, reverse,
                  preroll_override, npieces_override);
  return -4;
}

extern "C" int tile_emu_smem_bytes(int R) {
  return R == 1 ? (int)sizeof(TileShared<1>) : (int)sizeof(TileShared<2>);
}
