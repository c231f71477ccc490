// lookup_emu.cpp -- CPU replay of lookup_kernel (rawspeed_b200/csrc/lookup.cuh): the per-lane
// functions (lookup_core.h) and the job builder (lookup_host.h) are the library's own source;
// the loop mirrors the kernel's warp program (four rows per warp, lanes striding the groups of a
// row 32 apart, one modular jump between a lane's groups).  Test infrastructure.
#include "../../rawspeed_b200/csrc/lookup_host.h"

#include <cstring>
#include <vector>

using namespace rsb200;

namespace {
template <bool DITHER> void replay_quad(uint8_t* img, const LookupJobDev& j, uint32_t quad, const uint16_t* tables) {
  const uint32_t y0 = (quad - j.quad_begin) * SCALE_ROWS;
  uint8_t* const base = img + j.offset + (uint64_t)y0 * j.pitch;
  const uint16_t* const table = tables + (size_t)j.table * (DITHER ? 131072u : 65536u);
  const uint32_t iters = (j.ngroups + 31) / 32;
  const uint32_t jump = DITHER ? lut_powmod(248u) : 0u;
  for (int lane = 0; lane < 32; ++lane) {
    uint32_t st[SCALE_ROWS];
    for (int r = 0; r < SCALE_ROWS; ++r)
      st[r] = DITHER ? lut_mwc_state(j.width, y0 + r, 8u * (uint32_t)lane) : 0u;
    for (uint32_t it = 0; it < iters; ++it) {
      const uint32_t g = it * 32 + lane;
      if (g >= j.ngroups)
        continue;
      for (int r = 0; r < SCALE_ROWS; ++r) {
        if (y0 + r >= j.height)
          continue;
        uint8_t* p = base + (uint64_t)r * j.pitch + (uint64_t)g * 16;
        ScaleVec v;
        std::memcpy(v.w, p, 16);
        const ScaleVec o = lut_group<DITHER>(v, table, j.ncols, 8u * g, st[r]);
        std::memcpy(p, o.w, 16);
        if (DITHER)
          st[r] = lut_mwc_jump(st[r], 248u, jump);
      }
    }
  }
}
} // namespace

extern "C" int lookup_emu_run(uint8_t* img, const rsb200_lookup_job* jobs, int njobs, const uint16_t* tables,
                              int ntables, int dither, char* err, int errlen) {
  std::vector<LookupJobDev> hj((size_t)njobs);
  uint64_t quads = 0;
  for (int i = 0; i < njobs; ++i) {
    if (const char* why = lookup_build_job(jobs[i], ntables, (uint32_t)quads, &hj[i])) {
      std::strncpy(err, why, (size_t)errlen - 1);
      err[errlen - 1] = 0;
      return -1;
    }
    quads += lookup_job_quads(jobs[i]);
  }
  for (uint32_t quad = 0; quad < (uint32_t)quads; ++quad) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (hj[mid].quad_begin <= quad)
        lo = mid;
      else
        hi = mid - 1;
    }
    if (dither)
      replay_quad<true>(img, hj[lo], quad, tables);
    else
      replay_quad<false>(img, hj[lo], quad, tables);
  }
  return 0;
}

// the RSB200_LUT_SMEM=1 candidate (lookup_smem_kernel): persistent CTAs of 32 warps walking the
// quads with a grid stride, table staged once (here: used in place); plain lookup, one table
extern "C" int lookup_emu_run_smem(uint8_t* img, const rsb200_lookup_job* jobs, int njobs,
                                   const uint16_t* tables, int grid, char* err, int errlen) {
  std::vector<LookupJobDev> hj((size_t)njobs);
  uint64_t quads = 0;
  for (int i = 0; i < njobs; ++i) {
    if (const char* why = lookup_build_job(jobs[i], 1, (uint32_t)quads, &hj[i])) {
      std::strncpy(err, why, (size_t)errlen - 1);
      err[errlen - 1] = 0;
      return -1;
    }
    quads += lookup_job_quads(jobs[i]);
  }
  const uint32_t warps_per_cta = 32;
  std::vector<char> seen((size_t)quads, 0);
  for (int block = 0; block < grid; ++block)
    for (uint32_t warp = 0; warp < warps_per_cta; ++warp)
      for (uint32_t quad = (uint32_t)block * warps_per_cta + warp; quad < (uint32_t)quads;
           quad += (uint32_t)grid * warps_per_cta) {
        if (seen[quad]++)
          return -2; // a quad visited twice would be looked up twice
        int lo = 0, hi = njobs - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (hj[mid].quad_begin <= quad)
            lo = mid;
          else
            hi = mid - 1;
        }
        replay_quad<false>(img, hj[lo], quad, tables);
      }
  for (char c : seen)
    if (!c)
      return -3;
  return 0;
}

extern "C" uint32_t lookup_emu_mwc_direct(uint32_t width, uint32_t y, uint32_t x) {
  uint32_t v = (width + y * 13u) ^ 0x45694584u;
  while (x--)
    v = lut_mwc_step(v);
  return v;
}
extern "C" uint32_t lookup_emu_mwc_state(uint32_t width, uint32_t y, uint32_t x) { return lut_mwc_state(width, y, x); }
