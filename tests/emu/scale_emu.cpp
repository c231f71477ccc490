// scale_emu.cpp -- CPU replay of scale_kernel (rawspeed_b200/csrc/scale.cuh): the same
// per-lane functions (scale_core.h) and the same job builder (scale_host.h), driven by a
// loop that mirrors the kernel's warp program statement by statement (lanes run one after
// the other between the points where the kernel has __syncwarp()).  Test infrastructure:
// it checks the kernel's arithmetic and indexing where there is no GPU; the GPU test
// (tests/test_gpu_scale.py) is what establishes parity of the real kernel.
#include "../../rawspeed_b200/csrc/scale_host.h"

#include <cstring>
#include <vector>

using namespace rsb200;

namespace {
ScaleVec ld(const uint8_t* p) {
  ScaleVec v;
  std::memcpy(v.w, p, 16);
  return v;
}
void st(uint8_t* p, const ScaleVec& v) { std::memcpy(p, v.w, 16); }

template <int MODE> void replay_quad(uint8_t* img, const ScaleJobDev& j, uint32_t quad) {
  uint8_t rnd[SCALE_ROWS * SCALE_RND_STRIDE];
  const uint32_t y0 = (quad - j.quad_begin) * SCALE_ROWS;
  uint8_t* const base = img + j.offset + (uint64_t)(j.off_y + y0) * j.pitch + (uint64_t)j.group0 * 16;
  const uint32_t iters = (j.ngroups + 31) / 32;
  if (MODE == 0) {
    int32_t state[32];
    for (int lane = 0; lane < 32; ++lane)
      state[lane] = j.dither ? scale_sse2_seed(j.crop_w, y0 + (lane >> 3), lane & 7) : 0;
    for (uint32_t it = 0; it < iters; ++it) {
      if (j.dither)
        for (int lane = 0; lane < 32; ++lane)
          scale_sse2_advance(state[lane], lane, rnd);
      // __syncwarp()
      for (int lane = 0; lane < 32; ++lane) {
        const uint32_t g = it * 32 + lane;
        if (g < j.ngroups) {
          ScaleVec v[SCALE_ROWS];
          for (int r = 0; r < SCALE_ROWS; ++r)
            if (y0 + r < j.crop_h)
              v[r] = ld(base + (uint64_t)r * j.pitch + (uint64_t)g * 16);
          for (int r = 0; r < SCALE_ROWS; ++r) {
            if (y0 + r < j.crop_h) {
              uint32_t rb[2] = {0u, 0u};
              if (j.dither)
                std::memcpy(rb, rnd + r * SCALE_RND_STRIDE + lane * 8, 8);
              const ScaleVec o = scale_sse2_group(v[r], j, (j.off_y + y0 + r) & 1u, rb[0], rb[1]);
              st(base + (uint64_t)r * j.pitch + (uint64_t)g * 16, o);
            }
          }
        }
      }
      // __syncwarp()
    }
  } else {
    const uint32_t jump = scale_powmod(248u);
    for (int lane = 0; lane < 32; ++lane) {
      const int32_t x_first = (int32_t)(8u * (uint32_t)lane) - (int32_t)j.skip;
      uint32_t stt[SCALE_ROWS];
      for (int r = 0; r < SCALE_ROWS; ++r)
        stt[r] = j.dither ? scale_mwc_state(j.crop_w, y0 + r, (uint32_t)(x_first > 0 ? x_first : 0)) : 0u;
      for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t g = it * 32 + lane;
        if (g < j.ngroups) {
          ScaleVec v[SCALE_ROWS];
          for (int r = 0; r < SCALE_ROWS; ++r)
            if (y0 + r < j.crop_h)
              v[r] = ld(base + (uint64_t)r * j.pitch + (uint64_t)g * 16);
          const int32_t x0 = (int32_t)(8u * g) - (int32_t)j.skip;
          for (int r = 0; r < SCALE_ROWS; ++r) {
            if (y0 + r < j.crop_h) {
              const ScaleVec o = scale_plain_group(v[r], j, y0 + r, x0, stt[r]);
              st(base + (uint64_t)r * j.pitch + (uint64_t)g * 16, o);
              if (j.dither)
                stt[r] = scale_mwc_jump(stt[r], 248u, jump);
            }
          }
        }
      }
    }
  }
}
} // namespace

// Scales `img` as one launch of the scale plan would: returns -1 and a message for a refused
// descriptor, else the number of "launches" (1 or 2: one per loop kind present).
extern "C" int scale_emu_run(uint8_t* img, const rsb200_scale_job* jobs, int njobs, char* err,
                             int errlen) {
  std::vector<ScaleJobDev> dev[2];
  uint32_t quads[2] = {0, 0};
  for (int i = 0; i < njobs; ++i) {
    ScaleJobDev d;
    int mode = 0;
    // quad_begin depends on the mode, which the builder decides: build twice
    const char* why = scale_build_job(jobs[i], 0, &d, &mode);
    if (why) {
      std::strncpy(err, why, (size_t)errlen - 1);
      err[errlen - 1] = 0;
      return -1;
    }
    d.quad_begin = quads[mode];
    quads[mode] += scale_job_quads(jobs[i]);
    dev[mode].push_back(d);
  }
  int launches = 0;
  for (int mode = 0; mode < 2; ++mode) {
    if (dev[mode].empty())
      continue;
    ++launches;
    for (uint32_t quad = 0; quad < quads[mode]; ++quad) {
      // scale_find_job
      int lo = 0, hi = (int)dev[mode].size() - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (dev[mode][mid].quad_begin <= quad)
          lo = mid;
        else
          hi = mid - 1;
      }
      if (mode == 0)
        replay_quad<0>(img, dev[mode][lo], quad);
      else
        replay_quad<1>(img, dev[mode][lo], quad);
    }
  }
  return launches;
}

// the multiply-with-carry jump ahead against plain stepping (n steps from seed v0)
extern "C" uint32_t scale_emu_mwc_direct(uint32_t crop_w, uint32_t y, uint32_t x) {
  uint32_t v = crop_w + y * 36969u;
  while (x--)
    v = scale_mwc_step(v);
  return v;
}
extern "C" uint32_t scale_emu_mwc_state(uint32_t crop_w, uint32_t y, uint32_t x) {
  return scale_mwc_state(crop_w, y, x);
}
