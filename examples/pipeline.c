/* examples/pipeline.c -- a C consumer of the C ABI (include/rawspeed_b200.h), start to finish:
 * a 12-bit packed strip is unpacked (K1), linearised through a curve (K12), black/white scaled
 * (K9) and has two bad pixels interpolated (K11) -- what a rawspeed decoder does between
 * UncompressedDecompressor::readUncompressedRaw() and handing mRaw to its caller -- with host
 * buffers (rsb200_plan_run_host_image: the library owns the device staging).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/pipeline.c -Lrawspeed_b200 -l:librawspeed_b200.so \
 *       -Wl,-rpath,$PWD/rawspeed_b200 -o /tmp/pipeline && /tmp/pipeline
 *
 * A consumer that keeps frames in HBM calls rsb200_plan_run() with its own device pointers
 * instead (the in-place plans take d_in = NULL), see bench.py / tools/quick_time.py.
 * tests/test_examples.py compiles and links this file without a GPU and runs it on one. */
#include "rawspeed_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(call)                                                                   \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != RSB200_OK) {                                                           \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ctx ? rsb200_last_error(ctx) : ""); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(void) {
  enum { W = 4000, H = 300, BPS = 12 };
  const uint32_t in_pitch = W * BPS / 8;
  const uint32_t pitch = (W * 2 + 15) / 16 * 16; /* RawImageData::createData */
  rsb200_ctx* ctx = NULL;
  rsb200_plan *unpack = NULL, *lookup = NULL, *scale = NULL, *badpix = NULL;
  uint8_t* packed = (uint8_t*)malloc((size_t)in_pitch * H);
  uint8_t* image = (uint8_t*)calloc((size_t)pitch * H, 1);
  uint16_t* table = (uint16_t*)malloc(65536 * sizeof(uint16_t));
  uint32_t s = 1, i;
  if (!packed || !image || !table)
    return 1;
  for (i = 0; i < in_pitch * H; i++) { /* the survey's LCG */
    s = s * 1664525u + 1013904223u;
    packed[i] = (uint8_t)(s >> 24);
  }
  for (i = 0; i < 65536; i++) /* a square-root-like curve over the 12-bit range, flat above */
    table[i] = (uint16_t)(i < 4096 ? (i * 15u) : 61425u);

  CHECK(rsb200_create(0, &ctx));

  { /* K1: UncompressedDecompressor::readUncompressedRaw, MSB order */
    rsb200_unpack_job j;
    memset(&j, 0, sizeof j);
    j.in_offset = 0;
    j.in_size = (uint64_t)in_pitch * H;
    j.out_offset = 0;
    j.out_pitch = (int32_t)pitch;
    j.row0 = 0;
    j.rows = H;
    j.samples = W;
    j.out_col0 = 0;
    j.in_pitch = (int32_t)in_pitch;
    j.bps = BPS;
    j.order = RSB200_MSB;
    CHECK(rsb200_unpack_plan_create(ctx, &j, 1, &unpack));
    CHECK(rsb200_plan_run_host_image(unpack, packed, (size_t)in_pitch * H, image, pitch, W * 2, H, 0));
  }
  { /* K12: mRaw->setTable(curve, false); mRaw->sixteenBitLookup() */
    rsb200_lookup_job j;
    memset(&j, 0, sizeof j);
    j.pitch = pitch;
    j.width = W;
    j.height = H;
    j.cpp = 1;
    j.table = 0;
    CHECK(rsb200_lookup_plan_create(ctx, &j, 1, table, 1, /*dither=*/0, &lookup));
    CHECK(rsb200_plan_run_host_image(lookup, NULL, 0, image, pitch, W * 2, H, 1));
  }
  { /* K9: mRaw->scaleBlackWhite() once black / white are known */
    rsb200_scale_job j;
    memset(&j, 0, sizeof j);
    j.pitch = pitch;
    j.width = W;
    j.height = H;
    j.cpp = 1;
    j.crop_x = 8;
    j.crop_y = 2;
    j.crop_w = W - 16;
    j.crop_h = H - 4;
    j.black_separate[0] = j.black_separate[1] = j.black_separate[2] = j.black_separate[3] = 960;
    j.white_point = 61425;
    j.dither = 1;
    j.path = RSB200_SCALE_AUTO;
    CHECK(rsb200_scale_plan_create(ctx, &j, 1, &scale));
    CHECK(rsb200_plan_run_host_image(scale, NULL, 0, image, pitch, W * 2, H, 1));
  }
  { /* K11: mRaw->mBadPixelPositions = {...}; mRaw->fixBadPixels() */
    const uint32_t bad[2] = {(10u << 16) | 100u, (200u << 16) | 3999u};
    rsb200_badpix_job j;
    memset(&j, 0, sizeof j);
    j.pitch = pitch;
    j.width = W;
    j.height = H;
    j.is_cfa = 1;
    j.first_position = 0;
    j.num_positions = 2;
    j.prior_map = NULL;
    CHECK(rsb200_badpix_plan_create(ctx, &j, 1, bad, 2, &badpix));
    CHECK(rsb200_plan_run_host_image(badpix, NULL, 0, image, pitch, W * 2, H, 1));
  }
  {
    const uint16_t* px = (const uint16_t*)image;
    uint64_t sum = 0;
    uint32_t r, c;
    for (r = 0; r < H; r++)
      for (c = 0; c < W; c++)
        sum += px[(size_t)r * (pitch / 2) + c];
    printf("pipeline ok: %d x %d, pixel sum %llu, first pixels %u %u %u %u, kernels launched %llu\n", W, H,
           (unsigned long long)sum, px[0], px[1], px[2], px[3],
           (unsigned long long)rsb200_kernel_launches(ctx));
  }
  rsb200_plan_destroy(unpack);
  rsb200_plan_destroy(lookup);
  rsb200_plan_destroy(scale);
  rsb200_plan_destroy(badpix);
  rsb200_destroy(ctx);
  free(packed);
  free(image);
  free(table);
  return 0;
}
