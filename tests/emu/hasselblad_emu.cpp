// hasselblad_emu.cpp -- CPU replay of the K2H kernels (rawspeed_b200/csrc/hasselblad.cuh): the kernel
// bodies themselves, compiled by g++ against tests/emu/cuda_emu.h, in the order the plan launches
// them.  Test infrastructure (no GPU needed); parity of the real kernels is the GPU tests' job.
#include "cuda_emu.h"

#include "../../rawspeed_b200/csrc/hasselblad.cuh"
#include "../../rawspeed_b200/csrc/ljpeg_host.h"

#include <vector>

using namespace rsb200;

// rounds_out: (parse, link) rounds that changed something; max_rounds < H_ROUNDS exercises the serial walk
extern "C" int hass_emu_run(const uint8_t* in, uint32_t in_size, const rsb200_huff_table* table, int w, int h,
                            int out_pitch, int init_pred, uint8_t* out, uint32_t* status, uint32_t* consumed,
                            int max_rounds, int* rounds_out) {
  DevTable tab;
  if (!build_dev_table(*table, tab))
    return -2;
  DevHassJob j{};
  j.in_offset = 0;
  j.in_size = in_size;
  j.w = (uint32_t)w;
  j.h = (uint32_t)h;
  j.out_pitch = (uint32_t)out_pitch;
  j.out_offset = 0;
  j.init_pred = (uint32_t)init_pred;
  j.table = 0;
  j.seg_begin = 0;
  j.nseg = (uint32_t)((((uint64_t)in_size + 24) * 8 + H_SEG_BITS - 1) / H_SEG_BITS);
  j.cta_begin = 0;
  const uint32_t nseg = j.nseg, ncta = (nseg + H_NT - 1) / H_NT;
  std::vector<DevHassCta> ctas(ncta);
  for (uint32_t c = 0; c < ncta; ++c)
    ctas[c] = DevHassCta{0, c * H_NT};
  std::vector<uint32_t> seg_job(nseg, 0), start(nseg), parsed(nseg, 0xFFFFFFFFu), exitp(nseg, 0), count(nseg, 0);
  for (uint32_t g = 0; g < nseg; ++g)
    start[g] = g * H_SEG_BITS;
  std::vector<uint32_t> cta_sum(ncta), cta_base(ncta), changed((size_t)H_ROUNDS + 1, 0);
  DevHassState st{H_NOKEY, H_NOKEY, 0, 0};
  // the stream at a 16-byte aligned address, exactly in_size bytes readable
  std::vector<uint8_t> buf((size_t)in_size + 32);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(buf.data()) + 15) & ~(uintptr_t)15);
  memcpy(base, in, in_size);
  const int rounds = max_rounds < H_ROUNDS ? max_rounds : H_ROUNDS;
  int used = 0;
  for (int r = 0; r < rounds; ++r) {
    for (uint32_t c = 0; c < ncta; ++c)
      cuemu::run_cta(c, ncta, H_NT, sizeof(HassShared), false, [&](uint8_t* smem) {
        hass_parse_entry(*reinterpret_cast<HassShared*>(smem), base, &j, &tab, ctas.data(), start.data(),
                         parsed.data(), exitp.data(), count.data());
      });
    const uint32_t nb = (nseg + 255) / 256;
    for (uint32_t b = 0; b < nb; ++b)
      cuemu::run_cta(b, nb, 256, 16, false, [&](uint8_t*) {
        hass_link_entry(&j, 1, nseg, seg_job.data(), start.data(), exitp.data(), &changed[(size_t)r]);
      });
    if (changed[(size_t)r])
      used = r + 1;
  }
  uint32_t force = 1; // the serial walk checks everything again (a no-op once the chain has settled)
  cuemu::run_cta(0, 1, 32, 16, false, [&](uint8_t*) {
    hass_serial_entry(base, &j, &tab, start.data(), parsed.data(), exitp.data(), count.data(), &force);
  });
  if (rounds_out)
    *rounds_out = used;
  for (uint32_t c = 0; c < ncta; ++c)
    cuemu::run_cta(c, ncta, H_NT, sizeof(HassShared), false, [&](uint8_t* smem) {
      hass_ctasum_entry(*reinterpret_cast<HassShared*>(smem), &j, ctas.data(), count.data(), cta_sum.data());
    });
  cuemu::run_cta(0, 1, 32, 16, false,
                 [&](uint8_t*) { hass_ctascan_entry(&j, 1, cta_sum.data(), cta_base.data()); });
  for (uint32_t c = 0; c < ncta; ++c)
    cuemu::run_cta(c, ncta, H_NT, sizeof(HassShared), false, [&](uint8_t* smem) {
      hass_decode_entry(*reinterpret_cast<HassShared*>(smem), base, &j, &tab, ctas.data(), start.data(),
                        exitp.data(), count.data(), cta_base.data(), out, &st);
    });
  const uint32_t row_begin[2] = {0, (uint32_t)h};
  const uint32_t nrb = ((uint32_t)h * 32 + 255) / 256;
  for (uint32_t b = 0; b < nrb; ++b)
    cuemu::run_cta(b, nrb, 256, 16, false, [&](uint8_t*) { hass_rows_entry(&j, 1, row_begin, out); });
  *status = st.key_ioe != H_NOKEY && st.key_ioe <= st.key_bad ? 2u : (st.key_bad != H_NOKEY ? 1u : 0u);
  *consumed = st.consumed;
  return 0;
}
