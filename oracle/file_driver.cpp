// file_driver.cpp -- the reference's consumer API over a whole file (test infrastructure):
// RawParser(Buffer).getDecoder() -> checkSupport(empty CameraMetaData) -> decodeRaw(), as
// src/utilities/rsbench/main.cpp:129-186 does.  Linked once against the unmodified reference
// (libref_full.so) and once against the reference with the four hot-path bodies replaced by the
// rawspeed_b200 C ABI (libdropin.so); see oracle/Makefile.dropin.
#include "RawSpeed-API.h"
#include "decoders/RawDecoderException.h"
#include "io/IOException.h"

#include <cstdio>
#include <cstring>
#include <memory>

using namespace rawspeed;

static int g_threads = 1;
extern "C" int rawspeed_get_number_of_processor_cores() { return g_threads; }

extern "C" int rs_file_decode(const uint8_t* file, uint32_t size, uint8_t* out, uint64_t out_cap,
                              int32_t info[8], char* err, int errlen, int threads,
                              int fail_on_unknown, int uncorrected) {
  g_threads = threads > 0 ? threads : 1;
  try {
    const Buffer buf(file, size);
    RawParser parser(buf);
    std::unique_ptr<RawDecoder> d = parser.getDecoder();
    if (!d) {
      snprintf(err, (size_t)errlen, "no decoder");
      return 3;
    }
    CameraMetaData meta; // empty database, as rsbench without cameras.xml
    d->failOnUnknown = fail_on_unknown != 0;
    d->uncorrectedRawValues = uncorrected != 0; // a cameras.xml-style hint the hot path honours
    d->checkSupport(&meta);
    d->decodeRaw();
    RawImage r = d->mRaw;
    const auto dim = r->getUncroppedDim();
    info[0] = dim.x;
    info[1] = dim.y;
    info[2] = (int)r->getCpp();
    info[3] = r->pitch;
    info[4] = (int)r->getBpp();
    info[5] = r->getDataType() == RawImageType::UINT16 ? 0 : 1;
    info[6] = (int)r->getErrors().size();
    info[7] = r->isCFA ? 1 : 0;
    const uint64_t bytes = (uint64_t)r->pitch * (uint64_t)dim.y;
    if (bytes > out_cap) {
      snprintf(err, (size_t)errlen, "output buffer too small: %llu > %llu", (unsigned long long)bytes,
               (unsigned long long)out_cap);
      return 4;
    }
    const auto a = r->getByteDataAsUncroppedArray2DRef();
    for (int y = 0; y < dim.y; ++y)
      memcpy(out + (uint64_t)y * (uint64_t)r->pitch, &a(y, 0), (size_t)a.width());
    return 0;
  } catch (const RawDecoderException& e) {
    snprintf(err, (size_t)errlen, "%s", e.what());
    return 1;
  } catch (const IOException& e) {
    snprintf(err, (size_t)errlen, "%s", e.what());
    return 2;
  } catch (const RawspeedException& e) {
    snprintf(err, (size_t)errlen, "%s", e.what());
    return 5;
  }
}
