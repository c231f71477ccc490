/*
 * ref_driver.cpp -- thin extern "C" shim over the UNMODIFIED reference
 * (darktable-org/rawspeed), compiled from the sources where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libref.so.
 *
 * TEST INFRASTRUCTURE ONLY.  Used (a) to pin oracle/rs_oracle.c against the
 * real reference, (b) as the "reference" CPU baseline in bench.py.  Nothing in
 * rawspeed_b200/ may load it.  No reference source is copied: this file only
 * *calls* the reference's public classes:
 *   UncompressedDecompressor   decompressors/UncompressedDecompressor.h:64-75
 *   LJpegDecompressor          decompressors/LJpegDecompressor.h:44-94
 *   LJpegDecoder               decompressors/LJpegDecoder.h:31-48
 *   AbstractDngDecompressor    decompressors/AbstractDngDecompressor.h:135-150
 *   Cr2Decompressor            decompressors/Cr2Decompressor.h:125-174
 *   Cr2LJpegDecoder            decompressors/Cr2LJpegDecoder.h:30-40
 *   BitStreamer{LSB,MSB,MSB16,MSB32,JPEG}, PrefixCodeDecoder<>, HuffmanCode<>
 */
#include "rawspeedconfig.h"

#include "adt/Array1DRef.h"
#include "adt/PartitioningOutputIterator.h"
#include "adt/Point.h"
#include "bitstreams/BitStreamerJPEG.h"
#include "bitstreams/BitStreamerLSB.h"
#include "bitstreams/BitStreamerMSB.h"
#include "bitstreams/BitStreamerMSB16.h"
#include "bitstreams/BitStreamerMSB32.h"
#include "bitstreams/BitStreams.h"
#include "bitstreams/BitVacuumerJPEG.h"
#include "codes/HuffmanCode.h"
#include "codes/PrefixCodeDecoder.h"
#include "codes/PrefixCodeVectorEncoder.h"
#include "common/DngOpcodes.h"
#include "common/RawImage.h"
#include "common/RawspeedException.h"
#include "decoders/RawDecoderException.h"
#include "decompressors/AbstractDngDecompressor.h"
#include "decompressors/Cr2Decompressor.h"
#include "decompressors/Cr2LJpegDecoder.h"
#include "decompressors/LJpegDecoder.h"
#include "decompressors/LJpegDecompressor.h"
#include "decompressors/UncompressedDecompressor.h"
#include "interpolators/Cr2sRawInterpolator.h"
#include "decompressors/PentaxDecompressor.h"
#include "decompressors/SonyArw2Decompressor.h"
#include "decompressors/NikonDecompressor.h"
#include "decompressors/HasselbladDecompressor.h"
#include "decompressors/HasselbladLJpegDecoder.h"
#include "decompressors/PhaseOneDecompressor.h"
#include "decompressors/PanasonicV4Decompressor.h"
#include "decompressors/PanasonicV5Decompressor.h"
#include "decompressors/PanasonicV6Decompressor.h"
#include "decompressors/PanasonicV7Decompressor.h"
#include "io/Buffer.h"
#include "io/ByteStream.h"
#include "io/Endianness.h"
#include "io/IOException.h"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

using namespace rawspeed;

// The consumer supplies the thread count (common/Common.h:41).
static int g_threads = 1;
extern "C" int rawspeed_get_number_of_processor_cores() { return g_threads; }

namespace {

struct RefErr {
  int code;
  char msg[240];
};

template <typename F> int guarded(RefErr* e, F&& f) {
  if (e) {
    e->code = 0;
    e->msg[0] = 0;
  }
  try {
    f();
    return 0;
  } catch (const IOException& ex) {
    if (e) {
      e->code = 2;
      std::snprintf(e->msg, sizeof e->msg, "%s", ex.what());
    }
    return 2;
  } catch (const RawDecoderException& ex) {
    if (e) {
      e->code = 1;
      std::snprintf(e->msg, sizeof e->msg, "%s", ex.what());
    }
    return 1;
  } catch (const RawspeedException& ex) {
    if (e) {
      e->code = 3;
      std::snprintf(e->msg, sizeof e->msg, "%s", ex.what());
    }
    return 3;
  }
}

// wall time of the reference call inside the last post-decode driver (ref_scale_values,
// ref_scale_black_white, ref_sixteen_bit_lookup, ref_fix_bad_pixels, ref_dng_opcodes), copies
// in and out of the driver's image excluded
double g_last_ms = 0.0;
struct StageTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~StageTimer() {
    g_last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
};

RawImage makeImage(int w, int h, int cpp, bool isCfa, int subX, int subY) {
  RawImage img = RawImage::create(iPoint2D(w, h), RawImageType::UINT16, cpp);
  img->isCFA = isCfa;
  img->metadata.subsampling = iPoint2D(subX, subY);
  return img;
}

void copyOut(const RawImage& img, uint16_t* out, int outPitchBytes) {
  const auto a = img->getU16DataAsUncroppedArray2DRef();
  for (int r = 0; r < a.height(); ++r)
    std::memcpy(reinterpret_cast<uint8_t*>(out) +
                    static_cast<size_t>(r) * outPitchBytes,
                &a(r, 0), sizeof(uint16_t) * a.width());
}
void copyIn(const RawImage& img, const uint16_t* in, int inPitchBytes) {
  const auto a = img->getU16DataAsUncroppedArray2DRef();
  for (int r = 0; r < a.height(); ++r)
    std::memcpy(&a(r, 0),
                reinterpret_cast<const uint8_t*>(in) +
                    static_cast<size_t>(r) * inPitchBytes,
                sizeof(uint16_t) * a.width());
}

PrefixCodeDecoder<> makeHT(const uint8_t* ncpl, const uint8_t* values,
                           int nvalues, bool full, bool fix16) {
  HuffmanCode<BaselineCodeTag> hc;
  hc.setNCodesPerLength(Buffer(ncpl, 16));
  hc.setCodeValues(Array1DRef<const uint8_t>(values, nvalues));
  PrefixCodeDecoder<> ht(std::move(hc));
  ht.setup(full, fix16);
  return ht;
}

template <typename Pump>
void pumpGet(const uint8_t* data, int size, const int* lens, int n,
             uint32_t* out, int* pos) {
  Pump p(Array1DRef<const std::byte>(reinterpret_cast<const std::byte*>(data),
                                     size));
  for (int i = 0; i < n; ++i)
    out[i] = p.getBits(lens[i]);
  if (pos)
    *pos = p.getStreamPosition();
}

} // namespace

extern "C" {

double ref_last_ms(void) { return g_last_ms; }


int ref_image_pitch(int w, int h, int cpp) {
  RawImage img = RawImage::create(iPoint2D(w, h), RawImageType::UINT16, cpp);
  return img->pitch;
}

void ref_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

int ref_pump_getbits(int order, const uint8_t* data, int size, const int* lens,
                     int n, uint32_t* out, int* pos, RefErr* e) {
  return guarded(e, [&] {
    switch (static_cast<BitOrder>(order)) {
    case BitOrder::LSB:
      pumpGet<BitStreamerLSB>(data, size, lens, n, out, pos);
      break;
    case BitOrder::MSB:
      pumpGet<BitStreamerMSB>(data, size, lens, n, out, pos);
      break;
    case BitOrder::MSB16:
      pumpGet<BitStreamerMSB16>(data, size, lens, n, out, pos);
      break;
    case BitOrder::MSB32:
      pumpGet<BitStreamerMSB32>(data, size, lens, n, out, pos);
      break;
    case BitOrder::JPEG:
      pumpGet<BitStreamerJPEG>(data, size, lens, n, out, pos);
      break;
    }
  });
}

int ref_huff_check(const uint8_t* ncpl, const uint8_t* values, int nvalues,
                   int full, int fix16, RefErr* e) {
  return guarded(e, [&] { (void)makeHT(ncpl, values, nvalues, full, fix16); });
}

int ref_huff_decode(const uint8_t* ncpl, const uint8_t* values, int nvalues,
                    int full, int fix16, int order, const uint8_t* data,
                    int size, int n, int32_t* out, RefErr* e) {
  return guarded(e, [&] {
    auto ht = makeHT(ncpl, values, nvalues, full, fix16);
    const Array1DRef<const std::byte> in(
        reinterpret_cast<const std::byte*>(data), size);
    if (order == static_cast<int>(BitOrder::JPEG)) {
      BitStreamerJPEG bs(in);
      for (int i = 0; i < n; ++i)
        out[i] = full ? ht.decodeDifference(bs) : ht.decodeCodeValue(bs);
    } else {
      BitStreamerMSB bs(in);
      for (int i = 0; i < n; ++i)
        out[i] = full ? ht.decodeDifference(bs) : ht.decodeCodeValue(bs);
    }
  });
}

// Encode differences with the reference's own writer half.
int64_t ref_encode_diffs(const int32_t* diffs, uint64_t n, const uint8_t* ncpl,
                         const uint8_t* values, int nvalues, int fix16,
                         uint8_t* out, uint64_t cap) {
  HuffmanCode<BaselineCodeTag> hc;
  hc.setNCodesPerLength(Buffer(ncpl, 16));
  hc.setCodeValues(Array1DRef<const uint8_t>(values, nvalues));
  PrefixCodeVectorEncoder<BaselineCodeTag> enc(
      static_cast<PrefixCode<BaselineCodeTag>>(std::move(hc)));
  enc.setup(true, fix16);
  std::vector<uint8_t> buf;
  buf.reserve(n * 2 + 16);
  {
    auto bsInserter = PartitioningOutputIterator(std::back_inserter(buf));
    using BitVacuumer = BitVacuumerJPEG<decltype(bsInserter)>;
    auto bv = BitVacuumer(bsInserter);
    for (uint64_t i = 0; i < n; ++i)
      enc.encodeDifference(bv, diffs[i]);
  }
  if (buf.size() > cap)
    return -1;
  std::memcpy(out, buf.data(), buf.size());
  return static_cast<int64_t>(buf.size());
}

int ref_unpack(const uint8_t* in, uint32_t in_size, uint16_t* img_data, int w,
               int h, int cpp, int pitch, int crop_x, int crop_y, int crop_w,
               int crop_h, int in_pitch, int bps, int order, int reps,
               double* best_ms, RefErr* e) {
  return guarded(e, [&] {
    RawImage img = makeImage(w, h, cpp, true, 1, 1);
    copyIn(img, img_data, pitch);
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      UncompressedDecompressor u(
          ByteStream(DataBuffer(Buffer(in, in_size), Endianness::little)), img,
          iRectangle2D({crop_x, crop_y}, {crop_w, crop_h}), in_pitch, bps,
          static_cast<BitOrder>(order));
      u.readUncompressedRaw();
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(
          best, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    if (best_ms)
      *best_ms = best;
    copyOut(img, img_data, pitch);
  });
}

// UncompressedDecompressor: the members other than readUncompressedRaw() on a
// uint16 image.  form: 0 readUncompressedRaw (is_f32: F32 image), 1/2
// decode8BitRaw<false/true>, 3/4 decode12BitRawWithControl<big/little>, 5/6
// decode12BitRawUnpackedLeftAligned<big/little>.  curve != nullptr ->
// mRaw->setTable(curve, dither) first (RawImage.cpp setTable / TableLookUp.cpp).
int ref_unpack_form(const uint8_t* in, uint32_t in_size, void* img_data, int is_f32,
                    int w, int h, int cpp, int pitch, int crop_x, int crop_y, int crop_w,
                    int crop_h, int in_pitch, int bps, int order, int form,
                    const uint16_t* curve, int ncurve, int dither, int reps,
                    double* best_ms, RefErr* e) {
  return guarded(e, [&] {
    RawImage img = RawImage::create(iPoint2D(w, h),
                                    is_f32 ? RawImageType::F32 : RawImageType::UINT16, cpp);
    const int bpp = (is_f32 ? 4 : 2) * cpp;
    auto rowPtr = [&](int r) {
      if (is_f32)
        return reinterpret_cast<uint8_t*>(&img->getF32DataAsUncroppedArray2DRef()(r, 0));
      return reinterpret_cast<uint8_t*>(&img->getU16DataAsUncroppedArray2DRef()(r, 0));
    };
    for (int r = 0; r < h; ++r)
      std::memcpy(rowPtr(r), static_cast<const uint8_t*>(img_data) + static_cast<size_t>(r) * pitch,
                  static_cast<size_t>(w) * bpp);
    if (curve)
      img->setTable(std::vector<uint16_t>(curve, curve + ncurve), dither != 0);
    double best = 1e30;
    for (int rep = 0; rep < (reps < 1 ? 1 : reps); ++rep) {
      const auto t0 = std::chrono::steady_clock::now();
      UncompressedDecompressor u(
          ByteStream(DataBuffer(Buffer(in, in_size), Endianness::little)), img,
          iRectangle2D({crop_x, crop_y}, {crop_w, crop_h}), in_pitch, bps,
          static_cast<BitOrder>(order));
      switch (form) {
      case 0: u.readUncompressedRaw(); break;
      case 1: u.decode8BitRaw<false>(); break;
      case 2: u.decode8BitRaw<true>(); break;
      case 3: u.decode12BitRawWithControl<Endianness::big>(); break;
      case 4: u.decode12BitRawWithControl<Endianness::little>(); break;
      case 5: u.decode12BitRawUnpackedLeftAligned<Endianness::big>(); break;
      case 6: u.decode12BitRawUnpackedLeftAligned<Endianness::little>(); break;
      default: ThrowRDE("unknown form");
      }
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(best, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    if (best_ms)
      *best_ms = best;
    for (int r = 0; r < h; ++r)
      std::memcpy(static_cast<uint8_t*>(img_data) + static_cast<size_t>(r) * pitch, rowPtr(r),
                  static_cast<size_t>(w) * bpp);
  });
}

// HasselbladDecompressor(mRaw, {ht, initPred}, input).decompress(); the table is set up
// as a code-value table (full = false), as HasselbladLJpegDecoder::decode arranges.
int ref_hasselblad_decompress(uint16_t* img_data, int w, int h, int pitch, const uint8_t* ncpl,
                              const uint8_t* values, int nvalues, int full, int init_pred,
                              const uint8_t* data, uint32_t size, uint32_t* consumed, int reps,
                              double* best_ms, RefErr* e) {
  return guarded(e, [&] {
    RawImage img = makeImage(w, h, 1, true, 1, 1);
    copyIn(img, img_data, pitch);
    const PrefixCodeDecoder<> ht = makeHT(ncpl, values, nvalues, full, 0);
    const HasselbladDecompressor::PerComponentRecipe rec = {ht, static_cast<uint16_t>(init_pred)};
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      HasselbladDecompressor d(img, rec, Array1DRef<const uint8_t>(data, static_cast<int>(size)));
      const auto c = d.decompress();
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(best, std::chrono::duration<double, std::milli>(t1 - t0).count());
      if (consumed)
        *consumed = c;
    }
    if (best_ms)
      *best_ms = best;
    copyOut(img, img_data, pitch);
  });
}

// HasselbladLJpegDecoder(bs, img).decode(): the LJPEG container (SOI, DHT, SOF3, SOS, pair stream,
// EOI) around HasselbladDecompressor.
int ref_hasselblad_ljpeg_decode(const uint8_t* in, uint32_t in_size, uint16_t* img_data, int w, int h,
                                int pitch, RefErr* e) {
  return guarded(e, [&] {
    RawImage img = makeImage(w, h, 1, true, 1, 1);
    copyIn(img, img_data, pitch);
    HasselbladLJpegDecoder d(ByteStream(DataBuffer(Buffer(in, in_size), Endianness::big)), img);
    d.decode();
    copyOut(img, img_data, pitch);
  });
}

// RawImageDataU16::scaleBlackWhite() with blackLevelSeparate and whitePoint given (no
// estimation, no black areas): the SCALE_VALUES worker over the cropped rows.
int ref_scale_values(uint16_t* img_data, int w, int h, int pitch, int off_x, int off_y,
                     int crop_w, int crop_h, const int* black_sep, int white, int dither,
                     int nthreads, RefErr* e) {
  return guarded(e, [&] {
    ref_set_threads(nthreads);
    RawImage img = makeImage(w, h, 1, true, 1, 1);
    copyIn(img, img_data, pitch);
    img->subFrame(iRectangle2D(iPoint2D(off_x, off_y), iPoint2D(crop_w, crop_h)));
    img->blackLevelSeparate = Array2DRef<int>(img->blackLevelSeparateStorage.data(), 2, 2);
    for (int i = 0; i < 4; ++i)
      img->blackLevelSeparateStorage[i] = black_sep[i];
    img->whitePoint = white;
    img->mDitherScale = dither != 0;
    {
      StageTimer tm;
      img->scaleBlackWhite();
    }
    // copy out the whole uncropped buffer
    const auto a = img->getU16DataAsUncroppedArray2DRef();
    for (int r = 0; r < a.height(); ++r)
      std::memcpy(reinterpret_cast<uint8_t*>(img_data) + static_cast<size_t>(r) * pitch, &a(r, 0),
                  sizeof(uint16_t) * a.width());
  });
}

// RawImageDataU16::scaleBlackWhite() in full: blackLevel, optional blackLevelSeparate /
// whitePoint, blackAreas (triples is_vertical, offset, size); reports what it settled on.
int ref_scale_black_white(uint16_t* img_data, int w, int h, int cpp, int pitch, int is_cfa, int off_x,
                          int off_y, int crop_w, int crop_h, int black_level, int* black_sep,
                          int has_sep, int* white, int has_white, const int* areas, int n_areas,
                          int dither, int nthreads, int* sep_set, RefErr* e) {
  return guarded(e, [&] {
    ref_set_threads(nthreads);
    RawImage img = makeImage(w, h, cpp, is_cfa != 0, 1, 1);
    copyIn(img, img_data, pitch);
    img->subFrame(iRectangle2D(iPoint2D(off_x, off_y), iPoint2D(crop_w, crop_h)));
    img->blackLevel = black_level;
    if (has_sep) {
      img->blackLevelSeparate = Array2DRef<int>(img->blackLevelSeparateStorage.data(), 2, 2);
      for (int i = 0; i < 4; ++i)
        img->blackLevelSeparateStorage[i] = black_sep[i];
    }
    if (has_white)
      img->whitePoint = *white;
    for (int i = 0; i < n_areas; ++i)
      img->blackAreas.emplace_back(areas[3 * i + 1], areas[3 * i + 2], areas[3 * i] != 0);
    img->mDitherScale = dither != 0;
    {
      StageTimer tm;
      img->scaleBlackWhite();
    }
    copyOut(img, img_data, pitch);
    *sep_set = img->blackLevelSeparate.has_value();
    if (img->blackLevelSeparate)
      for (int i = 0; i < 4; ++i)
        black_sep[i] = img->blackLevelSeparateStorage[i];
    *white = img->whitePoint.has_value() ? *img->whitePoint : -1;
  });
}

// DngOpcodes(ri, bs) + applyOpCodes(ri) on a uint16 (is_f32 == 0) or float image with the crop
// crop[4] = (mOffset.x, mOffset.y, dim.x, dim.y); reports the crop and mBadPixelPositions
// afterwards.  stage: which half threw (1 constructor, 2 applyOpCodes), 0 if none.
int ref_dng_opcodes(void* img_data, int is_f32, int w, int h, int cpp, int pitch, int* crop,
                    const uint8_t* data, uint32_t size, uint32_t* bad, uint32_t bad_cap,
                    uint32_t* nbad, int* stage, RefErr* e) {
  *stage = 0;
  return guarded(e, [&] {
    RawImage img = RawImage::create(iPoint2D(w, h),
                                    is_f32 ? RawImageType::F32 : RawImageType::UINT16, cpp);
    if (img->pitch != pitch)
      ThrowRDE("driver: pitch mismatch (%d vs %d)", img->pitch, pitch);
    uint8_t* base = is_f32 ? reinterpret_cast<uint8_t*>(&img->getF32DataAsUncroppedArray2DRef()(0, 0))
                           : reinterpret_cast<uint8_t*>(&img->getU16DataAsUncroppedArray2DRef()(0, 0));
    std::memcpy(base, img_data, static_cast<size_t>(pitch) * h);
    if (crop[0] || crop[1] || crop[2] != w || crop[3] != h)
      img->subFrame(iRectangle2D(iPoint2D(crop[0], crop[1]), iPoint2D(crop[2], crop[3])));
    auto copyBack = [&] {
      std::memcpy(img_data, base, static_cast<size_t>(pitch) * h);
      const iPoint2D o = img->getCropOffset();
      crop[0] = o.x;
      crop[1] = o.y;
      crop[2] = img->dim.x;
      crop[3] = img->dim.y;
      *nbad = static_cast<uint32_t>(img->mBadPixelPositions.size());
      for (uint32_t i = 0; i < *nbad && i < bad_cap; ++i)
        bad[i] = img->mBadPixelPositions[i];
    };
    *stage = 1;
    DngOpcodes codes(img, ByteStream(DataBuffer(Buffer(data, size), Endianness::little)));
    *stage = 2;
    try {
      StageTimer tm;
      codes.applyOpCodes(img);
    } catch (...) {
      copyBack();
      throw;
    }
    *stage = 0;
    copyBack();
  });
}

// mRaw->setTable(curve, dither); mRaw->sixteenBitLookup() on an image cropped to crop[4]
int ref_sixteen_bit_lookup(uint16_t* img_data, int w, int h, int cpp, int pitch, const int* crop,
                           const uint16_t* curve, int ncurve, int dither, int nthreads, RefErr* e) {
  return guarded(e, [&] {
    ref_set_threads(nthreads);
    RawImage img = makeImage(w, h, cpp, true, 1, 1);
    copyIn(img, img_data, pitch);
    if (crop[0] || crop[1] || crop[2] != w || crop[3] != h)
      img->subFrame(iRectangle2D(iPoint2D(crop[0], crop[1]), iPoint2D(crop[2], crop[3])));
    if (curve)
      img->setTable(std::vector<uint16_t>(curve, curve + ncurve), dither != 0);
    {
      StageTimer tm;
      img->sixteenBitLookup();
    }
    copyOut(img, img_data, pitch);
  });
}

// RawImageData::fixBadPixels() with mBadPixelPositions = positions[0..n)
int ref_fix_bad_pixels(uint16_t* img_data, int w, int h, int cpp, int pitch, int is_cfa,
                       const uint32_t* positions, uint32_t n, int nthreads, RefErr* e) {
  return guarded(e, [&] {
    ref_set_threads(nthreads);
    RawImage img = makeImage(w, h, cpp, is_cfa != 0, 1, 1);
    copyIn(img, img_data, pitch);
    img->mBadPixelPositions.assign(positions, positions + n);
    {
      StageTimer tm;
      img->fixBadPixels();
    }
    copyOut(img, img_data, pitch);
  });
}

// PhaseOneDecompressor(mRaw, strips).decompress(): strip k = (row rown[k], bytes
// [off[k], off[k]+len[k]) of `file`), as IiqDecoder::DecodePhaseOneC builds them.
int ref_phaseone(uint16_t* img_data, int w, int h, int pitch, const uint8_t* file,
                 uint64_t file_size, const uint64_t* off, const uint32_t* len, const int32_t* rown,
                 int nstrips, int nthreads, int reps, double* best_ms, RefErr* e) {
  return guarded(e, [&] {
    ref_set_threads(nthreads);
    RawImage img = makeImage(w, h, 1, true, 1, 1);
    copyIn(img, img_data, pitch);
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
      std::vector<PhaseOneStrip> strips;
      strips.reserve(static_cast<size_t>(nstrips));
      for (int k = 0; k < nstrips; ++k) {
        if (off[k] + len[k] > file_size)
          ThrowIOE("Out of bounds access in ByteStream");
        strips.emplace_back(rown[k], ByteStream(DataBuffer(Buffer(file + off[k], len[k]),
                                                          Endianness::little)));
      }
      const auto t0 = std::chrono::steady_clock::now();
      PhaseOneDecompressor d(img, std::move(strips));
      d.decompress();
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(best, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    if (best_ms)
      *best_ms = best;
    copyOut(img, img_data, pitch);
  });
}

// PanasonicV4Decompressor(mRaw, input, zero_is_not_bad, section_split_offset).decompress();
// zero_pos receives mRaw->mBadPixelPositions (sorted: the threads append in any order).
int ref_panasonic_v4(uint16_t* img_data, int w, int h, int pitch, const uint8_t* data,
                     uint32_t size, int zero_is_not_bad, uint32_t section_split_offset,
                     uint32_t* zero_pos, uint32_t cap, uint32_t* nzero, int nthreads,
                     RefErr* e) {
  return guarded(e, [&] {
    ref_set_threads(nthreads);
    RawImage img = makeImage(w, h, 1, true, 1, 1);
    copyIn(img, img_data, pitch);
    PanasonicV4Decompressor d(img, ByteStream(DataBuffer(Buffer(data, size), Endianness::little)),
                              zero_is_not_bad != 0, section_split_offset);
    d.decompress();
    std::vector<uint32_t> z(img->mBadPixelPositions.begin(), img->mBadPixelPositions.end());
    std::sort(z.begin(), z.end());
    if (nzero)
      *nzero = static_cast<uint32_t>(z.size());
    for (size_t i = 0; i < z.size() && i < cap && zero_pos; ++i)
      zero_pos[i] = z[i];
    copyOut(img, img_data, pitch);
  });
}

// PanasonicV{5,6,7}Decompressor(mRaw, input[, bps]).decompress()
int ref_panasonic(int version, uint16_t* img_data, int w, int h, int pitch, const uint8_t* data,
                  uint32_t size, int bps, int nthreads, int reps, double* best_ms, RefErr* e) {
  return guarded(e, [&] {
    ref_set_threads(nthreads);
    RawImage img = makeImage(w, h, 1, true, 1, 1);
    copyIn(img, img_data, pitch);
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
      const ByteStream in(DataBuffer(Buffer(data, size), Endianness::little));
      const auto t0 = std::chrono::steady_clock::now();
      if (version == 5) {
        PanasonicV5Decompressor d(img, in, static_cast<uint32_t>(bps));
        d.decompress();
      } else if (version == 6) {
        PanasonicV6Decompressor d(img, in, static_cast<uint32_t>(bps));
        d.decompress();
      } else if (version == 7) {
        PanasonicV7Decompressor d(img, in);
        d.decompress();
      } else {
        ThrowRDE("unknown Panasonic version");
      }
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(best, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    if (best_ms)
      *best_ms = best;
    copyOut(img, img_data, pitch);
  });
}

// NikonDecompressor(mRaw, metadata, bitsPS).decompress(input, uncorrectedRawValues)
// (NikonDecompressor.h:51-55), as NefDecoder::DecodeNikonCompressed drives it.
int ref_nikon_decompress(uint16_t* img_data, int w, int h, int pitch, const uint8_t* meta,
                         uint32_t meta_size, int meta_be, int bitsPS, const uint8_t* data,
                         uint32_t size, int uncorrected, int reps, double* best_ms, RefErr* e) {
  return guarded(e, [&] {
    RawImage img = makeImage(w, h, 1, true, 1, 1);
    copyIn(img, img_data, pitch);
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      NikonDecompressor n(img,
                          ByteStream(DataBuffer(Buffer(meta, meta_size),
                                                meta_be ? Endianness::big : Endianness::little)),
                          static_cast<uint32_t>(bitsPS));
      n.decompress(Array1DRef<const uint8_t>(data, static_cast<int>(size)), uncorrected != 0);
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(best, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    if (best_ms)
      *best_ms = best;
    copyOut(img, img_data, pitch);
  });
}

// SonyArw2Decompressor(mRaw, input).decompress() (SonyArw2Decompressor.h); curve !=
// nullptr -> mRaw->setTable(curve, dither) first, as ArwDecoder does through
// RawImageCurveGuard (decoders/ArwDecoder.cpp).
int ref_sony_arw2(uint16_t* img_data, int w, int h, int pitch, const uint8_t* data,
                  uint32_t size, const uint16_t* curve, int ncurve, int dither, int nthreads,
                  int reps, double* best_ms, RefErr* e) {
  return guarded(e, [&] {
    ref_set_threads(nthreads);
    RawImage img = makeImage(w, h, 1, false, 1, 1);
    copyIn(img, img_data, pitch);
    if (curve)
      img->setTable(std::vector<uint16_t>(curve, curve + ncurve), dither != 0);
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      SonyArw2Decompressor a(img, ByteStream(DataBuffer(Buffer(data, size), Endianness::little)));
      a.decompress();
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(best, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    if (best_ms)
      *best_ms = best;
    copyOut(img, img_data, pitch);
  });
}

// Cr2sRawInterpolator(mRaw, input, sraw_coeffs, hue).interpolate(version)
// (interpolators/Cr2sRawInterpolator.h:36-60); out image: cpp 3, subsampling set.
int ref_sraw_interpolate(const uint16_t* in, int in_w, int in_h, int in_pitch,
                         uint16_t* out_data, int out_w, int out_h, int out_pitch, int sub_x,
                         int sub_y, const int* coeffs, int hue, int version, int nthreads,
                         int reps, double* best_ms, RefErr* e) {
  return guarded(e, [&] {
    ref_set_threads(nthreads);
    RawImage img = makeImage(out_w, out_h, 3, false, sub_x, sub_y);
    copyIn(img, out_data, out_pitch);
    const Array2DRef<const uint16_t> input(in, in_w, in_h, in_pitch / 2);
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      Cr2sRawInterpolator i(img, input, {coeffs[0], coeffs[1], coeffs[2]}, hue);
      i.interpolate(version);
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(best, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    if (best_ms)
      *best_ms = best;
    copyOut(img, out_data, out_pitch);
  });
}

// PentaxDecompressor(mRaw, metaData).decompress(data) (PentaxDecompressor.h)
int ref_pentax_decompress(uint16_t* img_data, int w, int h, int pitch, const uint8_t* meta,
                          int meta_size, int meta_be, const uint8_t* data, uint32_t size,
                          int reps, double* best_ms, RefErr* e) {
  return guarded(e, [&] {
    RawImage img = makeImage(w, h, 1, true, 1, 1);
    copyIn(img, img_data, pitch);
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      Optional<ByteStream> md;
      if (meta)
        md = ByteStream(DataBuffer(Buffer(meta, meta_size),
                                   meta_be ? Endianness::big : Endianness::little));
      PentaxDecompressor p(img, md);
      p.decompress(ByteStream(DataBuffer(Buffer(data, size), Endianness::little)));
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(best, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    if (best_ms)
      *best_ms = best;
    copyOut(img, img_data, pitch);
  });
}

struct RefHuffDesc {
  uint8_t ncpl[16];
  uint8_t values[162];
  int nvalues;
};

int ref_ljpeg_decompress(uint16_t* img_data, int w, int h, int cpp, int pitch,
                         int fx, int fy, int fw, int fh, int mcu_x, int mcu_y,
                         int dim_x, int dim_y, const RefHuffDesc* tabs,
                         const int* tab_of_comp, const uint16_t* init_pred,
                         int nrec, int fix16, int rows_per_restart,
                         const uint8_t* in, uint32_t in_size,
                         uint32_t* consumed, RefErr* e) {
  return guarded(e, [&] {
    RawImage img = makeImage(w, h, cpp, true, 1, 1);
    copyIn(img, img_data, pitch);
    std::vector<std::unique_ptr<PrefixCodeDecoder<>>> hts;
    int maxTab = 0;
    for (int i = 0; i < nrec; ++i)
      maxTab = std::max(maxTab, tab_of_comp[i]);
    for (int t = 0; t <= maxTab; ++t)
      hts.emplace_back(std::make_unique<PrefixCodeDecoder<>>(
          makeHT(tabs[t].ncpl, tabs[t].values, tabs[t].nvalues, true, fix16)));
    std::vector<LJpegDecompressor::PerComponentRecipe> rec;
    rec.reserve(nrec);
    for (int i = 0; i < nrec; ++i)
      rec.push_back({*hts[tab_of_comp[i]], init_pred[i]});
    LJpegDecompressor d(
        img, iRectangle2D({fx, fy}, {fw, fh}),
        LJpegDecompressor::Frame{iPoint2D(mcu_x, mcu_y), iPoint2D(dim_x, dim_y)},
        rec, rows_per_restart,
        Array1DRef<const uint8_t>(in, static_cast<int>(in_size)));
    const auto c = d.decode();
    if (consumed)
      *consumed = c;
    copyOut(img, img_data, pitch);
  });
}

int ref_ljpeg_decode(const uint8_t* in, uint32_t in_size, uint16_t* img_data,
                     int w, int h, int cpp, int pitch, uint32_t off_x,
                     uint32_t off_y, uint32_t tw, uint32_t th, int max_w,
                     int max_h, int fix16, RefErr* e) {
  return guarded(e, [&] {
    RawImage img = makeImage(w, h, cpp, true, 1, 1);
    copyIn(img, img_data, pitch);
    LJpegDecoder d(ByteStream(DataBuffer(Buffer(in, in_size), Endianness::little)),
                   img);
    d.decode(off_x, off_y, tw, th, iPoint2D(max_w, max_h), fix16);
    copyOut(img, img_data, pitch);
  });
}

int ref_dng_decompress(const uint8_t* file, uint64_t file_size,
                       const uint64_t* tile_off, const uint32_t* tile_len,
                       int ntiles, void* img_data, int is_f32, int w, int h, int cpp,
                       int pitch, int tile_w, int tile_h, int compression,
                       int fix_ljpeg, int bps, int big_endian, int nthreads,
                       int reps, double* best_ms, RefErr* e) {
  return guarded(e, [&] {
    ref_set_threads(nthreads);
    const Buffer whole(file, static_cast<Buffer::size_type>(file_size));
    double best = 1e30;
    RawImage img = RawImage::create(iPoint2D(w, h),
                                    is_f32 ? RawImageType::F32 : RawImageType::UINT16, cpp);
    img->isCFA = true;
    const int bpp = (is_f32 ? 4 : 2) * cpp;
    auto rowPtr = [&](int r) {
      if (is_f32)
        return reinterpret_cast<uint8_t*>(&img->getF32DataAsUncroppedArray2DRef()(r, 0));
      return reinterpret_cast<uint8_t*>(&img->getU16DataAsUncroppedArray2DRef()(r, 0));
    };
    for (int r = 0; r < h; ++r)
      std::memcpy(rowPtr(r), static_cast<const uint8_t*>(img_data) + static_cast<size_t>(r) * pitch,
                  static_cast<size_t>(w) * bpp);
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
      const iPoint2D dim(w, h);
      DngTilingDescription dsc(dim, tile_w, tile_h);
      AbstractDngDecompressor d(img, dsc, compression, fix_ljpeg, bps, 1);
      d.slices.reserve(ntiles);
      for (int n = 0; n < ntiles; ++n) {
        ByteStream bs(DataBuffer(
            whole.getSubView(static_cast<Buffer::size_type>(tile_off[n]),
                             tile_len[n]),
            big_endian ? Endianness::big : Endianness::little));
        d.slices.emplace_back(d.dsc, n, bs);
      }
      const auto t0 = std::chrono::steady_clock::now();
      d.decompress();
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(
          best, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    if (best_ms)
      *best_ms = best;
    for (int r = 0; r < h; ++r)
      std::memcpy(static_cast<uint8_t*>(img_data) + static_cast<size_t>(r) * pitch, rowPtr(r),
                  static_cast<size_t>(w) * bpp);
  });
}

int ref_cr2_decompress(uint16_t* img_data, int w, int h, int pitch, int is_cfa,
                       int n_comp, int x_s_f, int y_s_f, int frame_w,
                       int frame_h, int num_slices, int slice_w,
                       int last_slice_w, const RefHuffDesc* tabs,
                       const int* tab_of_comp, const uint16_t* init_pred,
                       int nrec, const uint8_t* in, uint32_t in_size,
                       uint32_t* consumed, int reps, double* best_ms,
                       RefErr* e) {
  return guarded(e, [&] {
    RawImage img = makeImage(w, h, 1, is_cfa, 1, 1);
    copyIn(img, img_data, pitch);
    std::vector<std::unique_ptr<PrefixCodeDecoder<>>> hts;
    int maxTab = 0;
    for (int i = 0; i < nrec; ++i)
      maxTab = std::max(maxTab, tab_of_comp[i]);
    for (int t = 0; t <= maxTab; ++t)
      hts.emplace_back(std::make_unique<PrefixCodeDecoder<>>(
          makeHT(tabs[t].ncpl, tabs[t].values, tabs[t].nvalues, true, false)));
    using D = Cr2Decompressor<PrefixCodeDecoder<>>;
    std::vector<D::PerComponentRecipe> rec;
    rec.reserve(nrec);
    for (int i = 0; i < nrec; ++i)
      rec.push_back({*hts[tab_of_comp[i]], init_pred[i]});
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      D d(img, std::make_tuple(n_comp, x_s_f, y_s_f), iPoint2D(frame_w, frame_h),
          Cr2SliceWidths(num_slices, slice_w, last_slice_w), rec,
          Array1DRef<const uint8_t>(in, static_cast<int>(in_size)));
      const auto c = d.decompress();
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(
          best, std::chrono::duration<double, std::milli>(t1 - t0).count());
      if (consumed)
        *consumed = c;
    }
    if (best_ms)
      *best_ms = best;
    copyOut(img, img_data, pitch);
  });
}

int ref_cr2_ljpeg_decode(const uint8_t* in, uint32_t in_size,
                         uint16_t* img_data, int w, int h, int pitch, int is_cfa,
                         int sub_x, int sub_y, int num_slices, int slice_w,
                         int last_slice_w, int reps, double* best_ms,
                         RefErr* e) {
  return guarded(e, [&] {
    RawImage img = makeImage(w, h, 1, is_cfa, sub_x, sub_y);
    copyIn(img, img_data, pitch);
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      Cr2LJpegDecoder d(
          ByteStream(DataBuffer(Buffer(in, in_size), Endianness::little)), img);
      if (num_slices == 0 && slice_w == 0 && last_slice_w == 0)
        d.decode(Cr2SliceWidths());
      else
        d.decode(Cr2SliceWidths(num_slices, slice_w, last_slice_w));
      const auto t1 = std::chrono::steady_clock::now();
      best = std::min(
          best, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    if (best_ms)
      *best_ms = best;
    copyOut(img, img_data, pitch);
  });
}

} // extern "C"
