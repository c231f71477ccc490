/*
 * rs_oracle.h -- CPU restatement of rawspeed's per-pixel decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rawspeed_b200/ may include, link,
 * import or execute this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Every function cites the reference file:line (relative to
 * /root/reference/src/librawspeed) whose behaviour it restates.  The
 * restatement is pinned against (a) the reference's own unit-test vectors
 * (tests/golden/ JSON files, transcribed from test/librawspeed/...) and (b) the
 * real reference compiled into oracle/_ref (see oracle/Makefile), see
 * tests/test_oracle_vs_ref.py.
 *
 * Plain C99, no dependencies beyond libc (+ OpenMP for the tile fan-out that
 * restates AbstractDngDecompressor.cpp:240-252).
 */
#ifndef RS_ORACLE_H
#define RS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Exception classes of the reference (common/RawspeedException.h:33-95). */
enum { RSO_OK = 0, RSO_RDE = 1 /* RawDecoderException */, RSO_IOE = 2 /* IOException */ };

/* bitstreams/BitStreams.h:28-35 (same numeric values as enum class BitOrder) */
enum { RSO_LSB = 0, RSO_MSB = 1, RSO_MSB16 = 2, RSO_MSB32 = 3, RSO_JPEG = 4 };

typedef struct {
  int code;      /* RSO_OK / RSO_RDE / RSO_IOE */
  char msg[240]; /* what() */
} rso_err;

/* uint16 image view == RawImageData (common/RawImage.cpp:68-113): `pitch` bytes
 * between rows, `w` pixels of `cpp` components each. */
typedef struct {
  uint16_t* data;
  int w, h, cpp;
  int pitch; /* bytes */
  int is_cfa; /* RawImageData::isCFA, only consulted by the CR2 path */
  int sub_x, sub_y; /* ImageMetaData::subsampling (RawImage.h:93), default 1,1 */
  int is_f32; /* RawImageType::F32: `data` holds 32-bit samples (only the
                 uncompressed paths accept it) */
} rso_image;

/* createData(): pitch = roundUp(w*cpp*2, 16) (RawImage.cpp:68-84) */
int rso_image_pitch(int w, int cpp);

/* ---- bit pumps (bitstreams/BitStreamer.h:135-326, BitStream.h:59-141) ---- */
/* Probe used by the golden-vector tests: construct pump of `order` over
 * data[0..size) and perform n getBits(lens[i]) calls -> out[i].
 * Returns RSO_OK or the exception class (message in e). */
int rso_pump_getbits(int order, const uint8_t* data, int size, const int* lens,
                     int n, uint32_t* out, rso_err* e);
/* Like above, but reports getStreamPosition() after the reads
 * (BitStreamer.h:229-232, BitStreamerJPEG.h:185-189). */
int rso_pump_getbits_pos(int order, const uint8_t* data, int size,
                         const int* lens, int n, uint32_t* out, int* stream_pos,
                         rso_err* e);

/* ---- Huffman (codes/HuffmanCode.h:66-166, PrefixCodeLUTDecoder.h:95-216,
 *               PrefixCodeLookupDecoder.h:97-164, AbstractPrefixCodeDecoder.h:43-76) */
typedef struct rso_huff rso_huff;
/* ncpl[16] = DHT counts for lengths 1..16, values[nvalues]. */
rso_huff* rso_huff_create(const uint8_t ncpl[16], const uint8_t* values,
                          int nvalues, int full_decode, int fix_dng16,
                          rso_err* e);
void rso_huff_destroy(rso_huff*);
/* generateCodeSymbols(): writes up to 162 (code,len) pairs; returns count. */
int rso_huff_symbols(const rso_huff*, uint16_t* codes, uint8_t* lens);
/* decodeDifference()/decodeCodeValue() n times from a JPEG (order=RSO_JPEG) or
 * MSB pump over data. */
int rso_huff_decode(const rso_huff*, int order, const uint8_t* data, int size,
                    int n, int32_t* out, rso_err* e);
/* extend() truth table probe (AbstractPrefixCodeDecoder.h:68-76) */
int rso_huff_extend(uint32_t diff, uint32_t len);

/* ---- UncompressedDecompressor (decompressors/UncompressedDecompressor.cpp:106-268) */
int rso_unpack(const uint8_t* in, uint32_t in_size, rso_image* img, int crop_x,
               int crop_y, int crop_w, int crop_h, int in_pitch, int bps,
               int order, rso_err* e);

/* The other members of UncompressedDecompressor (same constructor, then one of):
 *   RSO_FORM_READ               readUncompressedRaw() (with is_f32: the F32 image
 *                               branches, UncompressedDecompressor.cpp:214-247:
 *                               32-bit row copy, decodePackedFP<MSB/LSB, Binary16/24>)
 *   RSO_FORM_8BIT[_UNCORRECTED] decode8BitRaw<false/true>()            (:270-294)
 *   RSO_FORM_12BIT_CONTROL_*    decode12BitRawWithControl<big/little>() (:299-359)
 *   RSO_FORM_12BIT_LEFT_*       decode12BitRawUnpackedLeftAligned<e>()  (:366-390)
 * For an F32 image `img->data` points at 32-bit samples (pitch in bytes, bpp 4*cpp).
 * `table` = TableLookUp::tables of table 0 (65536 entries, or 2*65536 when
 * dithered: common/TableLookUp.cpp:40-85), NULL = RawImageData::table == nullptr. */
enum {
  RSO_FORM_READ = 0,
  RSO_FORM_8BIT = 1,
  RSO_FORM_8BIT_UNCORRECTED = 2,
  RSO_FORM_12BIT_CONTROL_BE = 3,
  RSO_FORM_12BIT_CONTROL_LE = 4,
  RSO_FORM_12BIT_LEFT_BE = 5,
  RSO_FORM_12BIT_LEFT_LE = 6
};
int rso_unpack_form(const uint8_t* in, uint32_t in_size, rso_image* img, int is_f32,
                    int crop_x, int crop_y, int crop_w, int crop_h, int in_pitch, int bps,
                    int order, int form, const uint16_t* table, int table_dither,
                    rso_err* e);

/* ---- LJpegDecompressor (decompressors/LJpegDecompressor.cpp:52-370) ---- */
typedef struct {
  int mcu_x, mcu_y; /* Frame::mcu */
  int dim_x, dim_y; /* Frame::dim (in MCUs) */
} rso_ljpeg_frame;

int rso_ljpeg_decompress(rso_image* img, int fx, int fy, int fw, int fh,
                         rso_ljpeg_frame frame, const rso_huff* const* ht,
                         const uint16_t* init_pred, int nrec,
                         int rows_per_restart, const uint8_t* in, uint32_t in_size,
                         uint32_t* consumed, rso_err* e);

/* ---- LJpegDecoder::decode (LJpegDecoder.cpp:66-165 + AbstractLJpegDecoder.cpp:65-291)
 * Parses a complete SOI..EOI LJPEG blob and decodes it into the tile. */
int rso_ljpeg_decode(const uint8_t* in, uint32_t in_size, rso_image* img,
                     uint32_t off_x, uint32_t off_y, uint32_t w, uint32_t h,
                     int max_w, int max_h, int fix_dng16, rso_err* e);

/* ---- AbstractDngDecompressor::decompress (AbstractDngDecompressor.cpp:54-131,240-252)
 * compression 1 (uncompressed) and 7 (LJPEG).  tile_off/tile_len index `file`.
 * `big_endian` = byte order of the tile ByteStreams (only matters for comp 1,
 * bps 8/16/32).  nthreads = rawspeed_get_number_of_processor_cores(). */
int rso_dng_decompress(const uint8_t* file, uint64_t file_size,
                       const uint64_t* tile_off, const uint32_t* tile_len,
                       int ntiles, rso_image* img, int tile_w, int tile_h,
                       int compression, int fix_ljpeg, int bps, int big_endian,
                       int nthreads, rso_err* e);

/* ---- PentaxDecompressor (decompressors/PentaxDecompressor.cpp:55-177) ----
 * meta == NULL: SetupPrefixCodeDecoder_Legacy (the built-in pentax_tree); else the
 * "modern" table description read from `meta` (:83-141; meta_be = byte order of that ByteStream).  Decodes `data` (plain MSB
 * bit stream) into the whole image; predictor = same-parity pixel two to the left,
 * row starts from two rows up (:158-176). */
int rso_pentax_decompress(rso_image* img, const uint8_t* meta, int meta_size, int meta_be,
                          const uint8_t* data, uint32_t size, rso_err* e);
/* Table the constructor would build: fills ncpl[16], values[<=16]; returns the
 * number of codes (<0 and e set on a throw). */
int rso_pentax_table(const uint8_t* meta, int meta_size, int meta_be, uint8_t* ncpl,
                     uint8_t* values, rso_err* e);
/* Writer for tests: Huffman-encode `diffs` with ONE table into a plain MSB stream
 * (no stuffing), zero-padded to a multiple of 4 bytes + 16 zero bytes. */
int64_t rso_encode_diffs_plain(const int32_t* diffs, uint64_t n, const rso_huff* ht,
                               uint8_t* out, uint64_t cap);

/* ---- NikonDecompressor (decompressors/NikonDecompressor.cpp:380-562) ----
 * Constructor (meta = the maker-note ByteStream, meta_be its byte order, bitsPS 12/14):
 * version bytes, huffSelect, the four start predictors pUp, createCurve (:380-441),
 * split.  decompress (:513-560): plain MSB bit stream, nikon_tree[huffSelect] full-decode
 * table, per-parity left predictor with rows starting from pUp[row & 1] (updated by
 * the first two pixels of every row), clampBits(value, 15), setWithLookUp with the
 * curve as a DITHERED table (RawImageCurveGuard) unless uncorrected != 0; the dither
 * state is seeded ONCE with the first 24 bits of the stream.
 * With a non-zero split the rows from `split` on are decoded with nikon_tree[huffSelect+1]
 * through the restated NikonLASDecompressor (:80-378: own table builder, 8/14-bit lookups,
 * (len | shl << 4) difference format).
 * Outputs (may be NULL): curve[<= 32769] + *ncurve, pup[4] = pUp[0][0], pUp[0][1],
 * pUp[1][0], pUp[1][1], *huff_select, *split. */
int rso_nikon_setup(const uint8_t* meta, int meta_size, int meta_be, int bitsPS, int img_w,
                    int img_h, uint16_t* curve, int* ncurve, int* pup, int* huff_select,
                    int* split, rso_err* e);
int rso_nikon_decompress(rso_image* img, const uint8_t* meta, int meta_size, int meta_be,
                         int bitsPS, const uint8_t* data, uint32_t size, int uncorrected,
                         rso_err* e);
/* nikon_tree[sel] as (ncpl[16], values[<=16]); returns the number of codes */
int rso_nikon_tree(int sel, uint8_t* ncpl, uint8_t* values);

/* ---- PanasonicV5 / V6 / V7 Decompressor ----
 * version 5 (decompressors/PanasonicV5Decompressor.cpp:58-266): 0x4000-byte blocks whose
 *   two sections (split at 0x1FF8) are swapped, 16-byte packets of 10 x 12 or 9 x 14 bits
 *   (LSB first), pixels numbered linearly over the image;
 * version 6 (PanasonicV6Decompressor.cpp:70-263): 16-byte blocks of 14 (12 bit) or 11
 *   (14 bit) pixels with per-triplet scale and an odd/even running reference;
 * version 7 (PanasonicV7Decompressor.cpp:40-106): 16-byte blocks of 9 x 14 bits.
 * bps: 12 or 14 (ignored for version 7). */
int rso_panasonic(int version, rso_image* img, const uint8_t* data, uint32_t size, int bps,
                  rso_err* e);

/* ---- HasselbladDecompressor (decompressors/HasselbladDecompressor.cpp:39-100) ----
 * (groundwork for the next "next" row: restated and pinned, no device kernel yet)
 * One MSB32 bit stream per image; pixels are packed two at a time:
 * [len1 code][len2 code][len1 bits][len2 bits]; `ht` must be a code-value table (built with
 * full = 0); each row starts from init_pred for both pixels of a pair; values are stored
 * truncated to 16 bits.  *consumed = BitStreamerMSB32::getStreamPosition(). */
int rso_hasselblad_decompress(rso_image* img, const rso_huff* ht, uint16_t init_pred,
                              const uint8_t* in, uint32_t in_size, uint32_t* consumed,
                              rso_err* e);

/* ---- PhaseOneDecompressor (decompressors/PhaseOneDecompressor.cpp:42-168) ----
 * One strip per image row (any order; strip k = bytes [off[k], off[k]+len[k]) of `file`,
 * decoding row rown[k]); a row is an MSB32 bit stream: every 8 pixels two code lengths
 * (unary prefix + 1 bit into {8,7,6,9,11,10,5,12,14,13}), then per pixel either a raw
 * 16-bit value (length 14) or a difference to the same-parity predecessor. */
int rso_phaseone(rso_image* img, const uint8_t* file, uint64_t file_size, const uint64_t* off,
                 const uint32_t* len, const int32_t* rown, int nstrips, rso_err* e);

/* ---- PanasonicV4Decompressor (decompressors/PanasonicV4Decompressor.cpp:49-277) ----
 * (groundwork: restated and pinned, no device kernel yet)
 * 0x4000-byte blocks whose two sections (split at section_split_offset) are swapped; every
 * 16-byte packet is read from its top bit down and holds 14 pixels: 8-bit steps scaled by a
 * 2-bit shift per triplet, per-parity predictor, 12-bit restarts.  zero_pos (may be NULL):
 * (row << 16 | col) of every zero pixel when !zero_is_not_bad, in pixel order; *nzero = how
 * many there are (only the first `cap` are stored). */
int rso_panasonic_v4(rso_image* img, const uint8_t* data, uint32_t size, int zero_is_not_bad,
                     uint32_t section_split_offset, uint32_t* zero_pos, uint32_t cap,
                     uint32_t* nzero, rso_err* e);

/* ---- RawImageDataU16::scaleValues (common/RawImageDataU16.cpp:185-399) ----
 * (groundwork for SURVEY 8(f)3: restated and pinned, no device kernel yet)
 * The black/white scaling of scaleBlackWhite() once black_sep[4] (blackLevelSeparate, index
 * 2*(row&1) + (col&1) in crop coordinates) and white are known.  img = the UNCROPPED buffer,
 * (off_x, off_y, crop_w, crop_h) = mOffset / dim.  sse2 != 0 restates scaleValues_SSE2
 * (:204-341, what x86 builds run when app_scale < 63: 10-bit fixed point, eight 16-bit
 * multiplicative dither states per row, whole uncropped rows in groups of 8 columns);
 * sse2 == 0 restates scaleValues_plain (:343-399: 14-bit fixed point, one
 * multiply-with-carry dither state per row, cropped columns only).  The reference picks:
 * SSE2 iff 65535 / (white - black_sep[0]) < 63 (rso_scale_uses_sse2). */
/* RawImageDataU16::scaleBlackWhite (:147-183) + calculateBlackAreas (:60-145): black_level =
 * RawImageData::blackLevel (-1 unset); has_sep / black_sep[4] = blackLevelSeparate (in when
 * has_sep, always out); has_white / *white = whitePoint (in when has_white, always out);
 * areas = blackAreas.  Estimates black/white from the crop's centre when they are unknown,
 * takes the per-CFA-position median of the masked areas (16-bit histogram counters and the
 * single sampled column / row of the FIXMEs included), then scales (force_sse2: -1 = the
 * reference's choice).  Returns RSO_OK also when the reference returns without scaling. */
typedef struct {
  uint32_t offset, size; /* BlackArea (metadata/BlackArea.h:27-34) */
  int is_vertical;
} rso_black_area;
int rso_scale_black_white(rso_image* img, int off_x, int off_y, int crop_w, int crop_h,
                          int black_level, int* black_sep, int has_sep, int* white, int has_white,
                          const rso_black_area* areas, int n_areas, int dither, int force_sse2,
                          rso_err* e);
int rso_scale_uses_sse2(const int* black_sep, int white);
int rso_scale_values(rso_image* img, int off_x, int off_y, int crop_w, int crop_h,
                     const int* black_sep, int white, int dither, int sse2, rso_err* e);

/* ---- DngOpcodes (common/DngOpcodes.cpp:62-798) ----
 * (restated and pinned; device pass = K10)
 * DngOpcodes(ri, bs) -- parse and validate the big-endian opcode list against the image and
 * its current crop -- then applyOpCodes(ri): FixBadPixelsConstant (4), FixBadPixelsList (5),
 * TrimBounds (6), MapTable (7), MapPolynomial (8), DeltaPerRow / Column (10 / 11),
 * ScalePerRow / Column (12 / 13) on uint16 or float images; 1, 2, 3, 9 are known but
 * unsupported (an error unless flagged optional).  img = the UNCROPPED buffer (uint16, or
 * 32-bit samples holding floats when is_f32); crop[4] = mOffset.x, mOffset.y, dim.x, dim.y,
 * updated by TrimBounds.  bad / *nbad: mRaw->mBadPixelPositions afterwards, in the reference's
 * order (starts empty; at most bad_cap entries are stored, *nbad is the full count).
 * An error raised by an opcode's setup()/apply() leaves the earlier opcodes applied, as in the
 * reference; *applied (may be NULL) = opcodes fully applied. */
int rso_dng_opcodes(rso_image* img, int* crop, const uint8_t* data, uint32_t size, uint32_t* bad,
                    uint32_t bad_cap, uint32_t* nbad, int* applied, rso_err* e);

/* ---- RawImageData::fixBadPixels (common/RawImage.cpp:201-239, :297-323;
 *      RawImageDataU16::fixBadPixel common/RawImageDataU16.cpp:399-485) ----
 * (restated and pinned; device pass = K11)
 * positions = mBadPixelPositions ((y << 16) | x, uncropped coordinates).  Every bad pixel in
 * the first ((w + 15) / 32) * 32 columns is replaced by the distance-weighted mean of the
 * nearest good pixels to the left / right / above / below, at step 2 for CFA images and 1
 * otherwise; for cpp > 1 the components are read at column x + component (not x * cpp +
 * component), as the reference does. */
int rso_fix_bad_pixels(rso_image* img, const uint32_t* positions, uint32_t npositions, rso_err* e);

/* ---- RawImageData::sixteenBitLookup (common/RawImage.cpp:373-378) + RawImageDataU16::doLookup
 *      (common/RawImageDataU16.cpp:487-520) ----
 * (restated and pinned; device pass = K12)
 * The APPLY_LOOKUP worker carries the FULL_IMAGE flag (common/RawImage.h:58-63, RawImage.cpp:
 * 272-279): every row of the UNCROPPED buffer, all w * cpp samples of each; `table` = the
 * storage build_table() / TableLookUp::setTable makes (65536 entries, or 2 * 65536 {base,
 * delta} when dithered: v = 15700 * (v & 65535) + (v >> 16) seeded (w + 13 y) ^ 0x45694584 per
 * row, pix = base + ((delta * (v & 2047) + 1024) >> 12)). */
int rso_sixteen_bit_lookup(rso_image* img, const uint16_t* table, int dither, rso_err* e);

/* ---- SonyArw2Decompressor (decompressors/SonyArw2Decompressor.cpp:41-150) ----
 * One byte per pixel: every row is an LSB-first bit stream of 128-bit blocks; a block
 * carries max(11) min(11) imax(4) imin(4) + 14 x 7-bit deltas for 16 same-parity
 * pixels (the blocks of the even and the odd pixels of 32 columns alternate).  Every
 * value goes through RawImageDataU16::setWithLookUp (common/RawImage.h:335-353) with
 * a per-row dither state seeded from the row's first 24 bits.
 * `table`/`table_dither` as in rso_unpack_form (NULL = image has no table). */
int rso_sony_arw2(rso_image* img, const uint8_t* data, uint32_t size, const uint16_t* table,
                  int table_dither, rso_err* e);

/* ---- Cr2sRawInterpolator (interpolators/Cr2sRawInterpolator.cpp:32-544) ----
 * in: the subsampled image as decoded (in_w uint16 per row = 4 or 6 per MCU);
 * out: 3-component image (out->sub_x/sub_y = ImageMetaData::subsampling selects
 * 4:2:2 (2,1) or 4:2:0 (2,2)); version 0..2 = the YUV_TO_RGB<version> variants. */
int rso_sraw_interpolate(const uint16_t* in, int in_w, int in_h, int in_pitch,
                         rso_image* out, const int* sraw_coeffs, int hue, int version,
                         rso_err* e);

/* ---- Cr2Decompressor (decompressors/Cr2DecompressorImpl.h:279-468) ---- */
int rso_cr2_decompress(rso_image* img, int n_comp, int x_s_f, int y_s_f,
                       int frame_w, int frame_h, int num_slices, int slice_w,
                       int last_slice_w, const rso_huff* const* ht,
                       const uint16_t* init_pred, int nrec, const uint8_t* in,
                       uint32_t in_size, uint32_t* consumed, rso_err* e);

/* ---- Cr2LJpegDecoder::decode (Cr2LJpegDecoder.cpp:58-167) ---- */
int rso_cr2_ljpeg_decode(const uint8_t* in, uint32_t in_size, rso_image* img,
                         int num_slices, int slice_w, int last_slice_w,
                         rso_err* e);

/* ===== test-input tooling (NOT part of the decode path) =====
 * The reference's writer half (BitVacuumerJPEG.h:44-96,
 * PrefixCodeVectorEncoder.h:79-90, AbstractPrefixCodeEncoder.h:47-58) is
 * restated so synthetic LJPEG/CR2 streams can be produced on the GPU box. */

/* Encode the entropy-coded segment for `n` differences (full decode tables);
 * comp_of[i % group] selects ht for sample i.  Output is FF00-stuffed, padded
 * with 1-bits to a byte.  Returns bytes written or <0 if cap too small. */
int64_t rso_encode_diffs(const int32_t* diffs, uint64_t n,
                         const rso_huff* const* ht, const uint8_t* comp_of,
                         int group, uint8_t* out, uint64_t cap);

/* Build a complete LJPEG blob (SOI,SOF3,DHT...,[DRI],SOS,data,EOI) for a tile.
 * samples: tile_rows x (frame_w*ncomp) uint16 (MCU mcu_x x mcu_y, ncomp=mcu_x*mcu_y;
 * for mcu_y==2 two image rows per LJPEG row).  tables: ntab (ncpl,values) pairs;
 * tab_of_comp selects the table id per component.  restart_rows: 0 = no DRI.
 * Returns bytes written, <0 on overflow. */
typedef struct {
  uint8_t ncpl[16];
  uint8_t values[162];
  int nvalues;
} rso_dht;

int64_t rso_ljpeg_encode(const uint16_t* samples, int src_pitch_elems,
                         int frame_w, int frame_h, int mcu_x, int mcu_y,
                         int prec, const rso_dht* tabs, int ntab,
                         const uint8_t* tab_of_comp, int restart_rows,
                         int fix_dng16, uint8_t* out, uint64_t cap);

/* Build a CR2-style LJPEG blob whose scan holds the image in Canon slice order
 * (inverse of Cr2DecompressorImpl.h:396-468) for formats <2,1,1>/<4,1,1>/<3,2,1>/<3,2,2>. */
int64_t rso_cr2_encode(const rso_image* img, int n_comp, int x_s_f, int y_s_f,
                       int frame_w, int frame_h, int num_slices, int slice_w,
                       int last_slice_w, int prec, const rso_dht* tabs, int ntab,
                       const uint8_t* tab_of_comp, uint8_t* out, uint64_t cap);


/* Synthetic frame of SURVEY 8(d) C3 (test inputs; same pixels as synth.image_model):
 * px(x,y) = (2000 + ((7x+3y)&1023) + noise6 - 32) & 0x3FFF, noise6 = bits 31..26 of the LCG
 * s' = s*1664525 + 1013904223 stepped once per pixel in raster order from `seed`.
 * sums[0] = sum of the pixels, sums[1] = sum of pixel * (((31x + 17y) & 0xFFFF) | 1), both mod 2^64. */
void rso_image_model(uint32_t w, uint32_t h, uint32_t seed, uint16_t* out, uint64_t sums[2]);

#ifdef __cplusplus
}
#endif
#endif
