/*
 * rs_oracle.c -- CPU restatement of rawspeed's per-pixel decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see rs_oracle.h).  Never linked into the product.
 *
 * This is a from-scratch C99 restatement of the *algorithms* of the reference
 * (C++20).  Each block cites the reference file:line it follows (paths
 * relative to /root/reference/src/librawspeed).  "Exceptions" are modelled with
 * setjmp/longjmp so that the control flow (what is checked, in which order,
 * which exception class results) reads like the reference.
 *
 * Parity status: PINNED -- checked against the reference's unit-test vectors
 * (tests/golden/) and differentially against the compiled reference
 * (oracle/_ref/libref.so; tests/test_oracle_vs_ref.py).
 */
#include "rs_oracle.h"

#include <limits.h>
#include <setjmp.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* exceptions: common/RawspeedException.h:33-95 (ThrowRDE / ThrowIOE)  */
/* ------------------------------------------------------------------ */
typedef struct {
  jmp_buf jb;
  rso_err* e;
} rso_ctx;

static void
#if defined(__GNUC__)
    __attribute__((noreturn, format(printf, 3, 4)))
#endif
    rso_throw(rso_ctx* c, int code, const char* fmt, ...) {
  va_list ap;
  c->e->code = code;
  va_start(ap, fmt);
  vsnprintf(c->e->msg, sizeof c->e->msg, fmt, ap);
  va_end(ap);
  longjmp(c->jb, 1);
}
#define THROW_RDE(c, ...) rso_throw((c), RSO_RDE, __VA_ARGS__)
#define THROW_IOE(c, ...) rso_throw((c), RSO_IOE, __VA_ARGS__)

#define RSO_ENTER(ctx, e)                                                      \
  rso_ctx ctx;                                                                 \
  rso_err rso_local_err_;                                                      \
  ctx.e = (e) ? (e) : &rso_local_err_;                                         \
  ctx.e->code = RSO_OK;                                                        \
  ctx.e->msg[0] = 0;                                                           \
  if (setjmp(ctx.jb))                                                          \
  return ctx.e->code

int rso_image_pitch(int w, int cpp) {
  /* common/RawImage.cpp:80-82: roundUp(dim.x * bpp, 16) */
  long v = (long)w * cpp * 2;
  return (int)((v + 15) / 16 * 16);
}

/* ------------------------------------------------------------------ */
/* ByteStream (io/ByteStream.h:42-140, io/Buffer.h:47-121): bounds-checked
 * cursor; multi-byte reads big-endian here (AbstractLJpegDecoder.cpp:50
 * sets Endianness::big).                                              */
/* ------------------------------------------------------------------ */
typedef struct {
  rso_ctx* c;
  const uint8_t* data;
  uint32_t size;
  uint32_t pos;
} bstream;

static void bs_check(const bstream* s, uint64_t bytes) {
  /* ByteStream.h:62-69 */
  if ((uint64_t)s->pos + bytes > (uint64_t)s->size)
    THROW_IOE(s->c, "Out of bounds access in ByteStream");
}
static uint32_t bs_remain(const bstream* s) {
  bs_check(s, 0);
  return s->size - s->pos;
}
static uint8_t bs_peek_byte(const bstream* s, uint32_t i) {
  if ((uint64_t)s->pos + i + 1 > (uint64_t)s->size)
    THROW_IOE(s->c, "Out of bounds access in ByteStream");
  return s->data[s->pos + i];
}
static uint8_t bs_get_byte(bstream* s) {
  uint8_t v = bs_peek_byte(s, 0);
  s->pos += 1;
  return v;
}
static uint16_t bs_peek_u16be(const bstream* s) {
  bs_check(s, 2);
  return (uint16_t)((s->data[s->pos] << 8) | s->data[s->pos + 1]);
}
static uint16_t bs_get_u16be(bstream* s) {
  uint16_t v = bs_peek_u16be(s);
  s->pos += 2;
  return v;
}
static void bs_skip(bstream* s, uint64_t n) {
  bs_check(s, n);
  s->pos += (uint32_t)n;
}
static bstream bs_get_stream(bstream* s, uint32_t n) {
  bstream r;
  bs_check(s, n);
  r.c = s->c;
  r.data = s->data + s->pos;
  r.size = n;
  r.pos = 0;
  s->pos += n;
  return r;
}

/* ------------------------------------------------------------------ */
/* Bit pumps.                                                          */
/*  cache:       bitstreams/BitStream.h:59-141                          */
/*  replenisher: bitstreams/BitStreamer.h:42-132                        */
/*  streamer:    bitstreams/BitStreamer.h:135-326                       */
/*  traits:      BitStream{MSB,LSB,MSB16,MSB32,JPEG}.h:31-43            */
/*  JPEG fill:   bitstreams/BitStreamerJPEG.h:106-189                   */
/* ------------------------------------------------------------------ */
typedef struct {
  rso_ctx* c;
  const uint8_t* data;
  int size;
  int pos;        /* replenisher position (bytes handed to the cache) */
  uint64_t cache; /* BitStreamCacheBase::cache */
  int fill;       /* BitStreamCacheBase::fillLevel */
  int order;
  int end_pos; /* JPEG: PosOrUnknown endOfStreamPos (-1 = unknown) */
} pump;

static int pump_max_process_bytes(int order) {
  return order == RSO_JPEG ? 8 : 4; /* BitStreamer*.h MaxProcessBytes */
}

static void pump_init(pump* p, rso_ctx* c, int order, const uint8_t* data,
                      int size) {
  p->c = c;
  p->data = data;
  p->size = size;
  p->pos = 0;
  p->cache = 0;
  p->fill = 0;
  p->order = order;
  p->end_pos = -1;
  /* BitStreamer.h:56-60 */
  if (size < pump_max_process_bytes(order))
    THROW_IOE(c, "Bit stream size is smaller than MaxProcessBytes");
}

/* BitStreamCacheRightInLeftOut::push (BitStream.h:91-113) */
static void cache_push_msb(pump* p, uint64_t bits, int count) {
  if (count != 0)
    p->cache |= bits << (64 - p->fill - count);
  p->fill += count;
}
/* BitStreamCacheLeftInRightOut::push (BitStream.h:60-68) */
static void cache_push_lsb(pump* p, uint64_t bits, int count) {
  p->cache |= bits << p->fill;
  p->fill += count;
}

/* BitStreamerForwardSequentialReplenisher::getInput (BitStreamer.h:100-131) */
static void pump_get_input(pump* p, uint8_t* tmp, int mp) {
  if (p->pos + mp <= p->size) {
    memcpy(tmp, p->data + p->pos, (size_t)mp);
    return;
  }
  if (p->pos > p->size + 2 * mp)
    THROW_IOE(p->c, "Buffer overflow read in BitStreamer");
  /* adt/VariableLengthLoad.h:148-173: zero padded tail */
  memset(tmp, 0, (size_t)mp);
  {
    int from = p->pos < p->size ? p->pos : p->size;
    int to = from + mp < p->size ? from + mp : p->size;
    if (to > from)
      memcpy(tmp, p->data + from, (size_t)(to - from));
  }
}

static uint32_t ld_le32(const uint8_t* b) {
  return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) |
         ((uint32_t)b[3] << 24);
}
static uint32_t ld_be32(const uint8_t* b) {
  return (uint32_t)b[3] | ((uint32_t)b[2] << 8) | ((uint32_t)b[1] << 16) |
         ((uint32_t)b[0] << 24);
}

/* BitStreamerJPEG::fillCache (BitStreamerJPEG.h:106-183); returns bytes consumed */
static int pump_fill_cache_jpeg(pump* p, const uint8_t* in) {
  int i, pp = 0;
  if (in[0] != 0xFF && in[1] != 0xFF && in[2] != 0xFF && in[3] != 0xFF) {
    cache_push_msb(p, ld_be32(in), 32);
    return 4;
  }
  for (i = 0; i < 4; ++i) {
    const int numBytesNeeded = 4 - i;
    const uint8_t c0 = in[pp + 0];
    cache_push_msb(p, c0, 8);
    if (c0 != 0xFF) {
      pp += 1;
      continue;
    }
    if (in[pp + 1] == 0x00) { /* FF 00 -> data byte FF */
      pp += 2;
      continue;
    }
    /* FF xx, xx != 0: end of stream (:155-179) */
    p->end_pos = p->pos + pp;
    p->fill -= 8;
    p->cache &= ~((~0ULL) >> p->fill); /* fill is never 64 here */
    p->fill = 64;
    pp = (p->size - p->pos) + numBytesNeeded;
    break;
  }
  return pp;
}

/* BitStreamer::fill (BitStreamer.h:216-229) + per-order fillCache (:155-182) */
static void pump_fill(pump* p, int nbits) {
  uint8_t tmp[8];
  if (p->fill >= nbits)
    return;
  pump_get_input(p, tmp, pump_max_process_bytes(p->order));
  switch (p->order) {
  case RSO_MSB: /* BitStreamMSB.h: u32 big-endian chunk */
    cache_push_msb(p, ld_be32(tmp), 32);
    p->pos += 4;
    break;
  case RSO_MSB32: /* BitStreamMSB32.h: u32 little-endian chunk, MSB-first */
    cache_push_msb(p, ld_le32(tmp), 32);
    p->pos += 4;
    break;
  case RSO_MSB16: /* BitStreamMSB16.h: two u16 little-endian chunks */
    cache_push_msb(p, (uint32_t)tmp[0] | ((uint32_t)tmp[1] << 8), 16);
    cache_push_msb(p, (uint32_t)tmp[2] | ((uint32_t)tmp[3] << 8), 16);
    p->pos += 4;
    break;
  case RSO_LSB: /* BitStreamLSB.h: u32 little-endian chunk, LSB-first */
    cache_push_lsb(p, ld_le32(tmp), 32);
    p->pos += 4;
    break;
  default: /* RSO_JPEG */
    p->pos += pump_fill_cache_jpeg(p, tmp);
    break;
  }
}

static uint32_t pump_peek_nofill(const pump* p, int n) {
  if (p->order == RSO_LSB) /* BitStream.h:70-78 */
    return (uint32_t)p->cache & (n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u));
  return (uint32_t)(p->cache >> (64 - n)); /* BitStream.h:115-125 */
}
static void pump_skip_nofill(pump* p, int n) {
  if (n == 0)
    return;
  if (p->order == RSO_LSB)
    p->cache >>= n; /* BitStream.h:80-89 */
  else
    p->cache = n >= 64 ? 0 : p->cache << n; /* BitStream.h:127-137 */
  p->fill -= n;
}
static uint32_t pump_get_nofill(pump* p, int n) {
  uint32_t v = pump_peek_nofill(p, n);
  pump_skip_nofill(p, n);
  return v;
}
static uint32_t pump_get_bits(pump* p, int n) { /* BitStreamer.h:294-301 */
  pump_fill(p, n);
  return pump_get_nofill(p, n);
}
static void pump_skip_bytes(pump* p, int nbytes) { /* BitStreamer.h:305-325 */
  int rem = 8 * nbytes;
  for (; rem >= 32; rem -= 32) {
    pump_fill(p, 32);
    pump_skip_nofill(p, 32);
  }
  if (rem > 0) {
    pump_fill(p, rem);
    pump_skip_nofill(p, rem);
  }
}
static int pump_stream_position(const pump* p) {
  if (p->order == RSO_JPEG) /* BitStreamerJPEG.h:185-189 */
    return p->end_pos >= 0 ? p->end_pos : p->pos;
  return p->pos - (p->fill >> 3); /* BitStreamer.h:229-232 */
}

int rso_pump_getbits_pos(int order, const uint8_t* data, int size,
                         const int* lens, int n, uint32_t* out, int* stream_pos,
                         rso_err* e) {
  pump p;
  int i;
  RSO_ENTER(c, e);
  pump_init(&p, &c, order, data, size);
  for (i = 0; i < n; ++i)
    out[i] = pump_get_bits(&p, lens[i]);
  if (stream_pos)
    *stream_pos = pump_stream_position(&p);
  return RSO_OK;
}
int rso_pump_getbits(int order, const uint8_t* data, int size, const int* lens,
                     int n, uint32_t* out, rso_err* e) {
  return rso_pump_getbits_pos(order, data, size, lens, n, out, NULL, e);
}

/* ------------------------------------------------------------------ */
/* Huffman tables                                                      */
/* ------------------------------------------------------------------ */
#define HUF_MAXSYM 162 /* AbstractPrefixCode.h:56 MaxNumCodeValues */
#define LUT_DEPTH 11   /* PrefixCodeLUTDecoder.h:91 LookupDepth */

struct rso_huff {
  int nsym;
  int maxlen;        /* maxCodeLength() */
  uint32_t ncpl[17]; /* nCodesPerLength, 1-based */
  uint16_t code[HUF_MAXSYM];
  uint8_t len[HUF_MAXSYM];
  uint8_t val[HUF_MAXSYM];
  uint16_t maxCodeOL[17];
  uint16_t codeOffsetOL[17];
  int32_t lut[1 << LUT_DEPTH];
  int full, fix16;
  /* encoder side */
  int idx_of_val[256];
};

int rso_huff_extend(uint32_t diff, uint32_t len) {
  /* AbstractPrefixCodeDecoder.h:68-76 (T.81 Figure F.12) */
  int32_t ret = (int32_t)diff;
  if ((diff & (1u << (len - 1))) == 0)
    ret -= (int32_t)((1u << len) - 1u);
  return ret;
}

/* HuffmanCode::setNCodesPerLength (HuffmanCode.h:100-147): returns the code count */
static unsigned huff_validate_counts(rso_ctx* c, const uint8_t ncpl[16]) {
  unsigned l, count = 0, maxCodes = 2, maxlen = 0;
  for (l = 1; l <= 16; ++l) {
    if (ncpl[l - 1])
      maxlen = l;
    count += ncpl[l - 1];
  }
  if (maxlen == 0)
    THROW_RDE(c, "Codes-per-length table is empty");
  if (count > HUF_MAXSYM)
    THROW_RDE(c, "Too big code-values table");
  for (l = 1; l <= maxlen; ++l) {
    const unsigned nCodes = ncpl[l - 1];
    if (nCodes > (1U << l))
      THROW_RDE(c, "Corrupt Huffman. Can never have %u codes in %u-bit len",
                nCodes, l);
    if (nCodes > maxCodes)
      THROW_RDE(c,
                "Corrupt Huffman. Can only fit %u out of %u codes in %u-bit len",
                maxCodes, nCodes, l);
    maxCodes -= nCodes;
    maxCodes *= 2;
  }
  return count;
}

static void huff_build(rso_ctx* c, rso_huff* h, const uint8_t ncpl[16],
                       const uint8_t* values, int nvalues, int full,
                       int fix16) {
  unsigned l, i, count;
  uint32_t code;
  int n;
  memset(h, 0, sizeof *h);
  count = huff_validate_counts(c, ncpl);
  h->maxlen = 0;
  for (l = 1; l <= 16; ++l) {
    h->ncpl[l] = ncpl[l - 1];
    if (ncpl[l - 1])
      h->maxlen = (int)l;
  }
  /* HuffmanCode::setCodeValues (HuffmanCode.h:149-164); a DHT segment always
   * supplies exactly `count` values (AbstractLJpegDecoder.cpp:252-256). */
  if ((unsigned)nvalues != count)
    THROW_RDE(c, "Malformed code");
  h->nsym = (int)count;
  memcpy(h->val, values, count);
  /* generateCodeSymbols (HuffmanCode.h:66-93): T.81 Figures C.1/C.2 */
  code = 0;
  n = 0;
  for (l = 1; l <= (unsigned)h->maxlen; ++l) {
    for (i = 0; i < h->ncpl[l]; ++i) {
      h->code[n] = (uint16_t)code;
      h->len[n] = (uint8_t)l;
      ++n;
      ++code;
    }
    code <<= 1;
  }
  /* PrefixCode ctor (PrefixCode.h:50-67): non-empty, sizes match: by construction.
   * verifyCodeSymbols (:70-102): Kraft repeated above; ordering and prefix
   * freedom hold for canonical codes by construction. */
  /* AbstractPrefixCodeTranscoder::setup (AbstractPrefixCodeTranscoder.h:49-83) */
  h->full = full;
  h->fix16 = fix16;
  if (full) {
    for (i = 0; i < count; ++i)
      if (h->val[i] > 16)
        THROW_RDE(c, "Corrupt Huffman code: difference length %u longer than %u",
                  h->val[i], 16);
  }
  /* PrefixCodeLookupDecoder::setup (PrefixCodeLookupDecoder.h:97-113), T.81 F.15 */
  for (l = 0; l <= 16; ++l) {
    h->maxCodeOL[l] = 0xFFFF;
    h->codeOffsetOL[l] = 0xFFFF;
  }
  {
    unsigned soFar = 0;
    for (l = 1; l <= (unsigned)h->maxlen; ++l) {
      if (!h->ncpl[l])
        continue;
      h->codeOffsetOL[l] = (uint16_t)(h->code[soFar] - soFar);
      soFar += h->ncpl[l];
      h->maxCodeOL[l] = h->code[soFar - 1];
    }
  }
  /* PrefixCodeLUTDecoder::setup (PrefixCodeLUTDecoder.h:95-148) */
  for (i = 0; i < count; ++i) {
    const unsigned code_l = h->len[i];
    uint32_t ll, ul, cc;
    uint32_t diff_l = h->val[i];
    if (code_l > LUT_DEPTH)
      break;
    ll = (uint16_t)(h->code[i] << (LUT_DEPTH - code_l));
    ul = (uint16_t)(ll | ((1u << (LUT_DEPTH - code_l)) - 1u));
    for (cc = ll; cc <= ul; ++cc) {
      if (!(cc < (1u << LUT_DEPTH)))
        THROW_RDE(c, "Corrupt Huffman");
      if (!full || (code_l + diff_l > LUT_DEPTH && diff_l != 16)) {
        h->lut[cc] = (int32_t)(diff_l << 9 | code_l);
        if (!full)
          h->lut[cc] |= 0x100;
      } else {
        h->lut[cc] = (int32_t)(0x100 | code_l);
        if (diff_l != 16 || fix16)
          h->lut[cc] += (int32_t)diff_l;
        if (diff_l) {
          uint32_t diff;
          if (diff_l != 16) {
            diff = cc >> (LUT_DEPTH - (code_l + diff_l));
            diff &= ((1u << diff_l) - 1u);
          } else
            diff = (uint32_t)-32768;
          h->lut[cc] |=
              (int32_t)((uint32_t)rso_huff_extend(diff, diff_l) << 9);
        }
      }
    }
  }
  for (i = 0; i < 256; ++i)
    h->idx_of_val[i] = -1;
  for (i = 0; i < count; ++i) /* first match wins (PrefixCodeVectorEncoder.h:52-62) */
    if (h->idx_of_val[h->val[i]] < 0)
      h->idx_of_val[h->val[i]] = (int)i;
}

rso_huff* rso_huff_create(const uint8_t ncpl[16], const uint8_t* values,
                          int nvalues, int full_decode, int fix_dng16,
                          rso_err* e) {
  rso_huff* volatile h = NULL;
  rso_ctx c;
  rso_err le;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  h = (rso_huff*)malloc(sizeof(rso_huff));
  if (!h)
    return NULL;
  if (setjmp(c.jb)) {
    free(h);
    return NULL;
  }
  huff_build(&c, h, ncpl, values, nvalues, full_decode, fix_dng16);
  return h;
}
void rso_huff_destroy(rso_huff* h) { free(h); }

int rso_huff_symbols(const rso_huff* h, uint16_t* codes, uint8_t* lens) {
  memcpy(codes, h->code, sizeof(uint16_t) * (size_t)h->nsym);
  memcpy(lens, h->len, (size_t)h->nsym);
  return h->nsym;
}

/* PrefixCodeLUTDecoder::decode (PrefixCodeLUTDecoder.h:172-216) with the
 * PrefixCodeLookupDecoder::finishReadingPartialSymbol slow path
 * (PrefixCodeLookupDecoder.h:133-164) and
 * AbstractPrefixCodeDecoder::processSymbol (AbstractPrefixCodeDecoder.h:43-66). */
static int huff_decode(const rso_huff* h, pump* bs, int full) {
  uint32_t code, lutEntry;
  int payload, len, code_len, codeValue;
  pump_fill(bs, 32);
  code = pump_peek_nofill(bs, LUT_DEPTH);
  lutEntry = (uint32_t)h->lut[code];
  payload = (int32_t)lutEntry >> 9;
  len = (int)(lutEntry & 0xff);
  pump_skip_nofill(bs, len);
  if (lutEntry & 0x100)
    return payload;
  if (lutEntry) {
    code_len = len;
    codeValue = payload;
  } else {
    pump_skip_nofill(bs, LUT_DEPTH);
    code_len = LUT_DEPTH;
    while (code_len < h->maxlen && (0xFFFF == h->maxCodeOL[code_len] ||
                                    code > h->maxCodeOL[code_len])) {
      uint32_t t = pump_get_nofill(bs, 1);
      code = (uint16_t)((code << 1) | t);
      code_len++;
    }
    if (code_len > h->maxlen || code > h->maxCodeOL[code_len])
      THROW_RDE(bs->c, "bad Huffman code: %u (len: %u)", code,
                (unsigned)code_len);
    codeValue = h->val[code - h->codeOffsetOL[code_len]];
  }
  if (!full)
    return codeValue;
  if (codeValue == 16) {
    if (h->fix16)
      pump_skip_nofill(bs, 16);
    return -32768;
  }
  return codeValue ? rso_huff_extend(pump_get_nofill(bs, codeValue),
                                     (uint32_t)codeValue)
                   : 0;
}

int rso_huff_decode(const rso_huff* h, int order, const uint8_t* data, int size,
                    int n, int32_t* out, rso_err* e) {
  pump p;
  int i;
  RSO_ENTER(c, e);
  pump_init(&p, &c, order, data, size);
  for (i = 0; i < n; ++i)
    out[i] = huff_decode(h, &p, h->full);
  return RSO_OK;
}

/* ------------------------------------------------------------------ */
/* UncompressedDecompressor                                            */
/* ------------------------------------------------------------------ */
/* extendBinaryFloatingPoint<Narrow, Binary32> (common/FloatingPoint.h:116-160);
 * fw/ew = fraction/exponent widths of the narrow type */
static uint32_t fp_extend(uint32_t narrow, int fw, int ew) {
  uint32_t sign = (narrow >> (fw + ew)) & 1u;
  uint32_t ne = (narrow >> fw) & ((1u << ew) - 1u);
  uint32_t nf = narrow & ((1u << fw) - 1u);
  int32_t bias = (1 << (ew - 1)) - 1;
  uint32_t we = (uint32_t)((int32_t)ne - bias + 127);
  uint32_t wf = nf << (23 - fw);
  if (ne == ((1u << ew) - 1u)) {
    we = 255; /* infinity or NaN; the fraction is kept/widened */
  } else if (ne == 0) {
    if (nf == 0) {
      we = 0;
      wf = 0;
    } else {
      we = (uint32_t)(1 - bias + 127);
      while (!(wf & (1u << 23))) {
        we -= 1;
        wf <<= 1;
      }
      wf &= (1u << 23) - 1u;
    }
  }
  return (sign << 31) | (we << 23) | wf;
}

/* sanityCheck(const uint32_t* h, int bytesPerLine) (UncompressedDecompressor.cpp:52-74) */
static void unpack_sanity(rso_ctx* c, const bstream* input, uint32_t h, uint32_t bpl) {
  uint32_t fullRows = bs_remain(input) / bpl;
  if (fullRows >= h)
    return;
  if (fullRows == 0)
    THROW_IOE(c, "Not enough data to decode a single line. Image file truncated.");
  THROW_IOE(c, "Image truncated, only %u of %u lines found", fullRows, h);
}

static void unpack_impl(rso_ctx* c, const uint8_t* in_data, uint32_t in_size,
                        rso_image* img, int crop_x, int crop_y, int crop_w,
                        int crop_h, int inputPitchBytes, int bitPerPixel,
                        int order, int f32, int form, const uint16_t* table,
                        int table_dither) {
  /* ctor: UncompressedDecompressor.cpp:106-169 */
  bstream all, input;
  uint32_t w, h, cpp;
  uint64_t ox, oy, outPixelBits, outPixelBytes;
  uint32_t skipBytes;
  all.c = c;
  all.data = in_data;
  all.size = in_size;
  all.pos = 0;
  /* input_.getStream(crop.dim.y, inputPitchBytes_) (ByteStream.h:117-121) */
  {
    uint32_t nmemb = (uint32_t)crop_h, sz = (uint32_t)inputPitchBytes;
    if (sz && nmemb > UINT32_MAX / sz)
      THROW_IOE(c, "Integer overflow when calculating stream length");
    input = bs_get_stream(&all, nmemb * sz);
  }
  if (!(crop_w > 0 && crop_h > 0))
    THROW_RDE(c, "Empty tile.");
  if (inputPitchBytes < 1)
    THROW_RDE(c, "Input pitch is non-positive");
  if (order == RSO_JPEG)
    THROW_RDE(c, "JPEG bit order not supported.");
  w = (uint32_t)crop_w;
  h = (uint32_t)crop_h;
  cpp = (uint32_t)img->cpp;
  ox = (uint64_t)crop_x;
  oy = (uint64_t)crop_y;
  if (cpp < 1 || cpp > 3)
    THROW_RDE(c, "Unsupported number of components per pixel: %u", cpp);
  if (bitPerPixel < 1 || bitPerPixel > 32 || (bitPerPixel > 16 && !f32 /* UINT16 image */))
    THROW_RDE(c, "Unsupported bit depth");
  outPixelBits = (uint64_t)w * cpp * (uint64_t)bitPerPixel;
  if (outPixelBits % 8 != 0)
    THROW_RDE(c, "Bad combination of cpp (%u), bps (%d) and width (%u)", cpp,
              bitPerPixel, w);
  outPixelBytes = outPixelBits / 8;
  if ((uint64_t)(unsigned)inputPitchBytes < outPixelBytes)
    THROW_RDE(c, "Specified pitch is smaller than minimally-required pitch");
  /* sanityCheck(&h, inputPitchBytes) (:52-74) */
  {
    uint32_t fullRows = bs_remain(&input) / (uint32_t)inputPitchBytes;
    if (fullRows < h) {
      if (fullRows == 0)
        THROW_IOE(c, "Not enough data to decode a single line. Image file "
                     "truncated.");
      THROW_IOE(c, "Image truncated, only %u of %u lines found", fullRows, h);
    }
  }
  skipBytes = (uint32_t)((uint64_t)inputPitchBytes - outPixelBytes);
  if (oy > (uint64_t)img->h)
    THROW_RDE(c, "Invalid y offset");
  if (ox + (uint64_t)crop_w > (uint64_t)img->w)
    THROW_RDE(c, "Invalid x offset");

  if (form != RSO_FORM_READ) {
    /* the fixed-layout members; all write out(row, col) from (0,0) and use only
     * `size` (UncompressedDecompressor.cpp:270-390) */
    uint32_t row, col;
    if (form == RSO_FORM_8BIT || form == RSO_FORM_8BIT_UNCORRECTED) {
      /* decode8BitRaw<uncorrected> (:270-294); setWithLookUp (RawImage.h:335-353) */
      uint32_t random = 0;
      unpack_sanity(c, &input, h, 1u * w);
      if ((uint64_t)w * h > bs_remain(&input))
        THROW_IOE(c, "Buffer overflow: image file may be truncated");
      for (row = 0; row < h; row++) {
        uint16_t* o = (uint16_t*)((uint8_t*)img->data + (size_t)row * (size_t)img->pitch);
        for (col = 0; col < w; col++) {
          uint16_t value = input.data[(size_t)row * w + col];
          if (form == RSO_FORM_8BIT_UNCORRECTED || !table) {
            o[col] = value;
          } else if (table_dither) {
            uint32_t base = table[2 * value + 0], delta = table[2 * value + 1];
            uint32_t r = random;
            uint32_t pix = base + ((delta * (r & 2047) + 1024) >> 12);
            random = 15700 * (r & 65535) + (r >> 16);
            o[col] = (uint16_t)pix;
          } else {
            o[col] = table[value];
          }
        }
      }
      return;
    }
    if (form == RSO_FORM_12BIT_CONTROL_BE || form == RSO_FORM_12BIT_CONTROL_LE) {
      /* decode12BitRawWithControl<e> (:299-359), bytesPerLine (:86-104) */
      const int little = form == RSO_FORM_12BIT_CONTROL_LE;
      uint32_t perline, x;
      if ((12 * w) % 8 != 0)
        THROW_IOE(c, "Bad image width");
      perline = (12 * w) / 8 + ((w + 2) / 10);
      unpack_sanity(c, &input, h, perline);
      if ((uint64_t)perline * h > bs_remain(&input))
        THROW_IOE(c, "Buffer overflow: image file may be truncated");
      for (row = 0; row < h; row++) {
        const uint8_t* in = input.data + (size_t)row * perline;
        uint16_t* o = (uint16_t*)((uint8_t*)img->data + (size_t)row * (size_t)img->pitch);
        col = 0;
        for (x = 0; x < w; x += 2) {
          uint32_t g1 = in[col + 0], g2 = in[col + 1], g3 = in[col + 2];
          /* process(x, invert=false, g1, g2); process(x+1, invert=true, g3, g2) */
          if (!little) {
            o[x] = (uint16_t)((g1 << 4) | (g2 >> 4));
            o[x + 1] = (uint16_t)(((g2 & 0x0f) << 8) | g3);
          } else {
            o[x] = (uint16_t)(((g2 & 0x0f) << 8) | g1);
            o[x + 1] = (uint16_t)((g3 << 4) | (g2 >> 4));
          }
          col += 3;
          if ((x % 10) == 8)
            col++;
        }
      }
      return;
    }
    if (form == RSO_FORM_12BIT_LEFT_BE || form == RSO_FORM_12BIT_LEFT_LE) {
      /* decode12BitRawUnpackedLeftAligned<e> (:366-390) */
      unpack_sanity(c, &input, h, 2u * w);
      if ((uint64_t)w * h * 2 > bs_remain(&input))
        THROW_IOE(c, "Buffer overflow: image file may be truncated");
      for (row = 0; row < h; row++) {
        const uint8_t* in = input.data + (size_t)row * 2 * w;
        uint16_t* o = (uint16_t*)((uint8_t*)img->data + (size_t)row * (size_t)img->pitch);
        for (col = 0; col < w; ++col) {
          uint32_t g1 = in[2 * col], g2 = in[2 * col + 1];
          uint16_t pix = form == RSO_FORM_12BIT_LEFT_LE ? (uint16_t)((g2 << 8) | g1)
                                                        : (uint16_t)((g1 << 8) | g2);
          o[col] = pix >> 4;
        }
      }
      return;
    }
    THROW_RDE(c, "unknown form");
  }

  if (f32) {
    /* readUncompressedRaw, F32 image (:214-247) */
    uint64_t y = oy, hh = h + oy;
    int rows, row;
    if (hh > (uint64_t)img->h)
      hh = (uint64_t)img->h;
    rows = (int)hh;
    row = (int)y;
    if (bitPerPixel == 32) {
      int r;
      uint64_t need = (uint64_t)inputPitchBytes * (uint64_t)(rows - row);
      if (need > bs_remain(&input))
        THROW_IOE(c, "Buffer overflow: image file may be truncated");
      for (r = row; r < rows; ++r)
        memcpy((uint8_t*)img->data + (size_t)r * (size_t)img->pitch +
                   (size_t)crop_x * cpp * 4,
               input.data + (size_t)(r - row) * (size_t)inputPitchBytes,
               (size_t)w * cpp * 4);
      return;
    }
    if ((order == RSO_MSB || order == RSO_LSB) && (bitPerPixel == 16 || bitPerPixel == 24)) {
      /* decodePackedFP<Pump, Binary16/24> (:171-186): out(row, offset.x + col) */
      pump bits;
      int cols = crop_w * (int)cpp, x;
      pump_init(&bits, c, order, input.data, (int)bs_remain(&input));
      for (; row < rows; row++) {
        uint32_t* o = (uint32_t*)((uint8_t*)img->data + (size_t)row * (size_t)img->pitch);
        for (x = 0; x < cols; x++) {
          uint32_t b = pump_get_bits(&bits, bitPerPixel);
          o[crop_x + x] = bitPerPixel == 16 ? fp_extend(b, 10, 5) : fp_extend(b, 16, 7);
        }
        pump_skip_bytes(&bits, (int)skipBytes);
      }
      return;
    }
    THROW_RDE(c, "Unsupported floating-point input bitwidth/bit packing: %d / %u",
              bitPerPixel, (unsigned)order);
  }

  /* readUncompressedRaw: UncompressedDecompressor.cpp:202-268 */
  {
    uint64_t y = oy;
    uint64_t hh = h + oy;
    int rows, row;
    if (hh > (uint64_t)img->h)
      hh = (uint64_t)img->h;
    rows = (int)hh;
    row = (int)y;
    if (order == RSO_LSB && bitPerPixel == 16) {
      /* copyPixels (:255-264): row memcpy, honours offset.x */
      int r;
      uint64_t need = (uint64_t)inputPitchBytes * (uint64_t)(rows - row);
      if (need > bs_remain(&input))
        THROW_IOE(c, "Buffer overflow: image file may be truncated");
      for (r = row; r < rows; ++r)
        memcpy((uint8_t*)img->data + (size_t)r * (size_t)img->pitch +
                   (size_t)crop_x * cpp * 2,
               input.data + (size_t)(r - row) * (size_t)inputPitchBytes,
               (size_t)w * cpp * 2);
      return;
    }
    /* decodePackedInt<Pump> (:188-200): NOTE writes out(row, x), i.e. the
     * crop's x offset is ignored for packed integer data. */
    {
      pump bits;
      int cols = crop_w * (int)cpp, x;
      pump_init(&bits, c, order, input.data, (int)bs_remain(&input));
      for (; row < rows; row++) {
        uint16_t* o =
            (uint16_t*)((uint8_t*)img->data + (size_t)row * (size_t)img->pitch);
        for (x = 0; x < cols; x++)
          o[x] = (uint16_t)pump_get_bits(&bits, bitPerPixel);
        pump_skip_bytes(&bits, (int)skipBytes);
      }
    }
  }
}

int rso_unpack(const uint8_t* in, uint32_t in_size, rso_image* img, int crop_x,
               int crop_y, int crop_w, int crop_h, int in_pitch, int bps,
               int order, rso_err* e) {
  RSO_ENTER(c, e);
  unpack_impl(&c, in, in_size, img, crop_x, crop_y, crop_w, crop_h, in_pitch,
              bps, order, 0, RSO_FORM_READ, NULL, 0);
  return RSO_OK;
}

int rso_unpack_form(const uint8_t* in, uint32_t in_size, rso_image* img, int is_f32,
                    int crop_x, int crop_y, int crop_w, int crop_h, int in_pitch, int bps,
                    int order, int form, const uint16_t* table, int table_dither,
                    rso_err* e) {
  RSO_ENTER(c, e);
  unpack_impl(&c, in, in_size, img, crop_x, crop_y, crop_w, crop_h, in_pitch,
              bps, order, is_f32, form, table, table_dither);
  return RSO_OK;
}

/* ------------------------------------------------------------------ */
/* LJpegDecompressor                                                   */
/* ------------------------------------------------------------------ */
static uint32_t ljpeg_decompress_impl(rso_ctx* c, rso_image* img, int fx, int fy,
                                      int fw, int fh, rso_ljpeg_frame frame,
                                      const rso_huff* const* ht,
                                      const uint16_t* init_pred, int nrec,
                                      int rowsPerRestart, const uint8_t* in,
                                      uint32_t in_size) {
  /* ctor: LJpegDecompressor.cpp:52-152 */
  const int cpp = img->cpp;
  int tileRequiredWidth, numFullMCUs, trailingPixels, i;
  if (cpp < 1 || cpp > 3)
    THROW_RDE(c, "Unexpected component count (%u)", (unsigned)cpp);
  if (!(img->w > 0 && img->h > 0))
    THROW_RDE(c, "Image has zero size");
  if (!(fw > 0 && fh > 0))
    THROW_RDE(c, "Tile has zero size");
  if (fx >= img->w)
    THROW_RDE(c, "X offset outside of image");
  if (fy >= img->h)
    THROW_RDE(c, "Y offset outside of image");
  if (fw > img->w)
    THROW_RDE(c, "Tile wider than image");
  if (fh > img->h)
    THROW_RDE(c, "Tile taller than image");
  if (fx + fw > img->w)
    THROW_RDE(c, "Tile overflows image horizontally");
  if (fy + fh > img->h)
    THROW_RDE(c, "Tile overflows image vertically");
  if (!(frame.dim_x > 0 && frame.dim_y > 0))
    THROW_RDE(c, "Frame has zero size");
  if (!((frame.mcu_x == 1 && frame.mcu_y == 1) ||
        (frame.mcu_x == 2 && frame.mcu_y == 1) ||
        (frame.mcu_x == 3 && frame.mcu_y == 1) ||
        (frame.mcu_x == 4 && frame.mcu_y == 1) ||
        (frame.mcu_x == 2 && frame.mcu_y == 2)))
    THROW_RDE(c, "Unexpected MCU size: {%i, %i}", frame.mcu_x, frame.mcu_y);
  if (nrec != frame.mcu_x * frame.mcu_y)
    THROW_RDE(c, "Must have exactly one recepie per component");
  for (i = 0; i < nrec; ++i)
    if (!ht[i]->full)
      THROW_RDE(c, "Huffman table is not of a full decoding variety");
  if (rowsPerRestart < 1)
    THROW_RDE(c, "Number of rows per restart interval must be positives");
  if ((int64_t)frame.mcu_x * frame.dim_x > INT_MAX ||
      (int64_t)frame.mcu_y * frame.dim_y > INT_MAX)
    THROW_RDE(c, "LJpeg frame is too big");
  if ((int64_t)cpp * fw > INT_MAX)
    THROW_RDE(c, "Img frame is too big");
  if (fw < frame.mcu_x || fh < frame.mcu_y)
    THROW_RDE(c, "Tile size is smaller than a single frame MCU");
  if (fh % frame.mcu_y != 0)
    THROW_RDE(c, "Output row count is not a multiple of MCU row count");
  tileRequiredWidth = cpp * fw;
  {
    const int mcusToConsume =
        (tileRequiredWidth + frame.mcu_x - 1) / frame.mcu_x;
    if (frame.dim_x < mcusToConsume || frame.mcu_y * frame.dim_y < fh ||
        frame.mcu_x * frame.dim_x < tileRequiredWidth)
      THROW_RDE(c, "LJpeg frame (%d, %d) is smaller than expected (%d, %d)",
                frame.mcu_x * frame.dim_x, frame.mcu_y * frame.dim_y,
                tileRequiredWidth, fh);
  }
  numFullMCUs = tileRequiredWidth / frame.mcu_x;
  trailingPixels = tileRequiredWidth % frame.mcu_x;

  /* decodeN<MCU> (LJpegDecompressor.cpp:254-339) */
  {
    const int MX = frame.mcu_x, MY = frame.mcu_y, N_COMP = MX * MY;
    const int pitchE = img->pitch / 2;
    uint16_t* const imgBase = img->data + (size_t)fy * (size_t)pitchE + (size_t)cpp * fx;
    const int imgWidth = cpp * fw; /* img.width() of the cropped view */
    const int numRestartIntervals =
        ((fh / MY) + rowsPerRestart - 1) / rowsPerRestart;
    bstream inputStream;
    int ri;
    inputStream.c = c;
    inputStream.data = in;
    inputStream.size = in_size;
    inputStream.pos = 0;
    for (ri = 0; ri != numRestartIntervals; ++ri) {
      uint16_t predStorage[4];
      const uint16_t* pred = predStorage; /* Array2DRef(pred, MX, MY) */
      int predPitch = MX;
      pump bs;
      int rr;
      for (i = 0; i < N_COMP; ++i)
        predStorage[i] = init_pred[i];
      if (ri != 0) {
        /* :286-297; peekMarker = JpegMarkers.h:110-117 */
        uint8_t c0 = bs_peek_byte(&inputStream, 0);
        uint8_t c1 = bs_peek_byte(&inputStream, 1);
        if (!(c0 == 0xFF && c1 != 0 && c1 != 0xFF))
          THROW_RDE(c, "Jpeg marker not encountered");
        if (c1 < 0xD0 || c1 > 0xD7)
          THROW_RDE(c, "Not a restart marker!");
        if ((c1 - 0xD0) != ((ri - 1) % 8))
          THROW_RDE(c, "Unexpected restart marker found");
        bs_skip(&inputStream, 2);
      }
      pump_init(&bs, c, RSO_JPEG, inputStream.data + inputStream.pos,
                (int)bs_remain(&inputStream));
      for (rr = 0; rr != rowsPerRestart; ++rr) {
        const int row = MY * (rowsPerRestart * ri + rr);
        uint16_t* outStripe;
        int mcuIdx = 0, r2, c2;
        if (row == fh)
          break; /* :309-313 */
        outStripe = imgBase + (size_t)row * (size_t)pitchE;
        /* decodeRowN (:184-251) */
        for (; mcuIdx < numFullMCUs; ++mcuIdx) {
          uint16_t* outTile = outStripe + MX * mcuIdx;
          for (r2 = 0; r2 != MY; ++r2)
            for (c2 = 0; c2 != MX; ++c2) {
              int cc = MX * r2 + c2;
              int prediction = pred[r2 * predPitch + c2];
              int diff = huff_decode(ht[cc], &bs, 1);
              outTile[(size_t)r2 * (size_t)pitchE + c2] =
                  (uint16_t)(prediction + diff);
            }
          pred = outTile; /* predictor = just-decoded MCU */
          predPitch = pitchE;
        }
        if (trailingPixels != 0) {
          for (r2 = 0; r2 != MY; ++r2)
            for (c2 = 0; c2 != MX; ++c2) {
              int cc = MX * r2 + c2;
              int prediction = pred[r2 * predPitch + c2];
              int diff = huff_decode(ht[cc], &bs, 1);
              int stripeCol = MX * mcuIdx + c2;
              if (stripeCol < imgWidth)
                outStripe[(size_t)r2 * (size_t)pitchE + stripeCol] =
                    (uint16_t)(prediction + diff);
            }
          ++mcuIdx;
        }
        for (; mcuIdx < frame.dim_x; ++mcuIdx) /* decode and discard */
          for (i = 0; i != N_COMP; ++i)
            (void)huff_decode(ht[i], &bs, 1);
        /* predictor for the next line = start of this line (:326-332) */
        pred = outStripe;
        predPitch = pitchE;
      }
      bs_skip(&inputStream, (uint64_t)(uint32_t)pump_stream_position(&bs));
    }
    bs_check(&inputStream, 0);
    return inputStream.pos;
  }
}

int rso_ljpeg_decompress(rso_image* img, int fx, int fy, int fw, int fh,
                         rso_ljpeg_frame frame, const rso_huff* const* ht,
                         const uint16_t* init_pred, int nrec,
                         int rows_per_restart, const uint8_t* in,
                         uint32_t in_size, uint32_t* consumed, rso_err* e) {
  uint32_t r;
  RSO_ENTER(c, e);
  r = ljpeg_decompress_impl(&c, img, fx, fy, fw, fh, frame, ht, init_pred, nrec,
                            rows_per_restart, in, in_size);
  if (consumed)
    *consumed = r;
  return RSO_OK;
}

/* ------------------------------------------------------------------ */
/* AbstractLJpegDecoder: marker walk, SOF3/DHT/SOS/DRI parsing          */
/* ------------------------------------------------------------------ */
typedef struct {
  uint32_t componentId, dcTblNo, superH, superV;
} comp_info;

typedef struct ljpeg_dec ljpeg_dec;
struct ljpeg_dec {
  rso_ctx* c;
  bstream input;
  rso_image* img;
  /* SOFInfo (AbstractLJpegDecoder.h:64-72) */
  comp_info compInfo[4];
  uint32_t frame_w, frame_h, cps, prec;
  int sof_initialized;
  rso_huff* store[4]; /* PrefixCodeDecoderStore */
  uint8_t store_ncpl[4][16];
  uint8_t store_vals[4][17];
  int store_n[4];
  int nstore;
  rso_huff* huff[4];
  uint32_t Pt;
  uint16_t numMCUsPerRestartInterval;
  uint32_t predictorMode;
  int fixDng16Bug;
  uint32_t (*decodeScan)(ljpeg_dec*);
  /* LJpegDecoder */
  uint32_t offX, offY, w, h;
  int maxDimX, maxDimY;
  /* Cr2LJpegDecoder */
  int numSlices, sliceWidth, lastSliceWidth;
};

static void ljd_free(ljpeg_dec* d) {
  int i;
  for (i = 0; i < d->nstore; ++i)
    free(d->store[i]);
  d->nstore = 0;
}

/* getNextMarker (AbstractLJpegDecoder.cpp:282-291, JpegMarkers.h:110-135) */
static uint8_t ljd_next_marker(ljpeg_dec* d, int allowskip) {
  bstream in = d->input;
  int found = 0;
  while (bs_remain(&in) >= 2) {
    uint8_t c0 = bs_peek_byte(&in, 0), c1 = bs_peek_byte(&in, 1);
    if (c0 == 0xFF && c1 != 0 && c1 != 0xFF) {
      found = 1;
      break;
    }
    if (!allowskip)
      break;
    bs_skip(&in, 1);
  }
  if (!found)
    THROW_RDE(d->c, "(Noskip) Expected marker not found. Probably corrupt file.");
  d->input = in;
  {
    uint8_t m = bs_peek_byte(&d->input, 1);
    bs_skip(&d->input, 2);
    return m;
  }
}

static void ljd_parse_sof(ljpeg_dec* d, bstream s) { /* :127-177 */
  uint32_t i;
  d->prec = bs_get_byte(&s);
  d->frame_h = bs_get_u16be(&s);
  d->frame_w = bs_get_u16be(&s);
  d->cps = bs_get_byte(&s);
  if (d->prec < 2 || d->prec > 16)
    THROW_RDE(d->c, "Invalid precision (%u).", d->prec);
  if (d->frame_h == 0 || d->frame_w == 0)
    THROW_RDE(d->c, "Frame width or height set to zero");
  if (d->cps > 4 || d->cps < 1)
    THROW_RDE(d->c, "Only from 1 to 4 components are supported.");
  if (d->cps < (uint32_t)d->img->cpp)
    THROW_RDE(d->c, "Component count should be no less than sample count (%u vs %u).",
              d->cps, (unsigned)d->img->cpp);
  if (d->cps > (uint32_t)d->img->w)
    THROW_RDE(d->c, "Component count should be no greater than row length (%u vs %d).",
              d->cps, d->img->w);
  if (bs_remain(&s) != 3 * d->cps)
    THROW_RDE(d->c, "Header size mismatch.");
  for (i = 0; i < d->cps; i++) {
    uint32_t subs, Tq;
    d->compInfo[i].componentId = bs_get_byte(&s);
    subs = bs_get_byte(&s);
    d->compInfo[i].superV = subs & 0xf;
    d->compInfo[i].superH = subs >> 4;
    if (d->compInfo[i].superV < 1 || d->compInfo[i].superV > 4)
      THROW_RDE(d->c, "Horizontal sampling factor is invalid.");
    if (d->compInfo[i].superH < 1 || d->compInfo[i].superH > 4)
      THROW_RDE(d->c, "Horizontal sampling factor is invalid.");
    Tq = bs_get_byte(&s);
    if (Tq != 0)
      THROW_RDE(d->c, "Quantized components not supported.");
  }
  /* mRaw->metadata.subsampling defaults to (1,1) (RawImage.h:93) */
  if ((int)d->compInfo[0].superH != d->img->sub_x ||
      (int)d->compInfo[0].superV != d->img->sub_y)
    THROW_RDE(d->c, "LJpeg's subsampling does not match image's subsampling.");
  d->sof_initialized = 1;
}

static void ljd_parse_dht(ljpeg_dec* d, bstream dht) { /* :230-273 */
  while (bs_remain(&dht) > 0) {
    uint32_t b = bs_get_byte(&dht), htIndex, nCodes, i;
    uint8_t ncpl[16], vals[17];
    int idx;
    static const rso_huff zero_huff;
    rso_huff tmp = zero_huff;
    if ((b >> 4) != 0)
      THROW_RDE(d->c, "Unsupported Table class.");
    htIndex = b & 0xf;
    if (htIndex >= 4)
      THROW_RDE(d->c, "Invalid huffman table destination id.");
    if (d->huff[htIndex] != NULL)
      THROW_RDE(d->c, "Duplicate table definition");
    bs_check(&dht, 16);
    for (i = 0; i < 16; ++i)
      ncpl[i] = bs_get_byte(&dht);
    /* hc.setNCodesPerLength() validates the counts before anything else */
    nCodes = huff_validate_counts(d->c, ncpl);
    /* spec says 16 different codes is max but Hasselblad violates that -> 17 */
    if (nCodes > 17)
      THROW_RDE(d->c, "Invalid DHT table.");
    bs_check(&dht, nCodes);
    for (i = 0; i < nCodes; ++i)
      vals[i] = bs_get_byte(&dht);
    /* reuse an identical table if already in the store (:258-262) */
    idx = -1;
    for (i = 0; i < (uint32_t)d->nstore; ++i)
      if (d->store_n[i] == (int)nCodes &&
          !memcmp(d->store_ncpl[i], ncpl, 16) &&
          !memcmp(d->store_vals[i], vals, nCodes))
        idx = (int)i;
    if (idx < 0) {
      rso_huff* nh;
      huff_build(d->c, &tmp, ncpl, vals, (int)nCodes, 1, d->fixDng16Bug);
      nh = (rso_huff*)malloc(sizeof(rso_huff));
      if (!nh)
        THROW_RDE(d->c, "out of memory");
      *nh = tmp;
      idx = d->nstore++;
      d->store[idx] = nh;
      memcpy(d->store_ncpl[idx], ncpl, 16);
      memcpy(d->store_vals[idx], vals, nCodes);
      d->store_n[idx] = (int)nCodes;
    }
    d->huff[htIndex] = d->store[idx];
  }
}

static void ljd_parse_sos(ljpeg_dec* d, bstream sos) { /* :179-228 */
  uint32_t i, soscps, scanLength;
  if (bs_remain(&sos) != 1 + 2 * d->cps + 3)
    THROW_RDE(d->c, "Invalid SOS header length.");
  soscps = bs_get_byte(&sos);
  if (d->cps != soscps)
    THROW_RDE(d->c, "Component number mismatch.");
  for (i = 0; i < d->cps; i++) {
    uint32_t cs = bs_get_byte(&sos);
    uint32_t td = (uint32_t)bs_get_byte(&sos) >> 4;
    int ciIndex = -1;
    uint32_t j;
    if (td >= 4 || !d->huff[td])
      THROW_RDE(d->c, "Invalid Huffman table selection.");
    for (j = 0; j < d->cps; ++j)
      if (d->compInfo[j].componentId == cs)
        ciIndex = (int)j;
    if (ciIndex == -1)
      THROW_RDE(d->c, "Invalid Component Selector");
    d->compInfo[ciIndex].dcTblNo = td;
  }
  d->predictorMode = bs_get_byte(&sos);
  if (d->predictorMode > 8)
    THROW_RDE(d->c, "Invalid predictor mode.");
  if (bs_get_byte(&sos) != 0)
    THROW_RDE(d->c, "Se/Ah not zero.");
  d->Pt = bs_get_byte(&sos);
  if (d->Pt > 15)
    THROW_RDE(d->c, "Invalid Point transform.");
  if (d->Pt != 0)
    THROW_RDE(d->c, "Point transform not supported.");
  scanLength = d->decodeScan(d);
  bs_skip(&d->input, scanLength);
}

static void ljd_decode_soi(ljpeg_dec* d) { /* :65-125 */
  int fDRI = 0, fDHT = 0, fSOF = 0, fSOS = 0;
  uint8_t m;
  if (ljd_next_marker(d, 0) != 0xD8)
    THROW_RDE(d->c, "Image did not start with SOI. Probably not an LJPEG");
  for (; (m = ljd_next_marker(d, 1)) != 0xD9;) {
    bstream data = bs_get_stream(&d->input, bs_peek_u16be(&d->input));
    bs_skip(&data, 2);
    switch (m) {
    case 0xC4: /* DHT */
      if (fSOS)
        THROW_RDE(d->c, "Found second DHT marker after SOS");
      ljd_parse_dht(d, data);
      fDHT = 1;
      break;
    case 0xC3: /* SOF3 */
      if (fSOS)
        THROW_RDE(d->c, "Found second SOF marker after SOS");
      if (fSOF)
        THROW_RDE(d->c, "Found second SOF marker");
      ljd_parse_sof(d, data);
      fSOF = 1;
      break;
    case 0xDA: /* SOS */
      if (fSOS)
        THROW_RDE(d->c, "Found second SOS marker");
      if (!fDHT)
        THROW_RDE(d->c, "Did not find DHT marker before SOS.");
      if (!fSOF)
        THROW_RDE(d->c, "Did not find SOF marker before SOS.");
      ljd_parse_sos(d, data);
      fSOS = 1;
      break;
    case 0xDB: /* DQT */
      THROW_RDE(d->c, "Not a valid RAW file.");
    case 0xDD: /* DRI (:275-280) */
      if (fDRI)
        THROW_RDE(d->c, "Found second DRI marker");
      if (bs_remain(&data) != 2)
        THROW_RDE(d->c, "Invalid DRI header length.");
      d->numMCUsPerRestartInterval = bs_get_u16be(&data);
      fDRI = 1;
      break;
    default:
      break;
    }
  }
  if (!fSOS)
    THROW_RDE(d->c, "Did not find SOS marker.");
}

static void ljd_recipes(ljpeg_dec* d, int N_COMP, const rso_huff** hts,
                        uint16_t* initPred) {
  int i;
  /* getPrefixCodeDecoders (AbstractLJpegDecoder.h:110-126) */
  for (i = 0; i < N_COMP; ++i) {
    const uint32_t t = d->compInfo[i].dcTblNo;
    if (t >= 4)
      THROW_RDE(d->c, "Decoding table %u for comp %i does not exist (tables = %u)",
                t, i, 4u);
    hts[i] = d->huff[t];
  }
  /* getInitialPredictors (:128-136) */
  if (d->prec < (d->Pt + 1))
    THROW_RDE(d->c, "Invalid precision (%u) and point transform (%u) combination!",
              d->prec, d->Pt);
  for (i = 0; i < N_COMP; ++i)
    initPred[i] = (uint16_t)(1u << (d->prec - d->Pt - 1));
}

/* LJpegDecoder::decodeScan (LJpegDecoder.cpp:104-165) */
static uint32_t ljpegdecoder_decode_scan(ljpeg_dec* d) {
  uint32_t i;
  int N_COMP;
  const rso_huff* hts[4];
  uint16_t initPred[4];
  rso_ljpeg_frame fr;
  int64_t maxResX, maxResY;
  int mcuX, mcuY, rowsPerRestart;
  if (d->predictorMode != 1)
    THROW_RDE(d->c, "Unsupported predictor mode: %u", d->predictorMode);
  for (i = 0; i < d->cps; i++)
    if (d->compInfo[i].superH != 1 || d->compInfo[i].superV != 1)
      THROW_RDE(d->c, "Unsupported subsampling");
  N_COMP = (int)d->cps;
  ljd_recipes(d, N_COMP, hts, initPred);
  if ((int64_t)d->maxDimX * d->img->cpp > INT_MAX)
    THROW_RDE(d->c, "Maximal output tile is too large");
  maxResX = (int64_t)d->img->cpp * d->maxDimX;
  maxResY = d->maxDimY;
  if ((uint64_t)(maxResX * maxResY) !=
      (uint64_t)N_COMP * ((uint64_t)d->frame_w * d->frame_h))
    THROW_RDE(d->c, "LJpeg frame area does not match maximal tile area");
  if (maxResX % (int64_t)d->frame_w != 0 || maxResY % (int64_t)d->frame_h != 0)
    THROW_RDE(d->c, "Maximal output tile size is not a multiple of LJpeg frame size");
  mcuX = (int)(maxResX / (int64_t)d->frame_w);
  mcuY = (int)(maxResY / (int64_t)d->frame_h);
  if ((int64_t)mcuX * mcuY != N_COMP)
    THROW_RDE(d->c, "Unexpected MCU size, does not match LJpeg component count");
  fr.mcu_x = mcuX;
  fr.mcu_y = mcuY;
  fr.dim_x = (int)d->frame_w;
  fr.dim_y = (int)d->frame_h;
  if (d->numMCUsPerRestartInterval == 0)
    rowsPerRestart = fr.dim_y;
  else {
    if (d->numMCUsPerRestartInterval % fr.dim_x != 0)
      THROW_RDE(d->c, "Restart interval is not a multiple of frame row size");
    rowsPerRestart = d->numMCUsPerRestartInterval / fr.dim_x;
  }
  return ljpeg_decompress_impl(d->c, d->img, (int)d->offX, (int)d->offY,
                               (int)d->w, (int)d->h, fr, hts, initPred, N_COMP,
                               rowsPerRestart, d->input.data + d->input.pos,
                               bs_remain(&d->input));
}

static void ljd_init(ljpeg_dec* d, rso_ctx* c, const uint8_t* in,
                     uint32_t in_size, rso_image* img) {
  memset(d, 0, sizeof *d);
  d->c = c;
  d->input.c = c;
  d->input.data = in;
  d->input.size = in_size;
  d->input.pos = 0;
  d->img = img;
  /* AbstractLJpegDecoder ctor (:47-63) */
  if (!(img->w > 0 && img->h > 0))
    THROW_RDE(c, "Image has zero size");
}

static void ljpeg_decode_impl(rso_ctx* c, ljpeg_dec* d, const uint8_t* in,
                              uint32_t in_size, rso_image* img, uint32_t offX,
                              uint32_t offY, uint32_t width, uint32_t height,
                              int maxW, int maxH, int fix16) {
  ljd_init(d, c, in, in_size, img);
  /* LJpegDecoder ctor (LJpegDecoder.cpp:46-64) */
  if (img->cpp < 1 || img->cpp > 3)
    THROW_RDE(c, "Unexpected component count (%u)", (unsigned)img->cpp);
  /* LJpegDecoder::decode (:66-102) */
  if (offX >= (unsigned)img->w)
    THROW_RDE(c, "X offset outside of image");
  if (offY >= (unsigned)img->h)
    THROW_RDE(c, "Y offset outside of image");
  if (width > (unsigned)img->w)
    THROW_RDE(c, "Tile wider than image");
  if (height > (unsigned)img->h)
    THROW_RDE(c, "Tile taller than image");
  if (offX + width > (unsigned)img->w)
    THROW_RDE(c, "Tile overflows image horizontally");
  if (offY + height > (unsigned)img->h)
    THROW_RDE(c, "Tile overflows image vertically");
  if (width == 0 || height == 0)
    return;
  if (!(maxW > 0 && maxH > 0) || (unsigned)maxW < width || (unsigned)maxH < height)
    THROW_RDE(c, "Requested tile is larger than tile's maximal dimensions");
  d->offX = offX;
  d->offY = offY;
  d->w = width;
  d->h = height;
  d->maxDimX = maxW;
  d->maxDimY = maxH;
  d->fixDng16Bug = fix16;
  d->decodeScan = ljpegdecoder_decode_scan;
  ljd_decode_soi(d);
}

int rso_ljpeg_decode(const uint8_t* in, uint32_t in_size, rso_image* img,
                     uint32_t off_x, uint32_t off_y, uint32_t w, uint32_t h,
                     int max_w, int max_h, int fix_dng16, rso_err* e) {
  ljpeg_dec* d = (ljpeg_dec*)calloc(1, sizeof(ljpeg_dec));
  rso_ctx c;
  rso_err le;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (!d)
    return RSO_RDE;
  if (setjmp(c.jb)) {
    ljd_free(d);
    free(d);
    return c.e->code;
  }
  ljpeg_decode_impl(&c, d, in, in_size, img, off_x, off_y, w, h, max_w, max_h,
                    fix_dng16);
  ljd_free(d);
  free(d);
  return RSO_OK;
}

/* ------------------------------------------------------------------ */
/* AbstractDngDecompressor                                             */
/* ------------------------------------------------------------------ */
int rso_dng_decompress(const uint8_t* file, uint64_t file_size,
                       const uint64_t* tile_off, const uint32_t* tile_len,
                       int ntiles, rso_image* img, int tile_w, int tile_h,
                       int compression, int fix_ljpeg, int bps, int big_endian,
                       int nthreads, rso_err* e) {
  /* DngTilingDescription (AbstractDngDecompressor.h:37-75) */
  const int tilesX = (img->w + tile_w - 1) / tile_w;
  rso_err first;
  int nerr = 0, n;
  (void)file_size;
  first.code = RSO_OK;
  first.msg[0] = 0;
  if (e) {
    e->code = RSO_OK;
    e->msg[0] = 0;
  }
  if (compression != 1 && compression != 7) {
    if (e) {
      e->code = RSO_RDE;
      snprintf(e->msg, sizeof e->msg,
               "Too many errors encountered. Giving up. First Error:\n"
               "AbstractDngDecompressor: Unknown compression");
    }
    return RSO_RDE;
  }
  if (nthreads < 1)
    nthreads = 1;
    /* decompress(): #pragma omp parallel num_threads(cores) if(slices.size()>1)
     * + #pragma omp for schedule(static) (AbstractDngDecompressor.cpp:54-131,240-246) */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads) if (ntiles > 1)
#endif
  for (n = 0; n < ntiles; ++n) {
    /* DngSliceElement (AbstractDngDecompressor.h:77-125) */
    const int column = n % tilesX, row = n / tilesX;
    const int lastColumn = (column + 1) == tilesX;
    const int tilesY = (img->h + tile_h - 1) / tile_h;
    const int lastRow = (row + 1) == tilesY;
    const unsigned offX = (unsigned)(tile_w * column), offY = (unsigned)(tile_h * row);
    const unsigned width = !lastColumn ? (unsigned)tile_w : (unsigned)img->w - offX;
    const unsigned height = !lastRow ? (unsigned)tile_h : (unsigned)img->h - offY;
    rso_err le;
    rso_ctx c;
    ljpeg_dec* volatile d = NULL;
    c.e = &le;
    le.code = RSO_OK;
    le.msg[0] = 0;
    if (setjmp(c.jb) == 0) {
      if (compression == 7) {
        d = (ljpeg_dec*)calloc(1, sizeof(ljpeg_dec));
        ljpeg_decode_impl(&c, d, file + tile_off[n], tile_len[n], img, offX, offY,
                          width, height, tile_w, tile_h, fix_ljpeg);
      } else {
        /* decompressThread<1> (:54-110) */
        int be = big_endian;
        uint32_t inputPixelBits;
        int inputPitchBits, inputPitch;
        if (bps != 8 && bps != 16 && bps != 32 && !img->is_f32)
          be = 1; /* UINT16 images only (AbstractDngDecompressor.cpp:66-77) */
        inputPixelBits = (uint32_t)img->cpp * (uint32_t)bps;
        if ((uint32_t)tile_w > (uint32_t)INT_MAX / inputPixelBits)
          THROW_IOE(&c, "Integer overflow when calculating input pitch");
        inputPitchBits = (int)(inputPixelBits * (uint32_t)tile_w);
        if (inputPitchBits % 8 != 0)
          THROW_RDE(&c, "Bad combination of cpp (%u), bps (%u) and width (%u)",
                    (unsigned)img->cpp, (unsigned)bps, width);
        inputPitch = inputPitchBits / 8;
        if (inputPitch == 0)
          THROW_RDE(&c, "Data input pitch is too short. Can not decode!");
        unpack_impl(&c, file + tile_off[n], tile_len[n], img, (int)offX,
                    (int)offY, (int)width, (int)height, inputPitch, bps,
                    be ? RSO_MSB : RSO_LSB, img->is_f32, RSO_FORM_READ, NULL, 0);
      }
    }
    if (d) {
      ljd_free(d);
      free(d);
    }
    if (le.code != RSO_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
      {
        if (nerr == 0)
          first = le;
        ++nerr;
      }
    }
  }
  if (nerr >= 1) { /* isTooManyErrors(1) (:247-251) */
    if (e) {
      e->code = RSO_RDE;
      snprintf(e->msg, sizeof e->msg,
               "Too many errors encountered. Giving up. First Error:\n%.150s",
               first.msg);
    }
    return RSO_RDE;
  }
  return RSO_OK;
}

/* ------------------------------------------------------------------ */
/* Cr2Decompressor                                                     */
/* ------------------------------------------------------------------ */
typedef struct {
  int x, y, w, h;
} rect;

typedef struct {
  int N_COMP, X_S_F, Y_S_F, subSampled, sliceColStep, pixelsPerGroup, groupSize,
      cpp, colsPerGroup;
} cr2_dsc;

static cr2_dsc cr2_make_dsc(int n, int x, int y) {
  /* Dsc (Cr2DecompressorImpl.h:250-275) */
  cr2_dsc d;
  d.N_COMP = n;
  d.X_S_F = x;
  d.Y_S_F = y;
  d.subSampled = (x != 1 || y != 1);
  d.sliceColStep = n * x;
  d.pixelsPerGroup = x * y;
  d.groupSize = !d.subSampled ? n : 2 + d.pixelsPerGroup;
  d.cpp = !d.subSampled ? 1 : 3;
  d.colsPerGroup = !d.subSampled ? d.cpp : d.groupSize;
  return d;
}

/* Cr2OutputTileIterator (Cr2DecompressorImpl.h:109-160): enumerate all output
 * tiles of the slices in stream order. Returns count (<= cap). */
static int cr2_all_tiles(int dimX, int dimY, int frameY, int numSlices,
                         int sliceW, int lastSliceW, rect* out, int cap) {
  int n = 0, sliceId = 0, sliceRow = 0, px = 0, py = 0;
  (void)dimX;
  while (sliceId < numSlices) {
    const int w = (sliceId + 1 == numSlices) ? lastSliceW : sliceW;
    rect t;
    int outRowsRemaining = dimY - py;
    int tileRowsRemaining = frameY - sliceRow;
    t.x = px;
    t.y = py;
    t.w = w;
    t.h = outRowsRemaining < tileRowsRemaining ? outRowsRemaining
                                               : tileRowsRemaining;
    if (n < cap)
      out[n] = t;
    ++n;
    sliceRow += t.h;
    py += t.h;
    if (sliceRow == frameY) {
      ++sliceId;
      sliceRow = 0;
    }
    if (py == dimY) {
      py = 0;
      px += t.w;
    }
    if (t.h <= 0 && n > 4 * (numSlices + 1) + 65536)
      break; /* defensive: cannot happen for validated inputs */
  }
  return n;
}

/* evaluateConsecutiveTiles (Cr2DecompressorImpl.h:62-74): 0 continues, 1 new column, 2 invalid */
static int cr2_eval_tiles(rect a, rect b) {
  if (a.x == b.x && a.y + a.h == b.y && a.x + a.w == b.x + b.w)
    return 0;
  if (b.y == 0 && b.x == a.x + a.w)
    return 1;
  return 2;
}

static uint32_t cr2_decompress_impl(rso_ctx* c, rso_image* img, int n_comp,
                                    int x_s_f, int y_s_f, int frameX, int frameY,
                                    int numSlices, int sliceWidth,
                                    int lastSliceWidth,
                                    const rso_huff* const* ht,
                                    const uint16_t* init_pred, int nrec,
                                    const uint8_t* in, uint32_t in_size) {
  cr2_dsc dsc;
  int dimX, dimY, i, ntiles, nvalid;
  rect* tiles;
  /* Cr2SliceWidths ctor (Cr2Decompressor.h:66-72) */
  if (numSlices < 1)
    THROW_RDE(c, "Bad slice count: %d", numSlices);
  /* ctor (Cr2DecompressorImpl.h:279-363) */
  if (img->cpp != 1)
    THROW_RDE(c, "Unexpected cpp: %u", (unsigned)img->cpp);
  if (!((n_comp == 3 && x_s_f == 2 && y_s_f == 2) ||
        (n_comp == 3 && x_s_f == 2 && y_s_f == 1) ||
        (n_comp == 2 && x_s_f == 1 && y_s_f == 1) ||
        (n_comp == 4 && x_s_f == 1 && y_s_f == 1)))
    THROW_RDE(c, "Unknown format <%i,%i,%i>", n_comp, x_s_f, y_s_f);
  dsc = cr2_make_dsc(n_comp, x_s_f, y_s_f);
  dimX = img->w;
  dimY = img->h;
  if (!(dimX > 0 && dimY > 0) || dimX % dsc.groupSize != 0)
    THROW_RDE(c, "Unexpected image dimension multiplicity");
  dimX /= dsc.groupSize;
  if (!(frameX > 0 && frameY > 0) || frameX % dsc.X_S_F != 0 ||
      frameY % dsc.Y_S_F != 0)
    THROW_RDE(c, "Unexpected LJpeg frame dimension multiplicity");
  frameX /= dsc.X_S_F;
  frameY /= dsc.Y_S_F;
  if (img->w > 19440 || img->h > 5920)
    THROW_RDE(c, "Unexpected image dimensions found: (%d; %d)", img->w, img->h);
  for (i = 0; i < numSlices; i++) {
    const int sw = (i + 1 == numSlices) ? lastSliceWidth : sliceWidth;
    if (sw <= 0)
      THROW_RDE(c, "Bad slice width: %i", sw);
  }
  if (dsc.subSampled == (img->is_cfa != 0))
    THROW_RDE(c, "Cannot decode subsampled image to CFA data or vice versa");
  if (nrec != dsc.N_COMP)
    THROW_RDE(c, "HT/Initial predictor count does not match component count");
  for (i = 0; i < nrec; ++i)
    if (!ht[i]->full)
      THROW_RDE(c, "Huffman table is not of a full decoding variety");
  if (sliceWidth % dsc.sliceColStep != 0)
    THROW_RDE(c, "Slice width (%d) should be multiple of pixel group size (%d)",
              sliceWidth, dsc.sliceColStep);
  sliceWidth /= dsc.sliceColStep;
  if (lastSliceWidth % dsc.sliceColStep != 0)
    THROW_RDE(c, "Slice width (%d) should be multiple of pixel group size (%d)",
              lastSliceWidth, dsc.sliceColStep);
  lastSliceWidth /= dsc.sliceColStep;
  if ((uint64_t)frameX * (uint64_t)frameY < (uint64_t)dimX * (uint64_t)dimY)
    THROW_RDE(c, "Frame area smaller than the image area");

  /* NOTE: a zero slice width after division (numSlices>1, sliceWidth 0 is caught
   * above as "Bad slice width" only before division) cannot occur: width>0 and
   * divisible => >= 1. */
  ntiles = cr2_all_tiles(dimX, dimY, frameY, numSlices, sliceWidth,
                         lastSliceWidth, NULL, 0);
  tiles = (rect*)malloc(sizeof(rect) * (size_t)(ntiles > 0 ? ntiles : 1));
  if (!tiles)
    THROW_RDE(c, "out of memory");
  cr2_all_tiles(dimX, dimY, frameY, numSlices, sliceWidth, lastSliceWidth, tiles,
                ntiles);
  {
    int haveLast = 0;
    rect lastTile = {0, 0, 0, 0};
    const char* err = NULL;
    nvalid = 0;
    for (i = 0; i < ntiles; ++i) {
      rect o = tiles[i];
      if (haveLast && cr2_eval_tiles(lastTile, o) == 2) {
        err = "Invalid tiling - slice width change mid-output row?";
        break;
      }
      if (o.x + o.w <= dimX && o.y + o.h <= dimY) {
        lastTile = o;
        haveLast = 1;
        nvalid = i + 1;
        continue;
      }
      if (o.x < dimX && o.y < dimY) {
        err = "Output tile partially outside of image";
        break;
      }
      break;
    }
    if (!err && !haveLast)
      err = "No tiles are provided";
    if (!err && !(lastTile.x + lastTile.w == dimX && lastTile.y + lastTile.h == dimY))
      err = "Tiles do not cover the entire image area.";
    if (err) {
      free(tiles);
      THROW_RDE(c, "%s", err);
    }
  }

  /* decompressN_X_Y (Cr2DecompressorImpl.h:396-468) */
  {
    const int pitchE = img->pitch / 2;
    uint16_t pred[4];
    const uint16_t* predNext = img->data; /* out[0].getCrop(0, groupSize) */
    pump bs;
    int globalFrameCol = 0, t;
    jmp_buf saved;
    uint32_t pos;
    /* make sure `tiles` is released if the pump throws */
    memcpy(&saved, &c->jb, sizeof(jmp_buf));
    if (setjmp(c->jb)) {
      free(tiles);
      memcpy(&c->jb, &saved, sizeof(jmp_buf));
      longjmp(c->jb, 1);
    }
    for (i = 0; i < dsc.N_COMP; ++i)
      pred[i] = init_pred[i];
    pump_init(&bs, c, RSO_JPEG, in, (int)in_size);
    /* getOutputTiles(): tiles[0..last] where last is the first tile whose
     * bottom-right == dim (:226-238); == nvalid found above.
     * getVerticalOutputStrips(): coalesce vertically adjacent tiles (:162-205). */
    t = 0;
    {
      int lastIdx = 0;
      while (lastIdx + 1 < ntiles &&
             !(tiles[lastIdx].x + tiles[lastIdx].w == dimX &&
               tiles[lastIdx].y + tiles[lastIdx].h == dimY))
        ++lastIdx;
      nvalid = lastIdx + 1;
    }
    while (t < nvalid) {
      rect strip = tiles[t];
      int num = 1, row, col;
      while (t + num < nvalid) {
        int s = cr2_eval_tiles(strip, tiles[t + num]);
        /* `strip` has accumulated height, compare against its bottom edge */
        if (s == 1)
          break;
        strip.h += tiles[t + num].h;
        ++num;
      }
      t += num;
      for (row = strip.y; row != strip.y + strip.h; ++row) {
        uint16_t* outRow = img->data + (size_t)row * (size_t)pitchE;
        for (col = strip.x; col != strip.x + strip.w;) {
          int colFrameEnd, rem;
          if (frameX - globalFrameCol == 0) {
            int cc;
            for (cc = 0; cc < dsc.N_COMP; ++cc)
              pred[cc] = predNext[cc == 0 ? cc : dsc.groupSize - (dsc.N_COMP - cc)];
            predNext = outRow + (size_t)dsc.groupSize * (size_t)col;
            globalFrameCol = 0;
          }
          rem = frameX - globalFrameCol;
          colFrameEnd = strip.x + strip.w < col + rem ? strip.x + strip.w : col + rem;
          for (; col != colFrameEnd; ++col, ++globalFrameCol) {
            int p;
            for (p = 0; p < dsc.groupSize; ++p) {
              int cc = p < dsc.pixelsPerGroup ? 0 : p - dsc.pixelsPerGroup + 1;
              pred[cc] = (uint16_t)(pred[cc] + huff_decode(ht[cc], &bs, 1));
              outRow[dsc.groupSize * col + p] = pred[cc];
            }
          }
        }
      }
    }
    pos = (uint32_t)pump_stream_position(&bs);
    memcpy(&c->jb, &saved, sizeof(jmp_buf));
    free(tiles);
    return pos;
  }
}

int rso_cr2_decompress(rso_image* img, int n_comp, int x_s_f, int y_s_f,
                       int frame_w, int frame_h, int num_slices, int slice_w,
                       int last_slice_w, const rso_huff* const* ht,
                       const uint16_t* init_pred, int nrec, const uint8_t* in,
                       uint32_t in_size, uint32_t* consumed, rso_err* e) {
  uint32_t r;
  RSO_ENTER(c, e);
  r = cr2_decompress_impl(&c, img, n_comp, x_s_f, y_s_f, frame_w, frame_h,
                          num_slices, slice_w, last_slice_w, ht, init_pred, nrec,
                          in, in_size);
  if (consumed)
    *consumed = r;
  return RSO_OK;
}

/* Cr2LJpegDecoder::decodeScan (Cr2LJpegDecoder.cpp:58-154) */
static uint32_t cr2decoder_decode_scan(ljpeg_dec* d) {
  int isSubSampled = 0, nc, xs, ys, N_COMP;
  uint32_t i;
  const rso_huff* hts[4];
  uint16_t initPred[4];
  if (d->numMCUsPerRestartInterval != 0)
    THROW_RDE(d->c, "Non-zero restart interval not supported.");
  if (d->predictorMode != 1)
    THROW_RDE(d->c, "Unsupported predictor mode.");
  if (d->numSlices == 0 && d->sliceWidth == 0 && d->lastSliceWidth == 0) {
    const int slicesWidth = (int)(d->frame_w * d->cps);
    if (slicesWidth > d->img->w)
      THROW_RDE(d->c, "Don't know slicing pattern, and failed to guess it.");
    d->numSlices = 1;
    d->sliceWidth = 0;
    d->lastSliceWidth = (uint16_t)slicesWidth;
  }
  for (i = 0; i < d->cps; i++)
    isSubSampled = isSubSampled || d->compInfo[i].superH != 1 ||
                   d->compInfo[i].superV != 1;
  if (d->cps != 3 && d->frame_w * d->cps > 2 * d->frame_h)
    d->frame_h *= 2; /* Canon double-height quirk (:80-87) */
  if (isSubSampled) {
    int ok;
    if (d->img->is_cfa)
      THROW_RDE(d->c, "Cannot decode subsampled image to CFA data");
    if (d->cps != 3)
      THROW_RDE(d->c, "Unsupported number of subsampled components: %u", d->cps);
    ok = d->compInfo[0].superH == 2;
    ok = ok && (d->compInfo[0].superV == 1 || d->compInfo[0].superV == 2);
    for (i = 1; i < d->cps; i++)
      ok = ok && d->compInfo[i].superH == 1 && d->compInfo[i].superV == 1;
    if (!ok)
      THROW_RDE(d->c, "Unsupported subsampling");
    if (d->compInfo[0].superV == 2) {
      nc = 3;
      xs = 2;
      ys = 2;
    } else {
      d->sliceWidth = d->sliceWidth * 3 / 2;
      d->lastSliceWidth = d->lastSliceWidth * 3 / 2;
      nc = 3;
      xs = 2;
      ys = 1;
    }
  } else {
    switch (d->cps) {
    case 2:
      nc = 2;
      break;
    case 4:
      nc = 4;
      break;
    default:
      THROW_RDE(d->c, "Unsupported number of components: %u", d->cps);
    }
    xs = ys = 1;
  }
  N_COMP = nc;
  ljd_recipes(d, N_COMP, hts, initPred);
  return cr2_decompress_impl(d->c, d->img, nc, xs, ys, (int)d->frame_w,
                             (int)d->frame_h, d->numSlices, d->sliceWidth,
                             d->lastSliceWidth, hts, initPred, N_COMP,
                             d->input.data + d->input.pos, bs_remain(&d->input));
}

int rso_cr2_ljpeg_decode(const uint8_t* in, uint32_t in_size, rso_image* img,
                         int num_slices, int slice_w, int last_slice_w,
                         rso_err* e) {
  ljpeg_dec* d = (ljpeg_dec*)calloc(1, sizeof(ljpeg_dec));
  rso_ctx c;
  rso_err le;
  int i;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (!d)
    return RSO_RDE;
  if (setjmp(c.jb)) {
    ljd_free(d);
    free(d);
    return c.e->code;
  }
  ljd_init(d, &c, in, in_size, img);
  /* Cr2LJpegDecoder ctor (:42-56) */
  if (img->cpp != 1)
    THROW_RDE(&c, "Unexpected cpp: %u", (unsigned)img->cpp);
  if (!img->w || !img->h || img->w > 19440 || img->h > 5920)
    THROW_RDE(&c, "Unexpected image dimensions found: (%d; %d)", img->w, img->h);
  /* Cr2LJpegDecoder::decode (:156-165); Cr2SliceWidths ctor rejects numSlices<1
   * unless default-constructed (empty) */
  if (!(num_slices == 0 && slice_w == 0 && last_slice_w == 0) && num_slices < 1)
    THROW_RDE(&c, "Bad slice count: %d", num_slices);
  d->numSlices = num_slices;
  d->sliceWidth = slice_w;
  d->lastSliceWidth = last_slice_w;
  for (i = 0; i < num_slices; i++) {
    const int sw = (i + 1 == num_slices) ? last_slice_w : slice_w;
    if (sw <= 0)
      THROW_RDE(&c, "Bad slice width: %i", sw);
  }
  d->decodeScan = cr2decoder_decode_scan;
  ljd_decode_soi(d);
  ljd_free(d);
  free(d);
  return RSO_OK;
}

/* ================================================================== */
/* Test-input tooling: the writer half                                 */
/* ================================================================== */
typedef struct {
  uint8_t* out;
  uint64_t cap, n;
  uint64_t acc; /* bit accumulator, MSB-first */
  int nbits;
  int overflow;
  int plain; /* no FF00 stuffing (BitVacuumerMSB-style byte order) */
} bitw;

static void bw_byte(bitw* w, uint8_t b) {
  /* BitVacuumerJPEG.h:60-92: FF is followed by a stuffed 00 */
  if (w->n < w->cap)
    w->out[w->n] = b;
  else
    w->overflow = 1;
  w->n++;
  if (b == 0xFF && !w->plain) {
    if (w->n < w->cap)
      w->out[w->n] = 0x00;
    else
      w->overflow = 1;
    w->n++;
  }
}
static void bw_put(bitw* w, uint32_t bits, int count) {
  if (count == 0)
    return;
  w->acc = (w->acc << count) | (bits & (count >= 32 ? 0xFFFFFFFFu : ((1u << count) - 1u)));
  w->nbits += count;
  while (w->nbits >= 8) {
    bw_byte(w, (uint8_t)(w->acc >> (w->nbits - 8)));
    w->nbits -= 8;
  }
}
static void bw_flush_ones(bitw* w) {
  if (w->nbits > 0) {
    int pad = 8 - w->nbits;
    bw_put(w, (1u << pad) - 1u, pad);
  }
}

/* AbstractPrefixCodeEncoder::reduce (AbstractPrefixCodeEncoder.h:47-58) */
static void enc_reduce(int32_t v, uint32_t* diff, int* len) {
  if (v >= 0) {
    uint32_t d = (uint32_t)v;
    int l = 0;
    while (l < 32 && (d >> l))
      ++l;
    *diff = d;
    *len = l;
    return;
  }
  {
    uint32_t d = (uint32_t)(v - 1);
    /* numSignificantBits(d) - 1 : number of bits below the run of sign bits */
    int l = 32;
    while (l > 0 && ((d >> (l - 1)) & 1u))
      --l;
    *len = l;
    *diff = l ? (d & ((l >= 32) ? 0xFFFFFFFFu : ((1u << l) - 1u))) : 0;
  }
}

/* PrefixCodeVectorEncoder::encodeDifference (PrefixCodeVectorEncoder.h:79-90) */
static int enc_diff(bitw* w, const rso_huff* h, int32_t v) {
  uint32_t diff;
  int len, idx;
  enc_reduce(v, &diff, &len);
  if (len > 16)
    return -1;
  idx = h->idx_of_val[len];
  if (idx < 0)
    return -1;
  bw_put(w, h->code[idx], h->len[idx]);
  if (len != 16 || h->fix16)
    bw_put(w, diff, len);
  return 0;
}

int64_t rso_encode_diffs(const int32_t* diffs, uint64_t n,
                         const rso_huff* const* ht, const uint8_t* comp_of,
                         int group, uint8_t* out, uint64_t cap) {
  bitw w;
  uint64_t i;
  memset(&w, 0, sizeof w);
  w.out = out;
  w.cap = cap;
  for (i = 0; i < n; ++i)
    if (enc_diff(&w, ht[comp_of[i % (uint64_t)group]], diffs[i]))
      return -2;
  bw_flush_ones(&w);
  return w.overflow ? -1 : (int64_t)w.n;
}

static void raw_put(uint8_t* out, uint64_t cap, uint64_t* n, const void* src,
                    size_t len, int* ovf) {
  if (*n + len <= cap)
    memcpy(out + *n, src, len);
  else
    *ovf = 1;
  *n += len;
}
static void raw_u8(uint8_t* out, uint64_t cap, uint64_t* n, uint8_t v, int* o) {
  raw_put(out, cap, n, &v, 1, o);
}
static void raw_u16(uint8_t* out, uint64_t cap, uint64_t* n, uint16_t v, int* o) {
  uint8_t b[2];
  b[0] = (uint8_t)(v >> 8);
  b[1] = (uint8_t)v;
  raw_put(out, cap, n, b, 2, o);
}

static uint64_t write_headers(uint8_t* out, uint64_t cap, int frame_w,
                              int frame_h, int ncomp, int prec,
                              const uint8_t* hv /* (H<<4|V) per comp */,
                              const rso_dht* tabs, int ntab,
                              const uint8_t* tab_of_comp, int dri_mcus,
                              int* ovf) {
  uint64_t n = 0;
  int i, k;
  raw_u16(out, cap, &n, 0xFFD8, ovf);
  raw_u16(out, cap, &n, 0xFFC3, ovf);
  raw_u16(out, cap, &n, (uint16_t)(8 + 3 * ncomp), ovf);
  raw_u8(out, cap, &n, (uint8_t)prec, ovf);
  raw_u16(out, cap, &n, (uint16_t)frame_h, ovf);
  raw_u16(out, cap, &n, (uint16_t)frame_w, ovf);
  raw_u8(out, cap, &n, (uint8_t)ncomp, ovf);
  for (i = 0; i < ncomp; ++i) {
    raw_u8(out, cap, &n, (uint8_t)(i + 1), ovf);
    raw_u8(out, cap, &n, hv ? hv[i] : 0x11, ovf);
    raw_u8(out, cap, &n, 0, ovf);
  }
  for (k = 0; k < ntab; ++k) {
    raw_u16(out, cap, &n, 0xFFC4, ovf);
    raw_u16(out, cap, &n, (uint16_t)(2 + 1 + 16 + tabs[k].nvalues), ovf);
    raw_u8(out, cap, &n, (uint8_t)k, ovf);
    raw_put(out, cap, &n, tabs[k].ncpl, 16, ovf);
    raw_put(out, cap, &n, tabs[k].values, (size_t)tabs[k].nvalues, ovf);
  }
  if (dri_mcus > 0) {
    raw_u16(out, cap, &n, 0xFFDD, ovf);
    raw_u16(out, cap, &n, 4, ovf);
    raw_u16(out, cap, &n, (uint16_t)dri_mcus, ovf);
  }
  raw_u16(out, cap, &n, 0xFFDA, ovf);
  raw_u16(out, cap, &n, (uint16_t)(6 + 2 * ncomp), ovf);
  raw_u8(out, cap, &n, (uint8_t)ncomp, ovf);
  for (i = 0; i < ncomp; ++i) {
    raw_u8(out, cap, &n, (uint8_t)(i + 1), ovf);
    raw_u8(out, cap, &n, (uint8_t)(tab_of_comp[i] << 4), ovf);
  }
  raw_u8(out, cap, &n, 1, ovf); /* predictor 1 */
  raw_u8(out, cap, &n, 0, ovf);
  raw_u8(out, cap, &n, 0, ovf);
  return n;
}

int64_t rso_ljpeg_encode(const uint16_t* samples, int src_pitch_elems,
                         int frame_w, int frame_h, int mcu_x, int mcu_y,
                         int prec, const rso_dht* tabs, int ntab,
                         const uint8_t* tab_of_comp, int restart_rows,
                         int fix_dng16, uint8_t* out, uint64_t cap) {
  const int ncomp = mcu_x * mcu_y;
  rso_huff* hts[4] = {0, 0, 0, 0};
  rso_err e;
  int ovf = 0, k, r, m, i, j;
  int64_t ret = -3;
  uint64_t n;
  bitw w;
  for (k = 0; k < ntab; ++k) {
    hts[k] = rso_huff_create(tabs[k].ncpl, tabs[k].values, tabs[k].nvalues, 1,
                             fix_dng16, &e);
    if (!hts[k])
      goto done;
  }
  n = write_headers(out, cap, frame_w, frame_h, ncomp, prec, NULL, tabs, ntab,
                    tab_of_comp, restart_rows > 0 ? restart_rows * frame_w : 0,
                    &ovf);
  memset(&w, 0, sizeof w);
  w.out = out;
  w.cap = cap;
  w.n = n;
  {
    uint16_t rowStart[4], left[4];
    int interval = 0;
    for (r = 0; r < frame_h; ++r) {
      if (restart_rows > 0 && r > 0 && r % restart_rows == 0) {
        bw_flush_ones(&w);
        raw_u8(out, cap, &w.n, 0xFF, &ovf);
        raw_u8(out, cap, &w.n, (uint8_t)(0xD0 + (interval % 8)), &ovf);
        ++interval;
      }
      for (m = 0; m < frame_w; ++m) {
        for (i = 0; i < mcu_y; ++i)
          for (j = 0; j < mcu_x; ++j) {
            const int cidx = mcu_x * i + j;
            const uint16_t v =
                samples[(size_t)(r * mcu_y + i) * (size_t)src_pitch_elems +
                        (size_t)(m * mcu_x + j)];
            uint16_t pred;
            int32_t d;
            if (m == 0) {
              if (r == 0 || (restart_rows > 0 && r % restart_rows == 0))
                pred = (uint16_t)(1u << (prec - 1));
              else
                pred = rowStart[cidx];
              rowStart[cidx] = v;
            } else
              pred = left[cidx];
            left[cidx] = v;
            d = (int32_t)(int16_t)(uint16_t)(v - pred);
            /* -32768 is encoded as ssss=16 */
            if (enc_diff(&w, hts[tab_of_comp[cidx]], d)) {
              ret = -2;
              goto done;
            }
          }
      }
    }
  }
  bw_flush_ones(&w);
  raw_u16(out, cap, &w.n, 0xFFD9, &ovf);
  ret = (ovf || w.overflow) ? -1 : (int64_t)w.n;
done:
  for (k = 0; k < 4; ++k)
    rso_huff_destroy(hts[k]);
  return ret;
}

int64_t rso_cr2_encode(const rso_image* img, int n_comp, int x_s_f, int y_s_f,
                       int frame_w, int frame_h, int num_slices, int slice_w,
                       int last_slice_w, int prec, const rso_dht* tabs, int ntab,
                       const uint8_t* tab_of_comp, uint8_t* out, uint64_t cap) {
  /* Inverse of decompressN_X_Y: walk the slices in the same order and emit the
   * differences the decoder will add back.  frame_w/frame_h are the SOF3
   * values; slice widths are the CANONCR2SLICE tag values (sample columns). */
  cr2_dsc dsc = cr2_make_dsc(n_comp, x_s_f, y_s_f);
  rso_huff* hts[4] = {0, 0, 0, 0};
  rso_err e;
  int ovf = 0, k, i, ntiles, t;
  int64_t ret = -3;
  rect* tiles = NULL;
  bitw w;
  uint8_t hv[4] = {0x11, 0x11, 0x11, 0x11};
  int dimX = img->w / dsc.groupSize, dimY = img->h;
  int frameX = frame_w / dsc.X_S_F, frameY = frame_h / dsc.Y_S_F;
  int sw = slice_w, lsw = last_slice_w;
  const int pitchE = img->pitch / 2;
  uint16_t pred[4];
  const uint16_t* predNext = img->data;
  int globalFrameCol = 0;
  /* mirror Cr2LJpegDecoder's quirks so that decode(encode(x)) == x */
  if (n_comp != 3 && frame_w * n_comp > 2 * frame_h)
    frameY = (frame_h * 2) / dsc.Y_S_F;
  if (dsc.subSampled) {
    hv[0] = (uint8_t)((x_s_f << 4) | y_s_f);
    if (y_s_f == 1) {
      sw = sw * 3 / 2;
      lsw = lsw * 3 / 2;
    }
  }
  sw /= dsc.sliceColStep;
  lsw /= dsc.sliceColStep;
  for (k = 0; k < ntab; ++k) {
    hts[k] = rso_huff_create(tabs[k].ncpl, tabs[k].values, tabs[k].nvalues, 1, 0, &e);
    if (!hts[k])
      goto done;
  }
  memset(&w, 0, sizeof w);
  w.out = out;
  w.cap = cap;
  w.n = write_headers(out, cap, frame_w, frame_h, n_comp, prec, hv, tabs, ntab,
                      tab_of_comp, 0, &ovf);
  ntiles = cr2_all_tiles(dimX, dimY, frameY, num_slices, sw, lsw, NULL, 0);
  tiles = (rect*)malloc(sizeof(rect) * (size_t)(ntiles > 0 ? ntiles : 1));
  if (!tiles)
    goto done;
  cr2_all_tiles(dimX, dimY, frameY, num_slices, sw, lsw, tiles, ntiles);
  for (i = 0; i < n_comp; ++i)
    pred[i] = (uint16_t)(1u << (prec - 1));
  for (t = 0; t < ntiles; ++t) {
    rect o = tiles[t];
    int row, col;
    if (!(o.x + o.w <= dimX && o.y + o.h <= dimY))
      break;
    for (row = o.y; row != o.y + o.h; ++row) {
      const uint16_t* inRow = img->data + (size_t)row * (size_t)pitchE;
      for (col = o.x; col != o.x + o.w; ++col, ++globalFrameCol) {
        int p;
        if (globalFrameCol == frameX) {
          int cc;
          for (cc = 0; cc < n_comp; ++cc)
            pred[cc] = predNext[cc == 0 ? cc : dsc.groupSize - (n_comp - cc)];
          predNext = inRow + (size_t)dsc.groupSize * (size_t)col;
          globalFrameCol = 0;
        }
        for (p = 0; p < dsc.groupSize; ++p) {
          int cc = p < dsc.pixelsPerGroup ? 0 : p - dsc.pixelsPerGroup + 1;
          uint16_t v = inRow[dsc.groupSize * col + p];
          int32_t d = (int32_t)(int16_t)(uint16_t)(v - pred[cc]);
          pred[cc] = v;
          if (enc_diff(&w, hts[tab_of_comp[cc]], d)) {
            ret = -2;
            goto done;
          }
        }
      }
    }
    if (o.x + o.w == dimX && o.y + o.h == dimY)
      break;
  }
  bw_flush_ones(&w);
  raw_u16(out, cap, &w.n, 0xFFD9, &ovf);
  ret = (ovf || w.overflow) ? -1 : (int64_t)w.n;
done:
  free(tiles);
  for (k = 0; k < 4; ++k)
    rso_huff_destroy(hts[k]);
  return ret;
}


/* ------------------------------------------------------------------ */
/* Cr2sRawInterpolator (interpolators/Cr2sRawInterpolator.cpp)          */
/* ------------------------------------------------------------------ */
typedef struct {
  int Y, Cb, Cr;
} ycc;

static uint16_t clamp16(int x) { return (uint16_t)(x < 0 ? 0 : (x > 65535 ? 65535 : x)); }

/* YUV_TO_RGB<version> + STORE_RGB (:455-497) */
static void sraw_store(const ycc* p, uint16_t* o, const int* k, int version) {
  int r, g, b;
  if (version == 0) { /* EOS 40D */
    r = k[0] * (p->Y + p->Cr - 512);
    g = k[1] * (p->Y + ((-778 * p->Cb - (p->Cr * 2048)) >> 12) - 512);
    b = k[2] * (p->Y + (p->Cb - 512));
  } else if (version == 1) {
    r = k[0] * (p->Y + ((50 * p->Cb + 22929 * p->Cr) >> 12));
    g = k[1] * (p->Y + ((-5640 * p->Cb - 11751 * p->Cr) >> 12));
    b = k[2] * (p->Y + ((29040 * p->Cb - 101 * p->Cr) >> 12));
  } else { /* EOS 5D Mk III */
    r = k[0] * (p->Y + p->Cr);
    g = k[1] * (p->Y + ((-778 * p->Cb - (p->Cr * 2048)) >> 12));
    b = k[2] * (p->Y + p->Cb);
  }
  o[0] = clamp16(r >> 8);
  o[1] = clamp16(g >> 8);
  o[2] = clamp16(b >> 8);
}

/* full chroma sample of MCU m of input row `row`: LoadCbCr + process(hue) (:66-86) */
static ycc sraw_chroma(const uint16_t* row, int per, int ys, int m, int hue) {
  ycc c;
  c.Y = 0;
  c.Cb = (int)row[per * m + ys] - 16384 + hue;
  c.Cr = (int)row[per * m + ys + 1] - 16384 + hue;
  return c;
}

int rso_sraw_interpolate(const uint16_t* in, int in_w, int in_h, int in_pitch,
                         rso_image* out, const int* k, int hue, int version,
                         rso_err* e) {
  RSO_ENTER(c, e);
  const int sx = out->sub_x, sy = out->sub_y;
  if (sy == 1 && sx == 2) {
    /* interpolate_422 (:96-187): rows independent */
    const int numMCUs = in_w / 4;
    int row, m;
    for (row = 0; row < out->h; row++) {
      const uint16_t* ir = (const uint16_t*)((const uint8_t*)in + (size_t)row * in_pitch);
      uint16_t* o = (uint16_t*)((uint8_t*)out->data + (size_t)row * out->pitch);
      for (m = 0; m < numMCUs; ++m) {
        ycc p0 = sraw_chroma(ir, 4, 2, m, hue), p1;
        p0.Y = ir[4 * m];
        p1.Y = ir[4 * m + 1];
        if (m + 1 < numMCUs) {
          ycc n = sraw_chroma(ir, 4, 2, m + 1, hue);
          p1.Cb = (p0.Cb + n.Cb) >> 1;
          p1.Cr = (p0.Cr + n.Cr) >> 1;
        } else { /* last pixel of the line keeps the previous chroma */
          p1.Cb = p0.Cb;
          p1.Cr = p0.Cr;
        }
        sraw_store(&p0, o + 6 * m, k, version);
        sraw_store(&p1, o + 6 * m + 3, k, version);
      }
    }
    return RSO_OK;
  }
  if (sy == 2 && sx == 2) {
    /* interpolate_420 (:189-453) */
    const int numMCUs = in_w / 6;
    int row, m, i, j;
    for (row = 0; row < in_h; ++row) {
      const uint16_t* r0 = (const uint16_t*)((const uint8_t*)in + (size_t)row * in_pitch);
      const uint16_t* r1 = (const uint16_t*)((const uint8_t*)r0 + in_pitch);
      const int lastRow = row + 1 == in_h;
      for (m = 0; m < numMCUs; ++m) {
        const int lastCol = m + 1 == numMCUs;
        ycc px[2][2];
        ycc c00 = sraw_chroma(r0, 6, 4, m, hue), c01, c10, c11;
        for (i = 0; i < 2; ++i)
          for (j = 0; j < 2; ++j)
            px[i][j].Y = r0[6 * m + 2 * i + j];
        px[0][0].Cb = c00.Cb;
        px[0][0].Cr = c00.Cr;
        if (!lastCol)
          c01 = sraw_chroma(r0, 6, 4, m + 1, hue);
        if (!lastRow)
          c10 = sraw_chroma(r1, 6, 4, m, hue);
        if (!lastRow && !lastCol)
          c11 = sraw_chroma(r1, 6, 4, m + 1, hue);
        if (!lastRow && !lastCol) {
          px[0][1].Cb = (c00.Cb + c01.Cb) >> 1;
          px[0][1].Cr = (c00.Cr + c01.Cr) >> 1;
          px[1][0].Cb = (c00.Cb + c10.Cb) >> 1;
          px[1][0].Cr = (c00.Cr + c10.Cr) >> 1;
          px[1][1].Cb = (c00.Cb + c01.Cb + c10.Cb + c11.Cb) >> 2;
          px[1][1].Cr = (c00.Cr + c01.Cr + c10.Cr + c11.Cr) >> 2;
        } else if (!lastRow) { /* last MCU of the line (:349-374) */
          px[1][0].Cb = (c00.Cb + c10.Cb) >> 1;
          px[1][0].Cr = (c00.Cr + c10.Cr) >> 1;
          px[0][1].Cb = px[0][0].Cb;
          px[0][1].Cr = px[0][0].Cr;
          px[1][1].Cb = px[1][0].Cb;
          px[1][1].Cr = px[1][0].Cr;
        } else if (!lastCol) { /* last line (:405-430) */
          px[0][1].Cb = (c00.Cb + c01.Cb) >> 1;
          px[0][1].Cr = (c00.Cr + c01.Cr) >> 1;
          px[1][0].Cb = px[0][0].Cb;
          px[1][0].Cr = px[0][0].Cr;
          px[1][1].Cb = px[0][1].Cb;
          px[1][1].Cr = px[0][1].Cr;
        } else { /* last MCU of the last line (:434-452) */
          for (i = 0; i < 2; ++i)
            for (j = 0; j < 2; ++j) {
              px[i][j].Cb = c00.Cb;
              px[i][j].Cr = c00.Cr;
            }
        }
        for (i = 0; i < 2; ++i) {
          uint16_t* o = (uint16_t*)((uint8_t*)out->data + (size_t)(2 * row + i) * out->pitch);
          for (j = 0; j < 2; ++j)
            sraw_store(&px[i][j], o + 6 * m + 3 * j, k, version);
        }
      }
    }
    return RSO_OK;
  }
  THROW_RDE(&c, "Unknown subsampling: (%i; %i)", sx, sy);
}

/* ------------------------------------------------------------------ */
/* PentaxDecompressor (decompressors/PentaxDecompressor.cpp)            */
/* ------------------------------------------------------------------ */
/* ByteStream::getU16 with the stream's byte order */
static uint16_t bs_get_u16e(bstream* s, int big_endian) {
  const uint16_t v = bs_get_u16be(s);
  return big_endian ? v : (uint16_t)((v >> 8) | (v << 8));
}
static const uint8_t pentax_tree_ncpl[16] = {0, 2, 3, 1, 1, 1, 1, 1, 1, 2, 0, 0, 0, 0, 0, 0};
static const uint8_t pentax_tree_vals[13] = {3, 4, 2, 5, 1, 6, 0, 7, 8, 9, 10, 11, 12};

/* SetupPrefixCodeDecoder_Legacy / _Modern (:69-141) -> (ncpl, values) */
static int pentax_table_impl(rso_ctx* c, const uint8_t* meta, int meta_size, int meta_be,
                             uint8_t* ncpl, uint8_t* values) {
  uint32_t depth, i, j;
  uint32_t v0[16], v1[16], v2[16];
  uint8_t n17[17];
  bstream st;
  if (!meta) {
    memcpy(ncpl, pentax_tree_ncpl, 16);
    memcpy(values, pentax_tree_vals, 13);
    return 13;
  }
  st.c = c;
  st.data = meta;
  st.size = (uint32_t)meta_size;
  st.pos = 0;
  /* meta_be: byte order of the ByteStream handed in (TiffEntry::getData(): the
   * file's, PefDecoder.cpp:102-113) */
  depth = (uint32_t)bs_get_u16e(&st, meta_be) + 12;
  if (depth > 15)
    THROW_RDE(c, "Depth of huffman table is too great (%u).", depth);
  bs_skip(&st, 12);
  for (i = 0; i < depth; i++)
    v0[i] = bs_get_u16e(&st, meta_be);
  for (i = 0; i < depth; i++) {
    bs_check(&st, 1); /* ByteStream::getByte -> check(1) */
    v1[i] = st.data[st.pos++];
    if (v1[i] == 0 || v1[i] > 12)
      THROW_RDE(c, "Data corrupt: v1[%u]=%u, expected [1..12]", depth, v1[i]);
  }
  memset(n17, 0, sizeof n17);
  for (i = 0; i < depth; i++) {
    /* extractHighBits(v0, v1, effectiveBitwidth=12) = v0 >> (12 - v1) */
    v2[i] = v0[i] >> (12 - v1[i]);
    n17[v1[i]]++;
  }
  memcpy(ncpl, n17 + 1, 16);
  /* "Find smallest": repeatedly take the LAST index holding the minimum */
  for (i = 0; i < depth; i++) {
    uint32_t sm_val = 0xfffffff, sm_num = 0xff;
    for (j = 0; j < depth; j++) {
      if (v2[j] <= sm_val) {
        sm_num = j;
        sm_val = v2[j];
      }
    }
    values[i] = (uint8_t)sm_num;
    v2[sm_num] = 0xffffffff;
  }
  return (int)depth;
}

int rso_pentax_table(const uint8_t* meta, int meta_size, int meta_be, uint8_t* ncpl,
                     uint8_t* values, rso_err* e) {
  RSO_ENTER(c, e);
  return pentax_table_impl(&c, meta, meta_size, meta_be, ncpl, values);
}

int rso_pentax_decompress(rso_image* img, const uint8_t* meta, int meta_size, int meta_be,
                          const uint8_t* data, uint32_t size, rso_err* e) {
  uint8_t ncpl[16], values[16];
  int nv, row, col;
  rso_huff* volatile h = NULL;
  pump bs;
  rso_ctx c;
  rso_err le;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (setjmp(c.jb)) {
    free((void*)h);
    return c.e->code;
  }
  /* ctor (:55-67) */
  if (img->cpp != 1 || img->is_f32)
    THROW_RDE(&c, "Unexpected component count / data type");
  if (!img->w || !img->h || img->w % 2 != 0 || img->w > 8384 || img->h > 6208)
    THROW_RDE(&c, "Unexpected image dimensions found: (%d; %d)", img->w, img->h);
  nv = pentax_table_impl(&c, meta, meta_size, meta_be, ncpl, values);
  h = (rso_huff*)malloc(sizeof(rso_huff));
  if (!h)
    THROW_RDE(&c, "out of memory");
  huff_build(&c, (rso_huff*)h, ncpl, values, nv, 1, 0);
  /* decompress (:158-176) */
  pump_init(&bs, &c, RSO_MSB, data, (int)size);
  for (row = 0; row < img->h; row++) {
    uint16_t* o = (uint16_t*)((uint8_t*)img->data + (size_t)row * (size_t)img->pitch);
    int pred[2] = {0, 0};
    if (row >= 2) {
      const uint16_t* up = (const uint16_t*)((const uint8_t*)img->data +
                                             (size_t)(row - 2) * (size_t)img->pitch);
      pred[0] = up[0];
      pred[1] = up[1];
    }
    for (col = 0; col < img->w; col++) {
      int value;
      pred[col & 1] += huff_decode((const rso_huff*)h, &bs, 1);
      value = pred[col & 1];
      if (((unsigned)value >> 16) != 0) /* !isIntN(value, 16) (adt/Bit.h:83-90): 0..65535 */
        THROW_RDE(&c, "decoded value out of bounds at %d:%d", col, row);
      o[col] = (uint16_t)value;
    }
  }
  free((void*)h);
  return RSO_OK;
}

int64_t rso_encode_diffs_plain(const int32_t* diffs, uint64_t n, const rso_huff* ht,
                               uint8_t* out, uint64_t cap) {
  bitw w;
  uint64_t i;
  memset(&w, 0, sizeof w);
  w.out = out;
  w.cap = cap;
  w.plain = 1;
  for (i = 0; i < n; ++i)
    if (enc_diff(&w, ht, diffs[i]))
      return -2;
  if (w.nbits > 0)
    bw_put(&w, 0, 8 - w.nbits);
  while (w.n % 4)
    bw_put(&w, 0, 8);
  for (i = 0; i < 16; ++i)
    bw_put(&w, 0, 8);
  return w.overflow ? -1 : (int64_t)w.n;
}

/* ------------------------------------------------------------------
 * SonyArw2Decompressor (decompressors/SonyArw2Decompressor.cpp)
 * ------------------------------------------------------------------ */
int rso_sony_arw2(rso_image* img, const uint8_t* data, uint32_t size, const uint16_t* table,
                  int table_dither, rso_err* e) {
  rso_ctx c;
  rso_err le;
  int row, failed = 0;
  char first[200];
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  first[0] = 0;
  if (setjmp(c.jb))
    return c.e->code;
  /* ctor (:41-56) */
  if (img->cpp != 1 || img->is_f32)
    THROW_RDE(&c, "Unexpected component count / data type");
  if (!(img->w > 0 && img->h > 0) || img->w % 32 != 0 || img->w > 9600 || img->h > 6376)
    THROW_RDE(&c, "Unexpected image dimensions found: (%d; %d)", img->w, img->h);
  /* input_.peekStream(dim.x * dim.y) (ByteStream.h) */
  if ((uint64_t)img->w * (uint64_t)img->h > (uint64_t)size)
    THROW_IOE(&c, "Out of bounds access in ByteStream");
  /* decompressRow (:58-112): rows are independent (OpenMP in the reference) */
  for (row = 0; row < img->h; row++) {
    const uint8_t* in = data + (size_t)row * (size_t)img->w;
    uint16_t* o = (uint16_t*)((uint8_t*)img->data + (size_t)row * (size_t)img->pitch);
    /* BitStreamerLSB: bit k of the row = bit k%8 of byte k/8 */
    uint32_t random = (uint32_t)in[0] | ((uint32_t)in[1] << 8) | ((uint32_t)in[2] << 16);
    uint64_t bitpos = 0;
    int col, bad = 0;
    for (col = 0; col < img->w && !bad; col += ((col & 1) != 0) ? 31 : 1) {
      int _max, _min, _imax, _imin, sh = 0, i;
#define ARW2_BITS(n, dst)                                                                \
  do {                                                                                   \
    uint32_t v_ = 0;                                                                     \
    int k_;                                                                              \
    for (k_ = 0; k_ < (n); k_++, bitpos++)                                               \
      v_ |= (uint32_t)((in[bitpos >> 3] >> (bitpos & 7)) & 1u) << k_;                    \
    (dst) = (int)v_;                                                                     \
  } while (0)
      ARW2_BITS(11, _max);
      ARW2_BITS(11, _min);
      ARW2_BITS(4, _imax);
      ARW2_BITS(4, _imin);
      if (_imax == _imin) {
        bad = 1; /* ThrowRDE inside the row; the row loop records it (:121-127) */
        break;
      }
      while (sh < 4 && (0x80 << sh) <= (_max - _min))
        sh++;
      for (i = 0; i < 16; i++) {
        int pv;
        uint16_t value;
        if (i == _imax)
          pv = _max;
        else if (i == _imin)
          pv = _min;
        else {
          int d;
          ARW2_BITS(7, d);
          pv = (d << sh) + _min;
          if (pv > 0x7ff)
            pv = 0x7ff;
        }
        value = (uint16_t)(pv << 1);
        /* setWithLookUp (RawImage.h:335-353) */
        if (!table) {
          o[col + i * 2] = value;
        } else if (table_dither) {
          uint32_t base = table[2 * value + 0], delta = table[2 * value + 1];
          uint32_t r = random;
          uint32_t pix = base + ((delta * (r & 2047) + 1024) >> 12);
          random = 15700 * (r & 65535) + (r >> 16);
          o[col + i * 2] = (uint16_t)pix;
        } else {
          o[col + i * 2] = table[value];
        }
      }
#undef ARW2_BITS
    }
    if (bad && !failed) {
      failed = 1;
      snprintf(first, sizeof first, "ARW2 invariant failed, same pixel is both min and max");
    }
  }
  /* decompress (:135-148): isTooManyErrors(1) */
  if (failed)
    THROW_RDE(&c, "Too many errors encountered. Giving up. First Error:\n%s", first);
  return RSO_OK;
}

/* ------------------------------------------------------------------
 * NikonDecompressor (decompressors/NikonDecompressor.cpp)
 * ------------------------------------------------------------------ */
static const uint8_t nikon_tree_tab[6][2][16] = {
    /* NikonDecompressor::nikon_tree (:47-67) */
    {{0, 1, 5, 1, 1, 1, 1, 1, 1, 2, 0, 0, 0, 0, 0, 0}, {5, 4, 3, 6, 2, 7, 1, 0, 8, 9, 11, 10, 12}},
    {{0, 1, 5, 1, 1, 1, 1, 1, 1, 2, 0, 0, 0, 0, 0, 0},
     {0x39, 0x5a, 0x38, 0x27, 0x16, 5, 4, 3, 2, 1, 0, 11, 12, 12}},
    {{0, 1, 4, 2, 3, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {5, 4, 6, 3, 7, 2, 8, 1, 9, 0, 10, 11, 12}},
    {{0, 1, 4, 3, 1, 1, 1, 1, 1, 2, 0, 0, 0, 0, 0, 0},
     {5, 6, 4, 7, 8, 3, 9, 2, 1, 0, 10, 11, 12, 13, 14}},
    {{0, 1, 5, 1, 1, 1, 1, 1, 1, 1, 2, 0, 0, 0, 0, 0},
     {8, 0x5c, 0x4b, 0x3a, 0x29, 7, 6, 5, 4, 3, 2, 1, 0, 13, 14}},
    {{0, 1, 4, 2, 2, 3, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0},
     {7, 6, 8, 5, 9, 4, 10, 3, 11, 12, 2, 0, 1, 13, 14}},
};

int rso_nikon_tree(int sel, uint8_t* ncpl, uint8_t* values) {
  int n = 0, i;
  if (sel < 0 || sel > 5)
    return -1;
  for (i = 0; i < 16; i++)
    n += nikon_tree_tab[sel][0][i];
  memcpy(ncpl, nikon_tree_tab[sel][0], 16);
  memcpy(values, nikon_tree_tab[sel][1], (size_t)n);
  return n;
}

typedef struct {
  uint32_t huffSelect, split;
  int pUp[2][2];
  uint16_t curve[32770];
  uint32_t ncurve;
} nikon_setup;

/* ctor (:478-511) + createCurve (:380-441) */
static void nikon_setup_impl(rso_ctx* c, const uint8_t* meta, int meta_size, int meta_be,
                             uint32_t bitsPS, int img_w, int img_h, nikon_setup* o) {
  bstream md;
  uint32_t v0, v1, csize, step = 0, i, n;
  md.c = c;
  md.data = meta;
  md.size = (uint32_t)meta_size;
  md.pos = 0;
  if (!(img_w > 0 && img_h > 0) || img_w % 2 != 0 || img_w > 8288 || img_h > 5520)
    THROW_RDE(c, "Unexpected image dimensions found: (%d; %d)", img_w, img_h);
  if (bitsPS != 12 && bitsPS != 14)
    THROW_RDE(c, "Invalid bpp found: %u", bitsPS);
  v0 = bs_get_byte(&md);
  v1 = bs_get_byte(&md);
  if (v0 == 73 || v1 == 88)
    bs_skip(&md, 2110);
  o->huffSelect = 0;
  o->split = 0;
  if (v0 == 70)
    o->huffSelect = 2;
  if (bitsPS == 14)
    o->huffSelect += 3;
  o->pUp[0][0] = bs_get_u16e(&md, meta_be);
  o->pUp[1][0] = bs_get_u16e(&md, meta_be);
  o->pUp[0][1] = bs_get_u16e(&md, meta_be);
  o->pUp[1][1] = bs_get_u16e(&md, meta_be);
  /* createCurve */
  if (v0 == 68 && v1 == 64)
    bitsPS -= 2; /* Nikon Z7 12/14 bit compressed hack */
  n = ((1u << bitsPS) & 0x7fffu) + 1u;
  for (i = 0; i < n; i++)
    o->curve[i] = (uint16_t)i;
  csize = bs_get_u16e(&md, meta_be);
  if (csize > 1)
    step = n / (csize - 1);
  if (v0 == 68 && (v1 == 32 || v1 == 64) && step > 0) {
    if ((csize - 1) * step != n - 1)
      THROW_RDE(c, "Bad curve segment count (%u)", csize);
    for (i = 0; i < csize; i++)
      o->curve[i * step] = bs_get_u16e(&md, meta_be);
    for (i = 0; i < n - 1; i++) {
      const uint32_t b_scale = i % step, a_pos = i - b_scale, b_pos = a_pos + step;
      const uint32_t a_scale = step - b_scale;
      o->curve[i] = (uint16_t)((a_scale * o->curve[a_pos] + b_scale * o->curve[b_pos]) / step);
    }
    /* metadata.setPosition(562) (ByteStream.h: check against the size) */
    if (562 > md.size)
      THROW_IOE(c, "Out of bounds access in ByteStream");
    md.pos = 562;
    o->split = bs_get_u16e(&md, meta_be);
  } else if (v0 != 70) {
    if (csize == 0 || csize > 0x4001)
      THROW_RDE(c, "Don't know how to compute curve! csize = %u", csize);
    n = csize + 1;
    for (i = 0; i < csize; i++)
      o->curve[i] = bs_get_u16e(&md, meta_be);
  }
  o->ncurve = n - 1; /* "and drop the last value" */
  if (o->split >= (uint32_t)img_h)
    o->split = 0;
}

int rso_nikon_setup(const uint8_t* meta, int meta_size, int meta_be, int bitsPS, int img_w,
                    int img_h, uint16_t* curve, int* ncurve, int* pup, int* huff_select,
                    int* split, rso_err* e) {
  rso_ctx c;
  rso_err le;
  nikon_setup* volatile s = NULL;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (setjmp(c.jb)) {
    free((void*)s);
    return c.e->code;
  }
  s = (nikon_setup*)malloc(sizeof(nikon_setup));
  if (!s)
    THROW_RDE(&c, "out of memory");
  nikon_setup_impl(&c, meta, meta_size, meta_be, (uint32_t)bitsPS, img_w, img_h, (nikon_setup*)s);
  if (curve)
    memcpy(curve, ((nikon_setup*)s)->curve, ((nikon_setup*)s)->ncurve * sizeof(uint16_t));
  if (ncurve)
    *ncurve = (int)((nikon_setup*)s)->ncurve;
  if (pup) {
    pup[0] = ((nikon_setup*)s)->pUp[0][0];
    pup[1] = ((nikon_setup*)s)->pUp[0][1];
    pup[2] = ((nikon_setup*)s)->pUp[1][0];
    pup[3] = ((nikon_setup*)s)->pUp[1][1];
  }
  if (huff_select)
    *huff_select = (int)((nikon_setup*)s)->huffSelect;
  if (split)
    *split = (int)((nikon_setup*)s)->split;
  free((void*)s);
  return RSO_OK;
}

/* NikonLASDecompressor (NikonDecompressor.cpp:80-378): the "lossy after split" decoder.
 * Its own table builder (T.81 C.1/C.2/F.15 without HuffmanCode's validation), an 8-bit and a
 * 14-bit lookup, and a difference format where a code value packs (len | shl << 4). */
typedef struct {
  uint32_t bits[17];
  uint32_t huffval[256];
  uint16_t mincode[17];
  int maxcode[18];
  int16_t valptr[17];
  uint32_t numbits[256];
  int bigTable[1 << 14];
} nikon_las;

static void las_create(rso_ctx* c, nikon_las* t, const uint8_t* ncpl, const uint8_t* values,
                       int nvalues) {
  int p, i, l, lastp, si, size, value, ll, ul;
  int8_t huffsize[257];
  uint16_t huffcode[257], code;
  memset(t, 0, sizeof *t);
  for (i = 0; i < 16; i++)
    t->bits[i + 1] = ncpl[i];
  for (i = 0; i < nvalues; i++)
    t->huffval[i] = values[i];
  /* createPrefixCodeDecoder (:112-210) */
  p = 0;
  for (l = 1; l <= 16; l++)
    for (i = 1; i <= (int)t->bits[l]; i++) {
      huffsize[p++] = (int8_t)l;
      if (p > 256)
        THROW_RDE(c, "LJpegDecoder::createPrefixCodeDecoder: Code length too long. Corrupt data.");
    }
  huffsize[p] = 0;
  lastp = p;
  code = 0;
  si = huffsize[0];
  p = 0;
  while (huffsize[p]) {
    while ((int)huffsize[p] == si) {
      huffcode[p++] = code;
      code++;
    }
    code = (uint16_t)(code << 1);
    si++;
    if (p > 256)
      THROW_RDE(c, "createPrefixCodeDecoder: Code length too long. Corrupt data.");
  }
  t->mincode[0] = 0;
  t->maxcode[0] = 0;
  p = 0;
  for (l = 1; l <= 16; l++) {
    if (t->bits[l]) {
      t->valptr[l] = (int16_t)p;
      t->mincode[l] = huffcode[p];
      p += (int)t->bits[l];
      t->maxcode[l] = huffcode[p - 1];
    } else {
      t->valptr[l] = 0xff;
      t->maxcode[l] = -1;
    }
    if (p > 256)
      THROW_RDE(c, "createPrefixCodeDecoder: Code length too long. Corrupt data.");
  }
  t->maxcode[17] = 0xFFFFF;
  for (p = 0; p < lastp; p++) {
    size = huffsize[p];
    if (size <= 8) {
      value = (int)t->huffval[p];
      code = huffcode[p];
      ll = code << (8 - size);
      ul = size < 8 ? (ll | (int)(0xffffffffu >> (24 + size))) : ll; /* bitMask[24 + size] */
      if (ul > 256 || ll > ul)
        THROW_RDE(c, "createPrefixCodeDecoder: Code length too long. Corrupt data.");
      for (i = ll; i <= ul; i++)
        t->numbits[i] = (uint32_t)(size | (value << 4));
    }
  }
  /* createBigTable (:224-283), bits = 14 */
  for (i = 0; i < (1 << 14); i++) {
    const uint16_t input = (uint16_t)(i << 2);
    int cd = input >> 8, rv, x;
    uint32_t val = t->numbits[cd];
    uint32_t ln = val & 15;
    if (ln) {
      rv = (int)(val >> 4);
    } else {
      ln = 8;
      while (cd > t->maxcode[ln]) {
        /* extractHighBits(input, ln, effectiveBitwidth = 15) & 1 */
        const int temp = (input >> (15 - ln)) & 1;
        cd = (cd << 1) | temp;
        ln++;
      }
      if (ln > 16 || t->valptr[ln] == 0xff) {
        t->bigTable[i] = 0xff;
        continue;
      }
      rv = (int)t->huffval[t->valptr[ln] + (cd - t->mincode[ln])];
    }
    if (rv == 16) {
      t->bigTable[i] = (-(32768 << 8)) | (int)ln;
      continue;
    }
    if (rv + (int)ln > 14) {
      t->bigTable[i] = 0xff;
      continue;
    }
    if (rv) {
      /* extractHighBits(input, ln + rv) of the 16-bit input */
      x = (input >> (16 - (ln + (uint32_t)rv))) & ((1 << rv) - 1);
      if ((x & (1 << (rv - 1))) == 0)
        x -= (1 << rv) - 1;
      t->bigTable[i] = (int)(((unsigned)x << 8) | (ln + (uint32_t)rv));
    } else {
      t->bigTable[i] = (int)ln;
    }
  }
}

/* NikonLASDecompressor::decodeDifference (:315-377) */
static int las_decode(const nikon_las* t, pump* bs) {
  int rv, l, code;
  unsigned val;
  uint32_t len, shl, nb;
  int diff;
  pump_fill(bs, 32);
  code = (int)pump_peek_nofill(bs, 14);
  val = (unsigned)t->bigTable[code];
  if ((val & 0xff) != 0xff) {
    pump_skip_nofill(bs, (int)(val & 0xff));
    return (int)val >> 8;
  }
  rv = 0;
  code = (int)pump_peek_nofill(bs, 8);
  val = t->numbits[code];
  l = (int)(val & 15);
  if (l) {
    pump_skip_nofill(bs, l);
    rv = (int)val >> 4;
  } else {
    pump_skip_nofill(bs, 8);
    l = 8;
    while (code > t->maxcode[l]) {
      const int temp = (int)pump_get_nofill(bs, 1);
      code = (code << 1) | temp;
      l++;
    }
    if (l > 16)
      THROW_RDE(bs->c, "Corrupt JPEG data: bad Huffman code:%d\n", l);
    rv = (int)t->huffval[t->valptr[l] + (code - t->mincode[l])];
  }
  if (rv == 16)
    return -32768;
  len = (uint32_t)rv & 15;
  shl = (uint32_t)rv >> 4;
  nb = len - shl;
  /* (bits.getBits(0) is not defined in the reference; real tables never get here with nb == 0) */
  diff = (int)(((((nb ? pump_get_bits(bs, (int)nb) : 0u) << 1) + 1) << shl) >> 1);
  if ((diff & (1 << (len - 1))) == 0)
    diff -= (1 << len) - !shl;
  return diff;
}

int rso_nikon_decompress(rso_image* img, const uint8_t* meta, int meta_size, int meta_be,
                         int bitsPS, const uint8_t* data, uint32_t size, int uncorrected,
                         rso_err* e) {
  rso_ctx c;
  rso_err le;
  nikon_setup* volatile s = NULL;
  rso_huff* volatile h = NULL;
  uint16_t* volatile tab = NULL; /* dithered TableLookUp storage: {base, delta} per value */
  nikon_las* volatile las = NULL;
  pump bs;
  uint8_t ncpl[16], values[16];
  int nv, row, col, split;
  uint32_t random, i;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (setjmp(c.jb)) {
    free((void*)s);
    free((void*)h);
    free((void*)tab);
    free((void*)las);
    return c.e->code;
  }
  if (img->cpp != 1 || img->is_f32)
    THROW_RDE(&c, "Unexpected component count / data type");
  s = (nikon_setup*)malloc(sizeof(nikon_setup));
  h = (rso_huff*)malloc(sizeof(rso_huff));
  if (!s || !h)
    THROW_RDE(&c, "out of memory");
  nikon_setup_impl(&c, meta, meta_size, meta_be, (uint32_t)bitsPS, img->w, img->h,
                   (nikon_setup*)s);
  /* RawImageCurveGuard + TableLookUp::setTable(curve, dither = true) (TableLookUp.cpp:49-84) */
  if (!uncorrected) {
    const uint16_t* cv = ((nikon_setup*)s)->curve;
    const int nf = (int)((nikon_setup*)s)->ncurve;
    int k;
    tab = (uint16_t*)malloc(2u * 65536u * sizeof(uint16_t));
    if (!tab)
      THROW_RDE(&c, "out of memory");
    for (k = 0; k < nf; k++) {
      int center = cv[k], lower = k > 0 ? cv[k - 1] : center,
          upper = k < nf - 1 ? cv[k + 1] : center, v;
      if (lower > center)
        lower = center;
      if (upper < center)
        upper = center;
      v = center - ((upper - lower + 2) / 4);
      ((uint16_t*)tab)[2 * k] = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
      ((uint16_t*)tab)[2 * k + 1] = (uint16_t)(upper - lower);
    }
    for (k = nf; k < 65536; k++) {
      ((uint16_t*)tab)[2 * k] = cv[nf - 1];
      ((uint16_t*)tab)[2 * k + 1] = 0;
    }
  }
  nv = rso_nikon_tree((int)((nikon_setup*)s)->huffSelect, ncpl, values);
  huff_build(&c, (rso_huff*)h, ncpl, values, nv, 1, 0);
  split = (int)((nikon_setup*)s)->split;
  if (split) {
    uint8_t n2[16], v2[16];
    const int nv2 = rso_nikon_tree((int)((nikon_setup*)s)->huffSelect + 1, n2, v2);
    las = (nikon_las*)malloc(sizeof(nikon_las));
    if (!las || nv2 < 0)
      THROW_RDE(&c, "out of memory");
    las_create(&c, (nikon_las*)las, n2, v2, nv2);
  }
  /* decompress (:540-560, :513-538) */
  pump_init(&bs, &c, RSO_MSB, data, (int)size);
  pump_fill(&bs, 24);
  random = pump_peek_nofill(&bs, 24);
  for (row = 0; row < img->h; row++) {
    uint16_t* o = (uint16_t*)((uint8_t*)img->data + (size_t)row * (size_t)img->pitch);
    int (*pUp)[2] = ((nikon_setup*)s)->pUp;
    int pred[2];
    pred[0] = pUp[row & 1][0];
    pred[1] = pUp[row & 1][1];
    for (col = 0; col < img->w; col++) {
      int v;
      uint16_t value;
      /* rows below the split: nikon_tree[huffSelect] through PrefixCodeDecoder<>; from the
       * split on: nikon_tree[huffSelect + 1] through NikonLASDecompressor (:549-556) */
      pred[col & 1] += (split == 0 || row < split) ? huff_decode((const rso_huff*)h, &bs, 1)
                                                   : las_decode((const nikon_las*)las, &bs);
      if (col < 2)
        pUp[row & 1][col & 1] = pred[col & 1];
      v = pred[col & 1]; /* clampBits(v, 15) */
      value = (uint16_t)(v < 0 ? 0 : (v > 32767 ? 32767 : v));
      if (!tab) {
        o[col] = value;
      } else {
        uint32_t base = ((uint16_t*)tab)[2 * value], delta = ((uint16_t*)tab)[2 * value + 1];
        uint32_t r = random;
        o[col] = (uint16_t)(base + ((delta * (r & 2047) + 1024) >> 12));
        random = 15700 * (r & 65535) + (r >> 16);
      }
    }
  }
  (void)i;
  free((void*)s);
  free((void*)h);
  free((void*)tab);
  free((void*)las);
  return RSO_OK;
}

/* ------------------------------------------------------------------
 * PanasonicV5 / V6 / V7 decompressors
 * ------------------------------------------------------------------ */
/* n bits at bit offset `off` of a 16-byte little-endian block (BitStreamerLSB order) */
static uint32_t pana_bits(const uint8_t* blk, uint32_t off, uint32_t n) {
  uint32_t v = 0, k;
  for (k = 0; k < n; k++, off++)
    v |= (uint32_t)((blk[off >> 3] >> (off & 7)) & 1u) << k;
  return v;
}

int rso_panasonic(int version, rso_image* img, const uint8_t* data, uint32_t size, int bps,
                  rso_err* e) {
  rso_ctx c;
  rso_err le;
  const uint64_t area = (uint64_t)(img->w > 0 ? img->w : 0) * (uint64_t)(img->h > 0 ? img->h : 0);
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (setjmp(c.jb))
    return c.e->code;
  if (img->cpp != 1 || img->is_f32)
    THROW_RDE(&c, "Unexpected component count / data type");
  if (version == 5) {
    /* ctor (PanasonicV5Decompressor.cpp:58-116) */
    const uint32_t BlockSize = 0x4000, split = 0x1FF8;
    uint32_t ppp;
    uint64_t numPackets, numBlocks, b;
    if (bps != 12 && bps != 14)
      THROW_RDE(&c, "Unsupported bps: %u", (unsigned)bps);
    ppp = 128u / (uint32_t)bps; /* pixels per 16-byte packet (padding bits left over) */
    if (!(img->w > 0 && img->h > 0) || (uint32_t)img->w % ppp != 0)
      THROW_RDE(&c, "Unexpected image dimensions found: (%i; %i)", img->w, img->h);
    numPackets = area / ppp;
    numBlocks = (numPackets + 1023) / 1024;
    if ((uint64_t)size / BlockSize < numBlocks)
      THROW_RDE(&c, "Insufficient count of input blocks for a given image");
    /* processBlock (:206-232) with ProxyStream's section swap (:147-186) */
    for (b = 0; b < numBlocks; b++) {
      const uint8_t* blk = data + b * BlockSize;
      uint32_t pk;
      for (pk = 0; pk < 1024; pk++) {
        uint8_t pkt[16];
        uint64_t idx = (b * 1024 + pk) * ppp;
        uint32_t i, j;
        if (idx >= area)
          break;
        for (j = 0; j < 16; j++)
          pkt[j] = blk[(pk * 16 + j + split) & (BlockSize - 1)];
        for (i = 0; i < ppp; i++, idx++) {
          uint16_t* o = (uint16_t*)((uint8_t*)img->data + (size_t)(idx / (uint64_t)img->w) * (size_t)img->pitch);
          o[idx % (uint64_t)img->w] = (uint16_t)pana_bits(pkt, i * (uint32_t)bps, (uint32_t)bps);
        }
      }
    }
    return RSO_OK;
  }
  if (version == 7) {
    /* PanasonicV7Decompressor.cpp:40-106: 9 pixels of 14 bits per 16-byte block */
    uint64_t numBlocks, b;
    if (!(img->w > 0 && img->h > 0) || img->w % 9 != 0)
      THROW_RDE(&c, "Unexpected image dimensions found: (%i; %i)", img->w, img->h);
    numBlocks = area / 9;
    if ((uint64_t)size / 16 < numBlocks)
      THROW_RDE(&c, "Insufficient count of input blocks for a given image");
    for (b = 0; b < numBlocks; b++) {
      const uint64_t idx = b * 9;
      uint16_t* o = (uint16_t*)((uint8_t*)img->data + (size_t)(idx / (uint64_t)img->w) * (size_t)img->pitch) +
                    idx % (uint64_t)img->w;
      uint32_t i;
      for (i = 0; i < 9; i++)
        o[i] = (uint16_t)pana_bits(data + b * 16, i * 14, 14);
    }
    return RSO_OK;
  }
  if (version == 6) {
    /* PanasonicV6Decompressor.cpp:146-239 */
    const int is14 = bps == 14;
    const uint32_t ppb = is14 ? 11u : 14u, PixelbaseZero = is14 ? 0x200u : 0x80u,
                   PixelbaseCompare = is14 ? 0x2000u : 0x800u, SpixCompare = is14 ? 0xffffu : 0x3fffu,
                   PixelMask = is14 ? 0x3fffu : 0xfffu;
    uint64_t numBlocks, b;
    if (bps != 12 && bps != 14)
      THROW_RDE(&c, "Unsupported bps: %u", (unsigned)bps);
    if (!(img->w > 0 && img->h > 0) || (uint32_t)img->w % ppb != 0)
      THROW_RDE(&c, "Unexpected image dimensions found: (%i; %i)", img->w, img->h);
    numBlocks = area / ppb;
    if ((uint64_t)size / 16 < numBlocks)
      THROW_RDE(&c, "Insufficient count of input blocks for a given image");
    for (b = 0; b < numBlocks; b++) {
      const uint8_t* blk = data + b * 16;
      const uint64_t idx = b * ppb;
      uint16_t* o = (uint16_t*)((uint8_t*)img->data + (size_t)(idx / (uint64_t)img->w) * (size_t)img->pitch) +
                    idx % (uint64_t)img->w;
      uint16_t buf[18];
      uint32_t off = 0, k, cur = 0, pix;
      uint32_t oddeven[2] = {0, 0}, nonzero[2] = {0, 0}, pmul = 0, pixel_base = 0;
      /* pana_cs6_page_decoder<B>::fillBuffer (:88-142): the buffer is filled from its END */
      if (is14) {
        off = 4;
        for (k = 14; k-- > 2;) {
          const uint32_t n = (k % 4 == 2) ? 2u : 10u;
          buf[k] = (uint16_t)pana_bits(blk, off, n);
          off += n;
        }
        buf[1] = (uint16_t)pana_bits(blk, off, 14);
        buf[0] = (uint16_t)pana_bits(blk, off + 14, 14);
      } else {
        for (k = 18; k-- > 2;) {
          const uint32_t n = (k % 4 == 2) ? 2u : 8u;
          buf[k] = (uint16_t)pana_bits(blk, off, n);
          off += n;
        }
        buf[1] = (uint16_t)pana_bits(blk, off, 12);
        buf[0] = (uint16_t)pana_bits(blk, off + 12, 12);
      }
      /* decompressBlock (:178-221) */
      for (pix = 0; pix < ppb; pix++) {
        uint16_t epixel;
        uint32_t spix;
        if (pix % 3 == 2) {
          uint16_t base = buf[cur++];
          if (base == 3)
            base = 4;
          pixel_base = PixelbaseZero << base;
          pmul = 1u << base;
        }
        epixel = buf[cur++];
        if (oddeven[pix % 2]) {
          epixel = (uint16_t)(epixel * pmul);
          if (pixel_base < PixelbaseCompare && nonzero[pix % 2] > pixel_base)
            epixel = (uint16_t)(epixel + (nonzero[pix % 2] - pixel_base));
          nonzero[pix % 2] = epixel;
        } else {
          oddeven[pix % 2] = epixel;
          if (epixel)
            nonzero[pix % 2] = epixel;
          else
            epixel = (uint16_t)nonzero[pix % 2];
        }
        spix = (uint32_t)((int)epixel - 0xf);
        if (spix <= SpixCompare)
          o[pix] = (uint16_t)(spix & SpixCompare);
        else {
          epixel = (uint16_t)((int)(epixel + 0x7ffffff1) >> 0x1f);
          o[pix] = (uint16_t)(epixel & PixelMask);
        }
      }
    }
    return RSO_OK;
  }
  THROW_RDE(&c, "unknown Panasonic version %d", version);
  return RSO_OK;
}

/* ------------------------------------------------------------------
 * PhaseOneDecompressor (decompressors/PhaseOneDecompressor.cpp)
 * ------------------------------------------------------------------ */
typedef struct {
  jmp_buf* outer;
} p1_dummy;

int rso_phaseone(rso_image* img, const uint8_t* file, uint64_t file_size, const uint64_t* off,
                 const uint32_t* len, const int32_t* rown, int nstrips, rso_err* e) {
  static const int length[10] = {8, 7, 6, 9, 11, 10, 5, 12, 14, 13};
  rso_ctx c;
  rso_err le;
  int* volatile order = NULL;
  volatile int failed = 0;
  char first[320];
  int k;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  first[0] = 0;
  if (setjmp(c.jb)) {
    free((void*)order);
    return c.e->code;
  }
  /* ctor (:42-59) */
  if (img->is_f32)
    THROW_RDE(&c, "Unexpected data type");
  if (img->cpp != 1)
    THROW_RDE(&c, "Unexpected cpp: %u", (unsigned)img->cpp);
  if (!(img->w > 0 && img->h > 0) || img->w % 2 != 0 || img->w > 11976 || img->h > 8854)
    THROW_RDE(&c, "Unexpected image dimensions found: (%d; %d)", img->w, img->h);
  /* prepareStrips (:61-83): exactly one strip per row */
  if (nstrips != img->h)
    THROW_RDE(&c, "Height (%d) vs strip count %zu mismatch", img->h, (size_t)nstrips);
  order = (int*)malloc(sizeof(int) * (size_t)img->h);
  if (!order)
    THROW_RDE(&c, "out of memory");
  for (k = 0; k < img->h; k++)
    ((int*)order)[k] = -1;
  for (k = 0; k < nstrips; k++) {
    if (rown[k] < 0 || rown[k] >= img->h || ((int*)order)[rown[k]] != -1)
      THROW_RDE(&c, "Strips validation issue.");
    if (off[k] + len[k] > file_size)
      THROW_IOE(&c, "Out of bounds access in ByteStream");
    ((int*)order)[rown[k]] = k;
  }
  /* decompressStrip (:85-135), one row at a time; a throwing row is recorded (:137-150) */
  for (k = 0; k < img->h; k++) {
    const int s = ((int*)order)[k];
    rso_ctx rc;
    rso_err re;
    rc.e = &re;
    re.code = RSO_OK;
    re.msg[0] = 0;
    if (setjmp(rc.jb)) {
      if (!failed) {
        failed = 1;
        snprintf(first, sizeof first, "%s", re.msg);
      }
      continue;
    }
    {
      pump bs;
      int32_t pred[2] = {0, 0};
      int ln[2] = {0, 0}, col;
      uint16_t* o = (uint16_t*)((uint8_t*)img->data + (size_t)k * (size_t)img->pitch);
      pump_init(&bs, &rc, RSO_MSB32, file + off[s], (int)len[s]);
      for (col = 0; col < img->w; col++) {
        int i;
        pump_fill(&bs, 32);
        if ((unsigned)col >= ((unsigned)img->w & ~7u)) {
          ln[0] = ln[1] = 14;
        } else if ((col & 7) == 0) {
          int t;
          for (t = 0; t < 2; t++) {
            int j = 0;
            for (; j < 5; j++) {
              if (pump_get_nofill(&bs, 1) != 0) {
                if (col == 0)
                  THROW_RDE(&rc, "Can not initialize lengths. Data is corrupt.");
                break;
              }
            }
            if (j > 0)
              ln[t] = length[2 * (j - 1) + (int)pump_get_nofill(&bs, 1)];
          }
        }
        i = ln[col & 1];
        if (i == 14) {
          pred[col & 1] = (int32_t)pump_get_nofill(&bs, 16);
          o[col] = (uint16_t)pred[col & 1];
        } else {
          pred[col & 1] += (int32_t)pump_get_nofill(&bs, i) + 1 - (1 << (i - 1));
          o[col] = (uint16_t)pred[col & 1];
        }
      }
    }
  }
  free((void*)order);
  order = NULL;
  if (failed)
    THROW_RDE(&c, "Too many errors encountered. Giving up. First Error:\n%s", first);
  return RSO_OK;
}

/* ------------------------------------------------------------------
 * HasselbladDecompressor (decompressors/HasselbladDecompressor.cpp)
 * ------------------------------------------------------------------ */
/* HasselbladDecompressor::getBits (:60-70) */
static int hassel_bits(pump* bs, int len) {
  int diff;
  if (!len)
    return 0;
  diff = (int)pump_get_bits(bs, len);
  diff = rso_huff_extend((uint32_t)diff, (uint32_t)len);
  if (diff == 65535)
    return -32768;
  return diff;
}

int rso_hasselblad_decompress(rso_image* img, const rso_huff* ht, uint16_t init_pred,
                              const uint8_t* in, uint32_t in_size, uint32_t* consumed,
                              rso_err* e) {
  rso_ctx c;
  rso_err le;
  pump bs;
  int row, col, k;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (setjmp(c.jb))
    return c.e->code;
  /* ctor (:39-58) */
  if (img->is_f32)
    THROW_RDE(&c, "Unexpected data type");
  if (img->cpp != 1)
    THROW_RDE(&c, "Unexpected cpp: %u", (unsigned)img->cpp);
  if (!(img->w > 0 && img->h > 0) || img->w % 2 != 0 || img->w > 12000 || img->h > 8842)
    THROW_RDE(&c, "Unexpected image dimensions found: (%d; %d)", img->w, img->h);
  if (ht->full)
    THROW_RDE(&c, "Huffman table is of a full decoding variety");
  /* ht.verifyCodeValuesAsDiffLengths() (codes/AbstractPrefixCode.h): every value <= 16 */
  for (k = 0; k < ht->nsym; k++)
    if (ht->val[k] > 16)
      THROW_RDE(&c, "Corrupt Huffman code: difference length %u longer than 16", (unsigned)ht->val[k]);
  /* decompress (:72-100) */
  pump_init(&bs, &c, RSO_MSB32, in, (int)in_size);
  for (row = 0; row < img->h; row++) {
    uint16_t* o = (uint16_t*)((uint8_t*)img->data + (size_t)row * (size_t)img->pitch);
    int p1 = init_pred, p2 = init_pred;
    for (col = 0; col < img->w; col += 2) {
      const int len1 = huff_decode(ht, &bs, 0);
      const int len2 = huff_decode(ht, &bs, 0);
      p1 += hassel_bits(&bs, len1);
      p2 += hassel_bits(&bs, len2);
      o[col] = (uint16_t)p1;
      o[col + 1] = (uint16_t)p2;
    }
  }
  if (consumed)
    *consumed = (uint32_t)pump_stream_position(&bs);
  return RSO_OK;
}

/* ------------------------------------------------------------------
 * PanasonicV4Decompressor (decompressors/PanasonicV4Decompressor.cpp)
 * ------------------------------------------------------------------ */
typedef struct {
  const uint8_t* buf; /* the block with its sections swapped back + one zero byte */
  int vbits;
} pana4_stream;

/* ProxyStream::getBits (:164-168) */
static uint32_t pana4_bits(pana4_stream* s, int nbits) {
  int byte;
  s->vbits = (s->vbits - nbits) & 0x1ffff;
  byte = (s->vbits >> 3) ^ 0x3ff0;
  return (uint32_t)((s->buf[byte] | s->buf[byte + 1] << 8) >> (s->vbits & 7)) & ~(uint32_t)(-(1 << nbits));
}

int rso_panasonic_v4(rso_image* img, const uint8_t* data, uint32_t size, int zero_is_not_bad,
                     uint32_t section_split_offset, uint32_t* zero_pos, uint32_t cap,
                     uint32_t* nzero, rso_err* e) {
  enum { BlockSize = 0x4000, PixelsPerPacket = 14, BytesPerPacket = 16 };
  rso_ctx c;
  rso_err le;
  uint8_t* volatile buf = NULL;
  uint64_t area, bytesTotal, bufSize, blocksTotal, b, pixel = 0;
  uint32_t nz = 0;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (setjmp(c.jb)) {
    free((void*)buf);
    return c.e->code;
  }
  /* ctor (:49-90) */
  if (img->cpp != 1 || img->is_f32)
    THROW_RDE(&c, "Unexpected component count / data type");
  if (!(img->w > 0 && img->h > 0) || img->w % PixelsPerPacket != 0)
    THROW_RDE(&c, "Unexpected image dimensions found: (%i; %i)", img->w, img->h);
  if (BlockSize < section_split_offset)
    THROW_RDE(&c, "Bad section_split_offset: %u, less than BlockSize (%u)", section_split_offset,
              (unsigned)BlockSize);
  area = (uint64_t)img->w * (uint64_t)img->h;
  bytesTotal = area / PixelsPerPacket * BytesPerPacket;
  bufSize = section_split_offset == 0 ? bytesTotal
                                      : (bytesTotal + BlockSize - 1) / BlockSize * BlockSize;
  if (bufSize > 0xFFFFFFFFull)
    THROW_RDE(&c, "Raw dimensions require input buffer larger than supported");
  if (bufSize > size) /* input_.peekStream(bufSize) */
    THROW_IOE(&c, "Out of bounds access in ByteStream");
  buf = (uint8_t*)malloc(BlockSize + 2);
  if (!buf)
    THROW_RDE(&c, "out of memory");
  /* chopInputIntoBlocks (:92-127) + processBlock (:216-236) */
  blocksTotal = (bufSize + BlockSize - 1) / BlockSize;
  for (b = 0; b < blocksTotal; b++) {
    const uint64_t off = b * BlockSize;
    const uint32_t blockSize = (uint32_t)(bufSize - off < BlockSize ? bufSize - off : BlockSize);
    const uint32_t packets = blockSize / BytesPerPacket;
    pana4_stream st;
    uint32_t k;
    /* ProxyStream::parseBlock (:136-159): second section first */
    {
      const uint32_t first = section_split_offset < blockSize ? section_split_offset : blockSize;
      memcpy((uint8_t*)buf, data + off + first, blockSize - first);
      memcpy((uint8_t*)buf + (blockSize - first), data + off, first);
      ((uint8_t*)buf)[blockSize] = 0;
    }
    st.buf = (const uint8_t*)buf;
    st.vbits = 0;
    for (k = 0; k < packets && pixel < area; k++) {
      /* processPixelPacket (:171-214) */
      int sh = 0, u = 0, p;
      int pred[2] = {0, 0}, nonz[2] = {0, 0};
      for (p = 0; p < PixelsPerPacket; p++, pixel++) {
        const int cc = p & 1;
        const uint32_t row = (uint32_t)(pixel / (uint64_t)img->w), col = (uint32_t)(pixel % (uint64_t)img->w);
        if (u == 2) {
          /* extractHighBits(4U, bits.getBits(2), effectiveBitwidth = 3) = 4 >> (3 - n) */
          sh = 4 >> (3 - (int)pana4_bits(&st, 2));
          u = -1;
        }
        if (nonz[cc]) {
          const int j = (int)pana4_bits(&st, 8);
          if (j) {
            pred[cc] -= 0x80 << sh;
            if (pred[cc] < 0 || sh == 4)
              pred[cc] &= ~(-(1 << sh));
            pred[cc] += j << sh;
          }
        } else {
          nonz[cc] = (int)pana4_bits(&st, 8);
          if (nonz[cc] || p > 11)
            pred[cc] = nonz[cc] << 4 | (int)pana4_bits(&st, 4);
        }
        ((uint16_t*)((uint8_t*)img->data + (size_t)row * (size_t)img->pitch))[col] = (uint16_t)pred[cc];
        if (!zero_is_not_bad && pred[cc] == 0) {
          if (zero_pos && nz < cap)
            zero_pos[nz] = (row << 16) | col;
          nz++;
        }
        u++;
      }
    }
  }
  if (nzero)
    *nzero = nz;
  free((void*)buf);
  return RSO_OK;
}

/* ------------------------------------------------------------------
 * RawImageDataU16::scaleValues (common/RawImageDataU16.cpp)
 * ------------------------------------------------------------------ */
int rso_scale_uses_sse2(const int* black_sep, int white) {
  /* scaleValues (:185-202): Cpuid::SSE2() && app_scale < 63 */
  const int depth_values = white - black_sep[0];
  const float app_scale = 65535.0F / (float)depth_values;
  return app_scale < 63;
}

int rso_scale_values(rso_image* img, int off_x, int off_y, int crop_w, int crop_h,
                     const int* black_sep, int white, int dither, int sse2, rso_err* e) {
  rso_ctx c;
  rso_err le;
  const int depth_values = white - black_sep[0];
  const float app_scale = 65535.0F / (float)depth_values;
  const int full_scale_fp = (int)(app_scale * 4.0F);    /* 30.2 fixed point */
  const int half_scale_fp = (int)(app_scale * 4095.0F); /* 18.14 fixed point */
  int y;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (setjmp(c.jb))
    return c.e->code;
  if (img->is_f32)
    THROW_RDE(&c, "Unexpected data type");
  if (off_x < 0 || off_y < 0 || crop_w <= 0 || crop_h <= 0 || off_x + crop_w > img->w ||
      off_y + crop_h > img->h)
    THROW_RDE(&c, "bad crop");
  if (!sse2) {
    /* scaleValues_plain (:343-399) */
    const int gw = crop_w * img->cpp;
    int mul[4], sub[4], i;
    for (i = 0; i < 4; i++) {
      int v = i;
      if (off_x & 1)
        v ^= 1;
      if (off_y & 1)
        v ^= 2;
      mul[i] = (int)(16384.0F * 65535.0F / (float)(white - black_sep[v]));
      sub[i] = black_sep[v];
    }
    for (y = 0; y < crop_h; y++) {
      uint16_t* row = (uint16_t*)((uint8_t*)img->data + (size_t)(off_y + y) * (size_t)img->pitch) +
                      off_x * img->cpp;
      int v = crop_w + y * 36969, x;
      for (x = 0; x < gw; x++) {
        int rnd = 0, val;
        if (dither) {
          v = 18000 * (v & 65535) + (v >> 16);
          rnd = half_scale_fp - (full_scale_fp * (v & 2047));
        }
        val = ((row[x] - sub[(2 * (y & 1)) + (x & 1)]) * mul[(2 * (y & 1)) + (x & 1)] + 8192 + rnd) >> 14;
        row[x] = (uint16_t)(val < 0 ? 0 : (val > 65535 ? 65535 : val)); /* clampBits(.., 16) */
      }
    }
    return RSO_OK;
  }
  /* scaleValues_SSE2 (:204-341), lane by lane */
  {
    uint32_t sub_even, mul_even, sub_odd, mul_odd;
    const int xend = img->w / 8 * 8; /* x < roundDown(uncropped_dim.x, 8): PIXELS, also when cpp > 1 (:309) */
    /* 10 bit fraction; the pair (column parity 0, 1) packed into one 32-bit lane */
    mul_even = (uint32_t)(int)(1024.0F * 65535.0F / (float)(white - black_sep[off_x & 1]));
    mul_even |= (uint32_t)(int)(1024.0F * 65535.0F / (float)(white - black_sep[(off_x + 1) & 1])) << 16;
    sub_even = (uint32_t)black_sep[off_x & 1] | ((uint32_t)black_sep[(off_x + 1) & 1] << 16);
    mul_odd = (uint32_t)(int)(1024.0F * 65535.0F / (float)(white - black_sep[2 + (off_x & 1)]));
    mul_odd |= (uint32_t)(int)(1024.0F * 65535.0F / (float)(white - black_sep[2 + ((off_x + 1) & 1)])) << 16;
    sub_odd = (uint32_t)black_sep[2 + (off_x & 1)] | ((uint32_t)black_sep[2 + ((off_x + 1) & 1)] << 16);
    for (y = 0; y < crop_h; y++) {
      uint16_t* row = (uint16_t*)((uint8_t*)img->data + (size_t)(off_y + y) * (size_t)img->pitch);
      const uint32_t subv = ((y + off_y) & 1) == 0 ? sub_even : sub_odd;
      const uint32_t mulv = ((y + off_y) & 1) == 0 ? mul_even : mul_odd;
      uint16_t rnd[8];
      int x, k;
      if (dither) {
        /* _mm_set_epi32(e3, e2, e1, e0): 32-bit lane 0 = e0 */
        const uint32_t l0 = (uint32_t)(crop_w * 1234 + y * 23464), l1 = (uint32_t)(crop_w * 4272 + y * 12123),
                       l2 = (uint32_t)(crop_w * 2342 + y * 34311), l3 = (uint32_t)(crop_w * 1676 + y * 18000);
        rnd[0] = (uint16_t)l0; rnd[1] = (uint16_t)(l0 >> 16);
        rnd[2] = (uint16_t)l1; rnd[3] = (uint16_t)(l1 >> 16);
        rnd[4] = (uint16_t)l2; rnd[5] = (uint16_t)(l2 >> 16);
        rnd[6] = (uint16_t)l3; rnd[7] = (uint16_t)(l3 >> 16);
      } else {
        memset(rnd, 0, sizeof rnd);
      }
      for (x = 0; x < xend; x += 8) {
        for (k = 0; k < 8; k++) {
          /* sserandom = mulhi_epi16(r, m) ^ mullo_epi16(r, m), m = 0x1d32 / 0x4d9f (0 without dither) */
          const int16_t m = dither ? (int16_t)((k & 1) ? 0x4d9f : 0x1d32) : 0;
          const int32_t prod = (int32_t)(int16_t)rnd[k] * (int32_t)m;
          rnd[k] = (uint16_t)((uint32_t)prod >> 16) ^ (uint16_t)prod;
        }
        for (k = 0; k < 8; k++) {
          const uint16_t sub16 = (uint16_t)((k & 1) ? subv >> 16 : subv);
          const uint16_t mul16 = (uint16_t)((k & 1) ? mulv >> 16 : mulv);
          const uint16_t pix = row[x + k] > sub16 ? (uint16_t)(row[x + k] - sub16) : 0; /* subs_epu16 */
          const uint32_t p32 = (uint32_t)pix * (uint32_t)mul16;      /* mulhi:mullo */
          const uint16_t r16 = (uint16_t)((rnd[k] & 0x00ff) * (uint16_t)full_scale_fp); /* mullo_epi16 */
          const uint32_t radd = (uint32_t)(half_scale_fp >> 4) - (uint32_t)r16;
          int32_t v = (int32_t)(p32 + 512u + radd); /* epi32 adds wrap */
          v >>= 10;                                  /* srai */
          v = (int32_t)((uint32_t)v - 32768u);       /* sub_epi32 */
          if (v < -32768)                            /* packs_epi32: signed saturation */
            v = -32768;
          if (v > 32767)
            v = 32767;
          row[x + k] = (uint16_t)((uint16_t)(int16_t)v ^ 0x8000u);
        }
      }
    }
  }
  return RSO_OK;
}

int rso_scale_black_white(rso_image* img, int off_x, int off_y, int crop_w, int crop_h,
                          int black_level, int* black_sep, int has_sep, int* white, int has_white,
                          const rso_black_area* areas, int n_areas, int dither, int force_sse2,
                          rso_err* e) {
  rso_ctx c;
  rso_err le;
  uint16_t* hist = NULL;
  const int skip = 250;
  int rc, i;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (setjmp(c.jb)) {
    free(hist);
    return c.e->code;
  }
  if (img->is_f32)
    THROW_RDE(&c, "Unexpected data type");
  if (off_x < 0 || off_y < 0 || crop_w < 0 || crop_h < 0 || off_x + crop_w > img->w ||
      off_y + crop_h > img->h)
    THROW_RDE(&c, "bad crop");
  /* scaleBlackWhite (:147-171): estimate from the crop minus a 250 pixel border */
  if ((n_areas == 0 && !has_sep && black_level < 0) || !has_white) {
    const int gw = (crop_w - skip) * img->cpp;
    int b = 65536, m = 0, row, col;
    for (row = skip; row < crop_h - skip; row++) {
      const uint16_t* p = (const uint16_t*)((const uint8_t*)img->data + (size_t)(off_y + row) * (size_t)img->pitch) +
                          off_x * img->cpp;
      for (col = skip; col < gw; col++) {
        const int pixel = p[skip + col];
        b = pixel < b ? pixel : b;
        m = pixel > m ? pixel : m;
      }
    }
    if (black_level < 0)
      black_level = b;
    if (!has_white) {
      *white = m;
      has_white = 1;
    }
  }
  /* (:173-177) nothing to do */
  if ((n_areas == 0 && black_level == 0 && *white == 65535 && !has_sep) || crop_w <= 0 || crop_h <= 0)
    return RSO_OK;
  if (!has_sep) {
    /* calculateBlackAreas (:60-145) */
    int totalpixels = 0;
    hist = (uint16_t*)calloc(4 * 65536, sizeof(uint16_t));
    if (!hist)
      THROW_RDE(&c, "out of memory");
    for (i = 0; i < n_areas; i++) {
      const uint32_t offset = areas[i].offset;
      const uint32_t size = areas[i].size - (areas[i].size & 1);
      if (!areas[i].is_vertical) {
        uint32_t y;
        int x;
        if ((int)offset + (int)size > img->h)
          THROW_RDE(&c, "Offset + size is larger than height of image");
        for (y = offset; y < offset + size; y++) {
          const uint16_t* p = (const uint16_t*)((const uint8_t*)img->data + (size_t)y * (size_t)img->pitch);
          for (x = off_x; x < crop_w + off_x; x++)
            hist[(size_t)((2 * (y & 1)) + (x & 1)) * 65536 + p[off_x]]++; /* one sampled column (:87-91) */
        }
        totalpixels += (int)(size * (uint32_t)crop_w);
      } else {
        int y;
        uint32_t x;
        if ((int)offset + (int)size > img->w)
          THROW_RDE(&c, "Offset + size is larger than width of image");
        for (y = off_y; y < crop_h + off_y; y++) {
          const uint16_t* p = (const uint16_t*)((const uint8_t*)img->data + (size_t)y * (size_t)img->pitch);
          for (x = offset; x < size + offset; x++)
            hist[(size_t)((2 * (y & 1)) + (x & 1)) * 65536 + p[offset]]++;
        }
        totalpixels += (int)(size * (uint32_t)crop_h);
      }
    }
    if (!totalpixels) {
      for (i = 0; i < 4; i++)
        black_sep[i] = black_level;
    } else {
      totalpixels /= 4 * 2;
      for (i = 0; i < 4; i++) {
        const uint16_t* h = hist + (size_t)i * 65536;
        int acc = h[0], v = 0;
        while (acc <= totalpixels && v < 65535) {
          v++;
          acc += h[v];
        }
        black_sep[i] = v;
      }
      if (!img->is_cfa) {
        int total = 0;
        for (i = 0; i < 4; i++)
          total += black_sep[i];
        for (i = 0; i < 4; i++)
          black_sep[i] = (total + 2) >> 2;
      }
    }
    free(hist);
    hist = NULL;
  }
  rc = rso_scale_values(img, off_x, off_y, crop_w, crop_h, black_sep, *white, dither,
                        force_sse2 < 0 ? rso_scale_uses_sse2(black_sep, *white) : force_sse2, c.e);
  return rc;
}

/* ------------------------------------------------------------------
 * DngOpcodes (common/DngOpcodes.cpp)
 * ------------------------------------------------------------------ */
typedef struct {
  int code;
  uint32_t value;            /* FixBadPixelsConstant */
  uint32_t* bad;             /* FixBadPixelsList */
  uint32_t nbad;
  int top, left, w, h;       /* ROI (pos, dim) */
  uint32_t firstPlane, planes, rowPitch, colPitch;
  uint16_t* lookup;          /* MapTable / MapPolynomial: 65536 entries */
  float* deltaF;             /* Delta / Scale per row / column */
  uint32_t ndelta;
} dng_op;

typedef struct {
  dng_op* ops;
  uint32_t nops;
  uint32_t* list; /* mBadPixelPositions */
  uint32_t nlist, caplist;
} dng_state;

static void dng_free(dng_state* st) {
  uint32_t i;
  for (i = 0; i < st->nops; i++) {
    free(st->ops[i].bad);
    free(st->ops[i].lookup);
    free(st->ops[i].deltaF);
  }
  free(st->ops);
  free(st->list);
  st->ops = NULL;
  st->list = NULL;
}

static uint32_t bs_get_u32be(bstream* s) {
  uint32_t v;
  bs_check(s, 4);
  v = ld_be32(s->data + s->pos);
  s->pos += 4;
  return v;
}
/* ByteStream::check(nmemb, size) / skipBytes(nmemb, size) (io/ByteStream.h:71-75, :130-132) */
static uint32_t bs_check2(const bstream* s, uint32_t nmemb, uint32_t size) {
  if (size && nmemb > 0xFFFFFFFFu / size)
    THROW_IOE(s->c, "Integer overflow when calculating stream length");
  bs_check(s, (uint64_t)nmemb * size);
  return nmemb * size;
}
static uint64_t dng_round_up_div(uint64_t a, uint64_t b) { return a ? 1 + (a - 1) / b : 0; }

/* ROIOpcode ctor (:193-226): rectangle inside {0, 0, dim} (inclusive), bottomRight >= topLeft */
static void dng_read_roi(rso_ctx* c, bstream* bs, int dim_x, int dim_y, dng_op* op) {
  const uint32_t top = bs_get_u32be(bs), left = bs_get_u32be(bs), bottom = bs_get_u32be(bs),
                 right = bs_get_u32be(bs);
  const int tx = (int)left, ty = (int)top, bx = (int)right, by = (int)bottom;
  const int ok = tx >= 0 && ty >= 0 && tx <= dim_x && ty <= dim_y && bx >= 0 && by >= 0 &&
                 bx <= dim_x && by <= dim_y && bx >= tx && by >= ty;
  if (!ok)
    THROW_RDE(c, "Rectangle (%d, %d, %d, %d) not inside image (%d, %d, %d, %d).", tx, ty, bx, by, 0,
              0, dim_x, dim_y);
  op->left = tx;
  op->top = ty;
  op->w = bx - tx;
  op->h = by - ty;
}

/* PixelOpcode ctor (:353-381) */
static void dng_read_pixel_op(rso_ctx* c, bstream* bs, int cpp, int dim_x, int dim_y, dng_op* op) {
  dng_read_roi(c, bs, dim_x, dim_y, op);
  op->firstPlane = bs_get_u32be(bs);
  op->planes = bs_get_u32be(bs);
  if (op->planes == 0 || op->firstPlane > (uint32_t)cpp || op->planes > (uint32_t)cpp ||
      op->firstPlane + op->planes > (uint32_t)cpp)
    THROW_RDE(c, "Bad plane params (first %u, num %u), got planes = %u", op->firstPlane, op->planes,
              (unsigned)cpp);
  op->rowPitch = bs_get_u32be(bs);
  op->colPitch = bs_get_u32be(bs);
  if (op->rowPitch < 1 || op->rowPitch > (uint32_t)op->h || op->colPitch < 1 ||
      op->colPitch > (uint32_t)op->w)
    THROW_RDE(c, "Invalid pitch");
}

static void dng_list_reserve(rso_ctx* c, dng_state* st, uint64_t extra) {
  if ((uint64_t)st->nlist + extra > st->caplist) {
    uint64_t ncap = ((uint64_t)st->nlist + extra) * 2 + 16;
    uint32_t* n = (uint32_t*)realloc(st->list, ncap * sizeof(uint32_t));
    if (!n || ncap > 0xFFFFFFFFull)
      THROW_RDE(c, "out of memory");
    st->list = n;
    st->caplist = (uint32_t)ncap;
  }
}

int rso_dng_opcodes(rso_image* img, int* crop, const uint8_t* data, uint32_t size, uint32_t* bad,
                    uint32_t bad_cap, uint32_t* nbad, int* applied, rso_err* e) {
  rso_ctx c;
  rso_err le;
  dng_state* volatile stp;
  dng_state st;
  bstream bs;
  uint32_t opcode_count, i;
  int int_x, int_y, int_w, int_h; /* integrated_subimg */
  const int cpp = img->cpp;
  volatile int done = 0;
  memset(&st, 0, sizeof st);
  stp = &st;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (applied)
    *applied = 0;
  if (nbad)
    *nbad = 0;
  if (setjmp(c.jb)) {
    dng_state* sp = (dng_state*)stp;
    if (applied)
      *applied = done;
    /* an apply-time error leaves the earlier opcodes' list entries in place */
    if (nbad)
      *nbad = sp->nlist;
    for (i = 0; i < sp->nlist && i < bad_cap && bad; i++)
      bad[i] = sp->list[i];
    dng_free(sp);
    return c.e->code;
  }
  /* DngOpcodes::DngOpcodes (:666-726), big endian */
  bs.c = &c;
  bs.data = data;
  bs.size = size;
  bs.pos = 0;
  opcode_count = bs_get_u32be(&bs);
  {
    const uint32_t orig = bs.pos;
    for (i = 0; i < opcode_count; i++) {
      uint32_t opcode_size;
      bs_skip(&bs, 4);
      bs_skip(&bs, 4);
      bs_skip(&bs, 4);
      opcode_size = bs_get_u32be(&bs);
      bs_skip(&bs, opcode_size);
    }
    bs.pos = orig;
  }
  st.ops = (dng_op*)calloc(opcode_count ? opcode_count : 1, sizeof(dng_op));
  if (!st.ops)
    THROW_RDE(&c, "out of memory");
  int_x = crop[0];
  int_y = crop[1];
  int_w = crop[2];
  int_h = crop[3];
  for (i = 0; i < opcode_count; i++) {
    const uint32_t code = bs_get_u32be(&bs);
    uint32_t flags, opcode_size;
    bstream ob;
    dng_op* op = &st.ops[st.nops];
    bs_skip(&bs, 4); /* version */
    flags = bs_get_u32be(&bs);
    opcode_size = bs_get_u32be(&bs);
    ob = bs_get_stream(&bs, opcode_size);
    memset(op, 0, sizeof *op);
    op->code = (int)code;
    switch (code) {
    case 1:
    case 2:
    case 3:
    case 9:
      /* known, not implemented (:751-757, :776): an error unless flagged optional */
      if (!(flags & 1)) {
        static const char* const names[] = {"", "WarpRectilinear", "WarpFisheye", "FixVignetteRadial",
                                            "", "", "", "", "", "GainMap"};
        THROW_RDE(&c, "Unsupported Opcode: %u (%s)", code, names[code]);
      }
      op = NULL;
      break;
    case 4: /* FixBadPixelsConstant (:149-185) */
      st.nops++;
      op->value = bs_get_u32be(&ob);
      (void)bs_get_u32be(&ob); /* Bayer phase */
      break;
    case 5: { /* FixBadPixelsList (:263-325): coordinates of the uncropped image */
      uint32_t npts, nrect, k;
      uint64_t n = 0, capn;
      st.nops++;
      (void)bs_get_u32be(&ob); /* phase */
      npts = bs_get_u32be(&ob);
      nrect = bs_get_u32be(&ob);
      {
        const uint32_t orig = ob.pos;
        bs_check(&ob, 0);
        ob.pos += bs_check2(&ob, npts, 8);
        ob.pos += bs_check2(&ob, nrect, 16);
        ob.pos = orig;
        bs_check(&ob, 0);
      }
      capn = (uint64_t)npts + 16;
      op->bad = (uint32_t*)malloc(capn * sizeof(uint32_t));
      if (!op->bad)
        THROW_RDE(&c, "out of memory");
      for (k = 0; k < npts; k++) {
        const uint32_t y = bs_get_u32be(&ob), x = bs_get_u32be(&ob);
        const int px = (int)x, py = (int)y;
        if (!(px >= 0 && py >= 0 && px < img->w && py < img->h))
          THROW_RDE(&c, "Bad point not inside image.");
        op->bad[n++] = y << 16 | x;
        op->nbad = (uint32_t)n;
      }
      for (k = 0; k < nrect; k++) {
        dng_op r;
        int y, x;
        memset(&r, 0, sizeof r);
        dng_read_roi(&c, &ob, img->w, img->h, &r);
        if (n + (uint64_t)r.w * r.h > capn) {
          uint32_t* nb;
          capn = (n + (uint64_t)r.w * r.h) * 2;
          nb = (uint32_t*)realloc(op->bad, capn * sizeof(uint32_t));
          if (!nb)
            THROW_RDE(&c, "out of memory");
          op->bad = nb;
        }
        for (y = 0; y < r.h; y++)
          for (x = 0; x < r.w; x++)
            op->bad[n++] = (uint32_t)(r.top + y) << 16 | (uint32_t)(r.left + x);
        op->nbad = (uint32_t)n;
      }
      break;
    }
    case 6: /* TrimBounds (:332-346) */
      st.nops++;
      dng_read_roi(&c, &ob, int_w, int_h, op);
      int_x += op->left;
      int_y += op->top;
      int_w = op->w;
      int_h = op->h;
      break;
    case 7: { /* MapTable (:446-466) */
      uint32_t count, k;
      st.nops++;
      dng_read_pixel_op(&c, &ob, cpp, int_w, int_h, op);
      count = bs_get_u32be(&ob);
      if (count == 0 || count > 65536)
        THROW_RDE(&c, "Invalid size of lookup table");
      op->lookup = (uint16_t*)calloc(65536, sizeof(uint16_t));
      if (!op->lookup)
        THROW_RDE(&c, "out of memory");
      for (k = 0; k < count; k++)
        op->lookup[k] = bs_get_u16be(&ob);
      for (k = count; k < 65536; k++)
        op->lookup[k] = op->lookup[count - 1];
      break;
    }
    case 8: { /* MapPolynomial (:473-505) */
      double poly[9];
      uint64_t psize;
      uint32_t k, j;
      st.nops++;
      dng_read_pixel_op(&c, &ob, cpp, int_w, int_h, op);
      psize = (uint64_t)bs_get_u32be(&ob) + 1;
      bs_check(&ob, (uint32_t)(8 * psize)); /* implicit_cast<size_type>(8UL * polynomial_size) */
      if (psize > 9)
        THROW_RDE(&c, "A polynomial with more than 8 degrees not allowed");
      for (k = 0; k < psize; k++) {
        uint64_t bits;
        bs_check(&ob, 8);
        bits = ((uint64_t)ld_be32(ob.data + ob.pos) << 32) | ld_be32(ob.data + ob.pos + 4);
        ob.pos += 8;
        memcpy(&poly[k], &bits, 8);
      }
      op->lookup = (uint16_t*)calloc(65536, sizeof(uint16_t));
      if (!op->lookup)
        THROW_RDE(&c, "out of memory");
      for (k = 0; k < 65536; k++) {
        double val = poly[0], t;
        for (j = 1; j < psize; j++)
          val += poly[j] * pow((double)k / 65536.0, (double)j);
        t = val * 65535.5;
        t = t < 0.0 ? 0.0 : (t > 65535.0 ? 65535.0 : t); /* std::clamp<double>(.., 0, 65535) */
        op->lookup[k] = (uint16_t)t;
      }
      break;
    }
    case 10:
    case 11:
    case 12:
    case 13: { /* DeltaRowOrCol (:535-589): 10 / 12 select the row index, 11 / 13 the column */
      uint32_t count, k;
      uint64_t expected;
      st.nops++;
      dng_read_pixel_op(&c, &ob, cpp, int_w, int_h, op);
      count = bs_get_u32be(&ob);
      (void)bs_check2(&ob, count, 4);
      expected = (code == 10 || code == 12) ? dng_round_up_div((uint64_t)op->h, op->rowPitch)
                                            : dng_round_up_div((uint64_t)op->w, op->colPitch);
      if (expected != count)
        THROW_RDE(&c, "Got unexpected number of elements (%llu), expected %u.",
                  (unsigned long long)expected, count);
      op->deltaF = (float*)malloc(((size_t)count + 1) * sizeof(float));
      if (!op->deltaF)
        THROW_RDE(&c, "out of memory");
      for (k = 0; k < count; k++) {
        const uint32_t bits = bs_get_u32be(&ob);
        float f;
        memcpy(&f, &bits, 4);
        if (!isfinite(f))
          THROW_RDE(&c, "Got bad float %f.", (double)f);
        op->deltaF[k] = f;
      }
      op->ndelta = count;
      break;
    }
    default:
      THROW_RDE(&c, "Unknown unhandled Opcode: %u", code);
    }
    if (ob.size - ob.pos != 0)
      THROW_RDE(&c, "Inconsistent length of opcode");
    (void)op;
  }

  /* applyOpCodes (:730-735): setup() then apply(), in order */
  for (i = 0; i < st.nops; i++) {
    const dng_op* op = &st.ops[i];
    const int ox = crop[0], oy = crop[1], dw = crop[2], dh = crop[3];
    switch (op->code) {
    case 4: {
      int row, col;
      if (img->is_f32)
        THROW_RDE(&c, "Only 16 bit images supported");
      if (cpp > 1)
        THROW_RDE(&c, "Only 1 component images supported");
      for (row = 0; row < dh; row++) {
        const uint16_t* p = (const uint16_t*)((const uint8_t*)img->data + (size_t)(oy + row) * (size_t)img->pitch) + ox;
        for (col = 0; col < dw; col++)
          if (p[col] == op->value) {
            dng_list_reserve(&c, &st, 1);
            st.list[st.nlist++] = ((uint32_t)ox | ((uint32_t)oy << 16)) + ((uint32_t)row << 16 | (uint32_t)col);
          }
      }
      break;
    }
    case 5: /* inserted at the BEGINNING of the list (:319-321) */
      dng_list_reserve(&c, &st, op->nbad);
      memmove(st.list + op->nbad, st.list, (size_t)st.nlist * sizeof(uint32_t));
      memcpy(st.list, op->bad, (size_t)op->nbad * sizeof(uint32_t));
      st.nlist += op->nbad;
      break;
    case 6: /* ri->subFrame(roi) (common/RawImage.cpp:175-199) */
      if (!(op->w > 0 && op->h > 0))
        THROW_RDE(&c, "No positive crop area");
      if (!(op->w <= dw - op->left && op->h <= dh - op->top))
        break; /* "Crop skipped." */
      crop[0] = ox + op->left;
      crop[1] = oy + op->top;
      crop[2] = op->w;
      crop[3] = op->h;
      break;
    default: {
      const uint64_t nax = dng_round_up_div((uint64_t)op->w, op->colPitch),
                     nay = dng_round_up_div((uint64_t)op->h, op->rowPitch);
      const int is_scale = op->code == 12 || op->code == 13;
      const int by_row = op->code == 10 || op->code == 12;
      int* deltaI = NULL;
      uint64_t y, x;
      uint32_t p;
      if (op->lookup) {
        if (img->is_f32)
          THROW_RDE(&c, "Only 16 bit images supported");
      } else if (!img->is_f32) {
        /* DeltaRowOrCol::setup (:538-552) */
        uint32_t k;
        const double absLimit = 65535.0 / (double)65535.0F;
        const double maxLimit = ((double)(2147483647 - 512) / 65535.0) / (double)1024.0F;
        deltaI = (int*)malloc(((size_t)op->ndelta + 1) * sizeof(int));
        if (!deltaI)
          THROW_RDE(&c, "out of memory");
        for (k = 0; k < op->ndelta; k++) {
          const float f = op->deltaF[k];
          const int ok = is_scale ? (f >= 0.0F && (double)f <= maxLimit) : ((double)fabsf(f) <= absLimit);
          if (!ok) {
            free(deltaI);
            THROW_RDE(&c, "Got float %f which is unacceptable.", (double)f);
          }
          deltaI[k] = (int)((is_scale ? 1024.0F : 65535.0F) * f);
        }
      }
      /* PixelOpcode::applyOP (:390-409) */
      for (y = 0; y < nay; y++) {
        uint8_t* rowp = (uint8_t*)img->data + (size_t)(oy + op->top + (int)(op->rowPitch * y)) * (size_t)img->pitch;
        for (x = 0; x < nax; x++) {
          for (p = 0; p < op->planes; p++) {
            const size_t s = (size_t)ox * cpp + op->firstPlane + (size_t)(op->left + (int)(op->colPitch * x)) * cpp + p;
            const uint64_t sel = by_row ? y : x;
            if (img->is_f32) {
              float* px = (float*)rowp + s;
              *px = is_scale ? op->deltaF[sel] * *px : op->deltaF[sel] + *px;
            } else {
              uint16_t* px = (uint16_t*)rowp + s;
              int v;
              if (op->lookup)
                v = op->lookup[*px];
              else if (is_scale)
                v = (deltaI[sel] * (int)*px + 512) >> 10;
              else
                v = deltaI[sel] + (int)*px;
              *px = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
            }
          }
        }
      }
      free(deltaI);
      break;
    }
    }
    done = (int)i + 1;
  }
  if (applied)
    *applied = done;
  if (nbad)
    *nbad = st.nlist;
  for (i = 0; i < st.nlist && i < bad_cap && bad; i++)
    bad[i] = st.list[i];
  dng_free(&st);
  return RSO_OK;
}

/* ------------------------------------------------------------------
 * RawImageData::fixBadPixels (common/RawImage.cpp, common/RawImageDataU16.cpp)
 * ------------------------------------------------------------------ */
static void fix_bad_pixel(rso_image* img, const uint8_t* bad, uint32_t bpitch, uint32_t x, uint32_t y,
                          int component) {
  /* RawImageDataU16::fixBadPixel (:399-485) */
  int values[4] = {-1, -1, -1, -1}, dist[4] = {0, 0, 0, 0}, weight[4] = {0, 0, 0, 0};
  const int step = img->is_cfa ? 2 : 1;
  int x_find, y_find, total_dist_x, total_dist_y, total_shifts = 7, total_pixel = 0, i;
#define PIX(r, c) (((uint16_t*)((uint8_t*)img->data + (size_t)(r) * (size_t)img->pitch))[(c)])
#define ISBAD(r, c) ((bad[(size_t)bpitch * (size_t)(r) + ((c) >> 3)] >> ((c)&7)) & 1)
  x_find = (int)x - step;
  while (x_find >= 0 && values[0] < 0) {
    if (!ISBAD(y, x_find)) {
      values[0] = PIX(y, x_find + component);
      dist[0] = (int)x - x_find;
    }
    x_find -= step;
  }
  x_find = (int)x + step;
  while (x_find < img->w && values[1] < 0) {
    if (!ISBAD(y, x_find)) {
      values[1] = PIX(y, x_find + component);
      dist[1] = x_find - (int)x;
    }
    x_find += step;
  }
  y_find = (int)y - step;
  while (y_find >= 0 && values[2] < 0) {
    if (!ISBAD(y_find, x)) {
      values[2] = PIX(y_find, x + component);
      dist[2] = (int)y - y_find;
    }
    y_find -= step;
  }
  y_find = (int)y + step;
  while (y_find < img->h && values[3] < 0) {
    if (!ISBAD(y_find, x)) {
      values[3] = PIX(y_find, x + component);
      dist[3] = y_find - (int)y;
    }
    y_find += step;
  }
  total_dist_x = dist[0] + dist[1];
  if (total_dist_x) {
    weight[0] = dist[0] ? (total_dist_x - dist[0]) * 256 / total_dist_x : 0;
    weight[1] = 256 - weight[0];
    total_shifts++;
  }
  total_dist_y = dist[2] + dist[3];
  if (total_dist_y) {
    weight[2] = dist[2] ? (total_dist_y - dist[2]) * 256 / total_dist_y : 0;
    weight[3] = 256 - weight[2];
    total_shifts++;
  }
  for (i = 0; i < 4; i++)
    if (values[i] >= 0)
      total_pixel += values[i] * weight[i];
  total_pixel >>= total_shifts;
  PIX(y, x + component) = (uint16_t)(total_pixel < 0 ? 0 : (total_pixel > 65535 ? 65535 : total_pixel));
  if (img->cpp > 1 && component == 0)
    for (i = 1; i < img->cpp; i++)
      fix_bad_pixel(img, bad, bpitch, x, y, i);
#undef PIX
#undef ISBAD
}

int rso_fix_bad_pixels(rso_image* img, const uint32_t* positions, uint32_t npositions, rso_err* e) {
  rso_ctx c;
  rso_err le;
  uint8_t* volatile map = NULL;
  uint32_t bpitch, i;
  int y, x, gw;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (setjmp(c.jb)) {
    free((void*)map);
    return c.e->code;
  }
  if (img->is_f32)
    THROW_RDE(&c, "restated for UINT16 images");
  /* transferBadPixelsToMap (:211-229) */
  if (!npositions)
    return RSO_OK;
  bpitch = (uint32_t)((((img->w + 7) / 8) + 15) / 16 * 16); /* createBadPixelMap (:201-209) */
  map = (uint8_t*)calloc((size_t)bpitch * (size_t)img->h, 1);
  if (!map)
    THROW_RDE(&c, "out of memory");
  for (i = 0; i < npositions; i++) {
    const uint32_t px = positions[i] & 0xffff, py = positions[i] >> 16;
    if ((int)px >= img->w || (int)py >= img->h) /* (assert in the reference) */
      THROW_RDE(&c, "bad pixel position outside the image");
    ((uint8_t*)map)[(size_t)bpitch * py + (px >> 3)] |= (uint8_t)(1 << (px & 7));
  }
  /* fixBadPixelsThread (:297-323): blocks of 32 pixels, (w + 15) / 32 of them per row */
  gw = (img->w + 15) / 32;
  for (y = 0; y < img->h; y++) {
    for (x = 0; x < gw; x++) {
      const uint8_t* block = (const uint8_t*)map + (size_t)bpitch * (size_t)y + (size_t)x * 4;
      int bi, bj;
      if (!(block[0] | block[1] | block[2] | block[3]))
        continue;
      for (bi = 0; bi < 4; bi++)
        for (bj = 0; bj < 8; bj++)
          if ((block[bi] >> bj) & 1)
            fix_bad_pixel(img, (const uint8_t*)map, bpitch, (uint32_t)(x * 32 + bi * 8 + bj), (uint32_t)y, 0);
    }
  }
  free((void*)map);
  return RSO_OK;
}

/* ------------------------------------------------------------------
 * RawImageData::sixteenBitLookup / RawImageDataU16::doLookup
 * ------------------------------------------------------------------ */
int rso_sixteen_bit_lookup(rso_image* img, const uint16_t* table, int dither, rso_err* e) {
  rso_ctx c;
  rso_err le;
  const int gw = img->w * img->cpp;
  int y, x;
  c.e = e ? e : &le;
  c.e->code = RSO_OK;
  c.e->msg[0] = 0;
  if (setjmp(c.jb))
    return c.e->code;
  if (img->is_f32)
    THROW_RDE(&c, "Unexpected data type");
  if (!table) /* sixteenBitLookup (:373-376): no table, nothing to do */
    return RSO_OK;
  for (y = 0; y < img->h; y++) { /* FULL_IMAGE: uncropped_dim.y rows */
    uint16_t* row = (uint16_t*)((uint8_t*)img->data + (size_t)y * (size_t)img->pitch);
    if (dither) {
      uint32_t v = (uint32_t)(img->w + y * 13) ^ 0x45694584u;
      for (x = 0; x < gw; x++) {
        const uint16_t p = row[x];
        const uint32_t base = table[2 * p + 0], delta = table[2 * p + 1];
        uint32_t pix;
        v = 15700u * (v & 65535u) + (v >> 16);
        pix = base + ((delta * (v & 2047u) + 1024u) >> 12);
        row[x] = (uint16_t)(pix > 65535u ? 65535u : pix); /* clampBits(pix, 16) */
      }
    } else {
      for (x = 0; x < gw; x++)
        row[x] = table[row[x]];
    }
  }
  return RSO_OK;
}

/* ---- synthetic frame generator (see rs_oracle.h) ---- */
void rso_image_model(uint32_t w, uint32_t h, uint32_t seed, uint16_t* out, uint64_t sums[2]) {
  uint32_t s = seed;
  uint64_t s0 = 0, s1 = 0;
  for (uint32_t y = 0; y < h; ++y) {
    uint16_t* row = out + (size_t)y * w;
    for (uint32_t x = 0; x < w; ++x) {
      s = s * 1664525u + 1013904223u;
      const uint32_t v = (2000u + ((7u * x + 3u * y) & 1023u) + (s >> 26) - 32u) & 0x3FFFu;
      row[x] = (uint16_t)v;
      s0 += v;
      s1 += (uint64_t)v * ((((uint64_t)31 * x + (uint64_t)17 * y) & 0xFFFFu) | 1u);
    }
  }
  if (sums) {
    sums[0] = s0;
    sums[1] = s1;
  }
}
