// Per-item arithmetic of the JPEG decoder (entropy decode of one image, IDCT of one block, colour
// of one pixel), written as plain host/device functions: jpeg.cu wraps them in kernels, and the CPU
// test tier compiles the very same functions with g++ (tests/jpeg_host_harness.cpp) to check them
// against Pillow before any GPU time is spent.  Nothing here is exported by the shared library.
//
// What has to be matched bit for bit is libjpeg-turbo as Pillow drives it for
// `Image.open(f).convert('RGB')` (oadp/oake/base.py:53): Huffman sequential decode (ITU T.81 F.2.2),
// the "islow" 8x8 inverse DCT (13-bit fixed-point Loeffler-Ligtenberg-Moschytz, two passes with 2
// extra bits kept between them), "fancy" (triangle-filter) chroma upsampling for 2x1 and 2x2
// sampling, and the 16-bit fixed-point YCbCr -> RGB tables.
#pragma once

#include <stdint.h>

#include "../../include/oake_b200.h"

#if defined(__CUDACC__)
#define OAKE_HD __host__ __device__ __forceinline__
#else
#define OAKE_HD inline
#endif

namespace oake {
namespace jpeg {

// natural (row-major) index -> position in the zig-zag sequence the file stores.  The entropy decoder
// writes coefficients in file order; the IDCT reads them through this map with compile-time indices.
OAKE_HD constexpr int zigzag_pos(int natural) {
  const uint8_t t[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                         41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                         46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
  return t[natural];
}

// ------------------------------------------------------------------------------- entropy decode
// The scan arrives "clean": the host has already dropped the 0x00 stuffed after every 0xFF data byte
// and cut the segment at the first marker that is not RSTn (jpeg_parse.h, stage()), the stream starts
// on a 4-byte boundary and is followed by at least 16 zero bytes (and then by the restart table).  That turns the refill into one
// aligned 32-bit load per 32 consumed bits, requested one refill ahead of its use -- what matters in a
// loop that is one long dependency chain run by a single thread.
OAKE_HD uint32_t load_be32(const uint32_t* p) {
  const uint32_t w = *p;
  return (w >> 24) | ((w >> 8) & 0xFF00u) | ((w << 8) & 0xFF0000u) | (w << 24);
}

struct BitReader {
  const uint32_t* words;
  uint32_t pos, limit;  // next word to request / first word beyond the zero padding
  uint64_t buf;         // left-aligned: bit 63 is the next bit of the stream
  int cnt;              // valid bits in buf
  uint32_t nxt;         // word `pos - 1`, loaded but not yet in buf

  OAKE_HD uint32_t fetch() { return pos < limit ? load_be32(words + pos++) : (++pos, 0u); }
  OAKE_HD void start(const uint8_t* stream, uint64_t len) {
    words = reinterpret_cast<const uint32_t*>(stream);
    pos = 0;
    limit = static_cast<uint32_t>((len + 15) / 4);
    buf = static_cast<uint64_t>(fetch()) << 32;
    buf |= fetch();
    cnt = 64;
    nxt = fetch();
  }
  // positions the reader on an arbitrary bit of the stream
  OAKE_HD void start_at(const uint8_t* stream, uint64_t len, uint32_t bit) {
    words = reinterpret_cast<const uint32_t*>(stream);
    pos = bit >> 5;
    limit = static_cast<uint32_t>((len + 15) / 4);
    buf = static_cast<uint64_t>(fetch()) << 32;
    buf |= fetch();
    cnt = 64;
    nxt = fetch();
    skip(static_cast<int>(bit & 31u));
    refill();
  }
  OAKE_HD void refill() {  // afterwards cnt > 32: a code (<= 16 bits) and its value (<= 16 bits) fit
    if (cnt <= 32) {
      buf |= static_cast<uint64_t>(nxt) << (32 - cnt);
      cnt += 32;
      nxt = fetch();
    }
  }
  OAKE_HD uint32_t top32() const { return static_cast<uint32_t>(buf >> 32); }
  OAKE_HD void skip(int k) {
    buf <<= k;
    cnt -= k;
  }
  // bits of the stream consumed so far
  OAKE_HD uint64_t consumed() const { return static_cast<uint64_t>(pos - 1) * 32 - static_cast<uint64_t>(cnt); }
};

// One Huffman symbol: returns (code length << 8) | symbol without consuming anything.  A code that is
// in no table comes back as length 16, symbol 0, and sets *bad.
OAKE_HD uint32_t peek_symbol(uint32_t top, const oake_jpeg_huff* t, bool* bad) {
  const uint32_t e = t->look[top >> 23];
  if (e != 0) return e;
  const uint32_t c16 = top >> 16;
  for (int l = 10; l <= 16; ++l) {
    const int32_t code = static_cast<int32_t>(c16 >> (16 - l));
    if (code <= t->maxcode[l]) return (static_cast<uint32_t>(l) << 8) | t->huffval[(code + t->valoff[l]) & 0xFF];
  }
  *bad = true;
  return 16u << 8;
}

// T.81 F.2.2.1 EXTEND of the s (1..16) bits that follow the `len`-bit code at the top of `top`
OAKE_HD int32_t extend_after(uint32_t top, int len, int s) {
  const uint32_t t = top << len;
  const int32_t v = static_cast<int32_t>(t >> (32 - s));
  // leading bit 0 -> negative: v - (2^s - 1)
  return v - (static_cast<int32_t>(~t) >> 31 & ((1 << s) - 1));
}

// Serial entropy decode of MCUs [mcu_lo, mcu_hi) from a known entry point: byte `start` of the stream,
// DC predictions zero -- the start of the scan or of a restart interval -- into zero-initialised
// coefficient blocks.  `tables`: the four tables of the descriptor, dc[0], dc[1], ac[0], ac[1]
// (contiguous there; the kernels copy them to shared memory).  `limit`: byte offset the data of this
// range must not reach beyond.  Returns 0, or non-zero if the data was damaged / ran out.
OAKE_HD int decode_mcu_range(const oake_jpeg_desc& d, const uint8_t* stream, const oake_jpeg_huff* tables,
                             uint8_t* scratch, uint32_t mcu_lo, uint32_t mcu_hi, uint32_t start, uint32_t limit) {
  BitReader br;
  br.start_at(stream, d.scan_len, start * 8u);
  bool bad = false;
  int32_t pred[3] = {0, 0, 0};
  uint32_t my = mcu_lo / d.mcus_x, mx = mcu_lo - my * d.mcus_x;
  for (uint32_t m = mcu_lo; m < mcu_hi; ++m) {
    for (uint32_t c = 0; c < d.ncomp; ++c) {
      const oake_jpeg_comp& k = d.comp[c];
      const oake_jpeg_huff* dc = tables + k.dc_tbl;
      const oake_jpeg_huff* ac = tables + 2 + k.ac_tbl;
      int16_t* plane = reinterpret_cast<int16_t*>(scratch + k.coef_off);
      for (uint32_t v = 0; v < k.v; ++v) {
        for (uint32_t h = 0; h < k.h; ++h) {
          int16_t* blk = plane + (static_cast<uint64_t>(my * k.v + v) * k.blocks_w + (mx * k.h + h)) * 64;
          br.refill();
          {
            const uint32_t top = br.top32();
            const uint32_t e = peek_symbol(top, dc, &bad);
            const int len = static_cast<int>(e >> 8), s = static_cast<int>(e & 15u);
            if (s != 0) pred[c] += extend_after(top, len, s);
            br.skip(len + s);
            blk[0] = static_cast<int16_t>(pred[c]);
          }
          for (int i = 1; i < 64;) {
            br.refill();
            const uint32_t top = br.top32();
            const uint32_t e = peek_symbol(top, ac, &bad);
            const int len = static_cast<int>(e >> 8), r = static_cast<int>((e >> 4) & 15u), s = static_cast<int>(e & 15u);
            if (s == 0) {
              br.skip(len);
              if (r != 15) break;  // end of block
              i += 16;
              continue;
            }
            i += r;
            if (i < 64) blk[i] = static_cast<int16_t>(extend_after(top, len, s));
            br.skip(len + s);
            ++i;
          }
        }
      }
    }
    if (++mx == d.mcus_x) {
      mx = 0;
      ++my;
    }
  }
  return (bad || br.consumed() > static_cast<uint64_t>(limit) * 8) ? 1 : 0;
}

// Restart intervals: the host (jpeg_parse.h, stage()) leaves a table of the byte offset at which every
// interval starts behind the stream; an interval nobody found is 0xFFFFFFFF.  Each interval is an
// independent entry point (byte aligned, predictions reset), so one thread decodes one interval.
constexpr uint32_t kNoInterval = 0xFFFFFFFFu;

OAKE_HD const uint32_t* restart_table(const oake_jpeg_desc& d, const uint8_t* stream) {
  return reinterpret_cast<const uint32_t*>(stream + ((d.scan_len + 3) & ~static_cast<uint64_t>(3)) + 16);
}

OAKE_HD int decode_interval(const oake_jpeg_desc& d, const uint8_t* stream, const oake_jpeg_huff* tables,
                            uint8_t* scratch, uint32_t k) {
  const uint32_t* table = restart_table(d, stream);
  const uint32_t n_mcus = d.mcus_x * d.mcus_y;
  const uint32_t lo = k * d.restart_interval;
  const uint32_t hi = lo + d.restart_interval < n_mcus ? lo + d.restart_interval : n_mcus;
  const uint32_t start = table[k];
  if (start == kNoInterval) return 1;
  // the data must end in front of the next RSTn marker (2 bytes), the last interval at the end of the scan
  uint32_t limit = static_cast<uint32_t>(d.scan_len);
  if (k + 1 < d.restart_count && table[k + 1] != kNoInterval) limit = table[k + 1] - 2;
  return decode_mcu_range(d, stream, tables, scratch, lo, hi, start, limit);
}

// The whole image by one thread (CPU harness; the kernels spread intervals / subsequences over threads).
OAKE_HD int decode_scan(const oake_jpeg_desc& d, const uint8_t* bytes, const oake_jpeg_huff* tables, uint8_t* scratch) {
  const uint8_t* stream = bytes + d.scan_off;
  if (d.restart_interval == 0)
    return decode_mcu_range(d, stream, tables, scratch, 0, d.mcus_x * d.mcus_y, 0, static_cast<uint32_t>(d.scan_len));
  int bad = 0;
  for (uint32_t k = 0; k < d.restart_count; ++k) bad |= decode_interval(d, stream, tables, scratch, k);
  return bad;
}

// ---------------------------------------------------------------- parallel entropy decode
// A scan without restart markers has no marked entry points, but Huffman streams re-synchronise: a
// decoder started on a wrong bit falls into step with the true symbol sequence after a few dozen
// symbols.  So the stream is cut into subsequences of kSubBits bits, every subsequence gets its own
// thread, and the threads find their true entry states by iteration (Klein & Wiseman; for JPEG
// Weissenberger & Schmidt, "Accelerating JPEG decompression on GPUs"):
//   1. thread i decodes the symbols that START inside [entry_i, (i+1) kSubBits) -- entry_0 is the true
//      start, the others guess "block boundary at i kSubBits" -- and publishes where it stopped: bit
//      position, block of the MCU, zig-zag index;
//   2. repeat: thread i takes the stop state of thread i-1 as its entry; if that differs from the entry
//      it used, it decodes again.  Entry i is final after at most i rounds (in practice 2-3 in all);
//   3. the blocks completed per subsequence, prefix-summed, say which block each entry state is in;
//   4. every thread decodes once more, now writing coefficients (DC as the difference it reads);
//   5. the DC differences are prefix-summed per component in scan order.
// decode_subsequence is steps 1, 2 and 4 for one thread; jpeg.cu holds the CTA-level orchestration, the
// CPU harness the same thing with plain loops.
constexpr uint32_t kSubBits = 1024;

// which component / block of it each block of an MCU is: one nibble per block of the MCU (<= 6)
struct McuMap {
  uint32_t bpm;  // blocks per MCU
  uint32_t comp4, v4, h4;
  OAKE_HD uint32_t comp(uint32_t b) const { return (comp4 >> (4 * b)) & 15u; }
  OAKE_HD uint32_t v(uint32_t b) const { return (v4 >> (4 * b)) & 15u; }
  OAKE_HD uint32_t h(uint32_t b) const { return (h4 >> (4 * b)) & 15u; }
};

OAKE_HD McuMap make_mcu_map(const oake_jpeg_desc& d) {
  McuMap m;
  m.bpm = m.comp4 = m.v4 = m.h4 = 0;
  for (uint32_t c = 0; c < d.ncomp; ++c)
    for (uint32_t v = 0; v < d.comp[c].v; ++v)
      for (uint32_t h = 0; h < d.comp[c].h; ++h) {
        m.comp4 |= c << (4 * m.bpm);
        m.v4 |= v << (4 * m.bpm);
        m.h4 |= h << (4 * m.bpm);
        ++m.bpm;
      }
  return m;
}

// decoder state between two symbols: bit position | block of the MCU << 32 | zig-zag index << 40
OAKE_HD uint64_t pack_state(uint32_t bit, uint32_t b, uint32_t z) {
  return static_cast<uint64_t>(bit) | (static_cast<uint64_t>(b) << 32) | (static_cast<uint64_t>(z) << 40);
}

// coefficient block number `B` of the scan (MCU-major, then the MCU's block order)
OAKE_HD int16_t* scan_block(const oake_jpeg_desc& d, const McuMap& map, uint8_t* scratch, uint32_t B) {
  const uint32_t mcu = B / map.bpm, kb = B - mcu * map.bpm;
  const uint32_t my = mcu / d.mcus_x, mx = mcu - my * d.mcus_x;
  const oake_jpeg_comp& k = d.comp[map.comp(kb)];
  const uint64_t idx = static_cast<uint64_t>(my * k.v + map.v(kb)) * k.blocks_w + (mx * k.h + map.h(kb));
  return reinterpret_cast<int16_t*>(scratch + k.coef_off) + idx * 64;
}

struct SubResult {
  uint64_t exit;   // state after the last symbol that started inside the subsequence
  uint32_t count;  // blocks completed
  bool bad;        // WRITE only: a code outside the tables inside a real block
};

template <bool WRITE>
OAKE_HD SubResult decode_subsequence(const oake_jpeg_desc& d, const McuMap& map, const uint8_t* stream,
                                     const oake_jpeg_huff* tables, uint64_t entry, uint32_t end_bit, uint32_t B,
                                     uint8_t* scratch) {
  const uint32_t total_bits = static_cast<uint32_t>(d.scan_len * 8);
  if (end_bit > total_bits) end_bit = total_bits;
  uint32_t b = static_cast<uint32_t>(entry >> 32) & 0xFFu, z = static_cast<uint32_t>(entry >> 40) & 0xFFu;
  BitReader br;
  br.start_at(stream, d.scan_len, static_cast<uint32_t>(entry));
  SubResult r;
  r.count = 0;
  r.bad = false;
  bool bad = false;
  int16_t* blk = nullptr;
  if (WRITE && B < d.total_blocks) blk = scan_block(d, map, scratch, B);
  uint32_t bit = static_cast<uint32_t>(entry);
  // tables of the block being decoded (they change with the block, not with the symbol)
  const oake_jpeg_huff* dc = tables + d.comp[map.comp(b)].dc_tbl;
  const oake_jpeg_huff* ac = tables + 2 + d.comp[map.comp(b)].ac_tbl;
  while (bit < end_bit) {
    br.refill();
    const uint32_t top = br.top32();
    int used;
    if (z == 0) {
      const uint32_t e = peek_symbol(top, dc, &bad);
      const int len = static_cast<int>(e >> 8), s = static_cast<int>(e & 15u);
      if (WRITE && blk != nullptr) blk[0] = static_cast<int16_t>(s != 0 ? extend_after(top, len, s) : 0);
      used = len + s;
      z = 1;
    } else {
      const uint32_t e = peek_symbol(top, ac, &bad);
      const int len = static_cast<int>(e >> 8), run = static_cast<int>((e >> 4) & 15u), s = static_cast<int>(e & 15u);
      if (s == 0) {
        used = len;
        z = run == 15 ? z + 16 : 64;
      } else {
        z += static_cast<uint32_t>(run);
        if (WRITE && blk != nullptr && z < 64) blk[z] = static_cast<int16_t>(extend_after(top, len, s));
        used = len + s;
        ++z;
      }
    }
    br.skip(used);
    bit += static_cast<uint32_t>(used);
    if (WRITE && blk != nullptr && bad) r.bad = true;
    bad = false;
    if (z >= 64) {
      z = 0;
      b = b + 1 == map.bpm ? 0 : b + 1;
      dc = tables + d.comp[map.comp(b)].dc_tbl;
      ac = tables + 2 + d.comp[map.comp(b)].ac_tbl;
      ++r.count;
      if (WRITE) {
        ++B;
        blk = B < d.total_blocks ? scan_block(d, map, scratch, B) : nullptr;
      }
    }
  }
  r.exit = pack_state(bit, b, z);
  return r;
}

// ----------------------------------------------------------------------------------------- IDCT
constexpr int kConstBits = 13;
constexpr int kPass1Bits = 2;

// The IDCT adds and multiplies modulo 2^32 (unsigned): for a sound file nothing comes near the limits
// and the values are those of libjpeg's signed arithmetic; for damaged data the sums just wrap instead
// of being undefined.  Only the final shifts look at the value as a signed number.
typedef uint32_t idct_t;

OAKE_HD int32_t descale(idct_t x, int n) { return static_cast<int32_t>(x + (1u << (n - 1))) >> n; }

// the sample range-limit table of libjpeg behind the IDCT, as arithmetic: index = (x & 1023)
OAKE_HD uint8_t idct_range_limit(int32_t x) {
  const int32_t i = x & 1023;
  if (i < 128) return static_cast<uint8_t>(i + 128);
  if (i < 512) return 255;
  if (i < 896) return 0;
  return static_cast<uint8_t>(i - 896);
}

OAKE_HD idct_t fix(int32_t c) { return static_cast<idct_t>(c); }  // a (possibly negative) constant mod 2^32

// One 1-D pass over eight values (already dequantised / from the workspace); results left unscaled.
OAKE_HD void idct_1d(const idct_t in[8], idct_t out[8]) {
  // even part
  idct_t z2 = in[2], z3 = in[6];
  idct_t z1 = (z2 + z3) * fix(4433);
  const idct_t e2 = z1 + z3 * fix(-15137);
  const idct_t e3 = z1 + z2 * fix(6270);
  z2 = in[0];
  z3 = in[4];
  const idct_t e0 = (z2 + z3) << kConstBits;
  const idct_t e1 = (z2 - z3) << kConstBits;
  const idct_t t10 = e0 + e3, t13 = e0 - e3, t11 = e1 + e2, t12 = e1 - e2;
  // odd part
  idct_t o0 = in[7], o1 = in[5], o2 = in[3], o3 = in[1];
  z1 = o0 + o3;
  z2 = o1 + o2;
  z3 = o0 + o2;
  idct_t z4 = o1 + o3;
  const idct_t z5 = (z3 + z4) * fix(9633);
  o0 *= fix(2446);
  o1 *= fix(16819);
  o2 *= fix(25172);
  o3 *= fix(12299);
  z1 *= fix(-7373);
  z2 *= fix(-20995);
  z3 *= fix(-16069);
  z4 *= fix(-3196);
  z3 += z5;
  z4 += z5;
  o0 += z1 + z3;
  o1 += z2 + z4;
  o2 += z2 + z3;
  o3 += z1 + z4;
  out[0] = t10 + o3;
  out[7] = t10 - o3;
  out[1] = t11 + o2;
  out[6] = t11 - o2;
  out[2] = t12 + o1;
  out[5] = t12 - o1;
  out[3] = t13 + o0;
  out[4] = t13 - o0;
}

// coef: 64 int16 in zig-zag (file) order; quant: 64 uint16 in natural order; dst: 8 rows of 8 bytes.
OAKE_HD void idct_block(const int16_t* coef, const uint16_t* quant, uint8_t* dst, uint32_t pitch) {
  idct_t ws[64];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    idct_t in[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
      in[r] = static_cast<idct_t>(static_cast<int32_t>(coef[zigzag_pos(r * 8 + c)])) * static_cast<idct_t>(quant[r * 8 + c]);
    idct_1d(in, o);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[r * 8 + c] = static_cast<idct_t>(descale(o[r], kConstBits - kPass1Bits));
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    idct_t o[8];
    idct_1d(ws + r * 8, o);
#pragma unroll
    for (int c = 0; c < 8; ++c) dst[r * pitch + c] = idct_range_limit(descale(o[c], kConstBits + kPass1Bits + 3));
  }
}

// ------------------------------------------------------------------- chroma upsampling + colour
OAKE_HD uint8_t clamp_u8(int32_t v) { return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// Full-resolution chroma sample (X, Y) of a component stored at half horizontal (and, for v2, half
// vertical) resolution: libjpeg's h2v1 / h2v2 "fancy" upsampling, which weighs the nearer sample 3/4
// and the further one 1/4 in each subsampled direction and rounds alternately down / up.
OAKE_HD int32_t upsample_h2v1(const uint8_t* plane, uint32_t pitch, uint32_t n, uint32_t X, uint32_t Y) {
  const uint8_t* row = plane + static_cast<uint64_t>(Y) * pitch;
  const uint32_t i = X >> 1;
  const int32_t cur = row[i];
  if ((X & 1u) == 0u) return i == 0 ? cur : (3 * cur + row[i - 1] + 1) >> 2;
  return i == n - 1 ? cur : (3 * cur + row[i + 1] + 2) >> 2;
}

OAKE_HD int32_t upsample_h2v2(const uint8_t* plane, uint32_t pitch, uint32_t n, uint32_t rows, uint32_t X,
                              uint32_t Y) {
  const uint32_t j = Y >> 1;
  // the other row: above for the upper output row, below for the lower one; replicated at the edges
  const uint32_t jn = (Y & 1u) == 0u ? (j == 0 ? 0 : j - 1) : (j + 1 < rows ? j + 1 : rows - 1);
  const uint8_t* r0 = plane + static_cast<uint64_t>(j) * pitch;
  const uint8_t* r1 = plane + static_cast<uint64_t>(jn) * pitch;
  const uint32_t i = X >> 1;
  const int32_t cur = 3 * r0[i] + r1[i];
  if ((X & 1u) == 0u) {
    if (i == 0) return (cur * 4 + 8) >> 4;
    return (cur * 3 + (3 * r0[i - 1] + r1[i - 1]) + 8) >> 4;
  }
  if (i == n - 1) return (cur * 4 + 7) >> 4;
  return (cur * 3 + (3 * r0[i + 1] + r1[i + 1]) + 7) >> 4;
}

// libjpeg's YCbCr -> RGB: 16-bit fixed point, the red / blue terms rounded per table entry, the two
// green terms summed before the shift.
OAKE_HD void ycc_to_rgb(int32_t y, int32_t cb, int32_t cr, uint8_t* rgb) {
  const int32_t u = cb - 128, v = cr - 128;
  rgb[0] = clamp_u8(y + ((91881 * v + 32768) >> 16));
  rgb[1] = clamp_u8(y + ((-22554 * u + 32768 - 46802 * v) >> 16));
  rgb[2] = clamp_u8(y + ((116130 * u + 32768) >> 16));
}

// RGB of pixel (X, Y) of image `d` from its component planes.
OAKE_HD void pixel_rgb(const oake_jpeg_desc& d, const uint8_t* scratch, uint32_t X, uint32_t Y, uint8_t* rgb) {
  const oake_jpeg_comp& k0 = d.comp[0];
  const int32_t y = scratch[k0.plane_off + static_cast<uint64_t>(Y) * (k0.blocks_w * 8) + X];
  if (d.ncomp == 1) {
    rgb[0] = rgb[1] = rgb[2] = static_cast<uint8_t>(y);
    return;
  }
  int32_t c[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const oake_jpeg_comp& k = d.comp[1 + q];
    const uint8_t* plane = scratch + k.plane_off;
    const uint32_t pitch = k.blocks_w * 8;
    if (d.hmax == 1) c[q] = plane[static_cast<uint64_t>(Y) * pitch + X];
    else if (d.vmax == 1) c[q] = upsample_h2v1(plane, pitch, k.width, X, Y);
    else c[q] = upsample_h2v2(plane, pitch, k.width, k.height, X, Y);
  }
  ycc_to_rgb(y, c[0], c[1], rgb);
}

}  // namespace jpeg
}  // namespace oake
