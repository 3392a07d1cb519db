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
// MSB-first bit reader over the entropy-coded segment: removes the 0x00 stuffed after each 0xFF and
// stops at any marker (from then on it supplies zero bits and counts them as padding).
struct BitReader {
  const uint8_t* src;
  uint64_t pos, end;
  uint64_t bits;  // the low `n` bits are valid
  int n;
  int pad;  // zero bits appended after the data ran out

  OAKE_HD void reset() {
    bits = 0;
    n = 0;
    pad = 0;
  }
  OAKE_HD void refill() {  // brings n to >= 57 (so 32 bits can be used between two calls)
    while (n <= 56) {
      uint32_t b = 0;
      if (pos < end) {
        b = src[pos];
        if (b != 0xFFu) {
          ++pos;
        } else if (pos + 1 < end && src[pos + 1] == 0u) {
          pos += 2;
        } else {  // a marker (or a truncated file): stay on it
          b = 0;
          pad += 8;
        }
      } else {
        pad += 8;
      }
      bits = (bits << 8) | b;
      n += 8;
    }
  }
  OAKE_HD uint32_t peek16() const { return static_cast<uint32_t>(bits >> (n - 16)) & 0xFFFFu; }
  OAKE_HD void skip(int k) { n -= k; }
  OAKE_HD int32_t receive_extend(int s) {  // T.81 F.2.2.1 / F.2.4.3; s in 0..16
    if (s == 0) return 0;
    const int32_t v = static_cast<int32_t>((bits >> (n - s)) & ((1u << s) - 1u));
    n -= s;
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
  }
  // true when bits that were never in the file have been consumed
  OAKE_HD bool overran() const { return n < pad; }
  // end of a restart interval: drop the bits left in the last byte and step over RSTn
  OAKE_HD bool restart() {
    const bool bad = overran();
    reset();
    while (pos < end && src[pos] != 0xFFu) ++pos;                      // (nothing to skip in a sound file)
    while (pos + 1 < end && src[pos] == 0xFFu && src[pos + 1] == 0xFFu) ++pos;  // fill bytes
    if (pos + 1 < end && src[pos] == 0xFFu && (src[pos + 1] & 0xF8u) == 0xD0u) {
      pos += 2;
      return !bad;
    }
    return false;
  }
};

// One Huffman symbol.  `look`, `maxcode`, `valoff`, `huffval`: the four parts of an oake_jpeg_huff
// (possibly copied to faster memory).  Returns the symbol; a code that is in no table returns 0 and
// sets *bad.
OAKE_HD int decode_symbol(BitReader& br, const uint16_t* look, const int32_t* maxcode, const int32_t* valoff,
                          const uint8_t* huffval, bool* bad) {
  const uint32_t c16 = br.peek16();
  const uint32_t e = look[c16 >> 7];
  if (e != 0) {
    br.skip(static_cast<int>(e >> 8));
    return static_cast<int>(e & 0xFFu);
  }
  for (int l = 10; l <= 16; ++l) {
    const int32_t code = static_cast<int32_t>(c16 >> (16 - l));
    if (code <= maxcode[l]) {
      br.skip(l);
      return huffval[(code + valoff[l]) & 0xFF];
    }
  }
  *bad = true;
  br.skip(16);
  return 0;
}

struct HuffView {
  const uint16_t* look;
  const int32_t* maxcode;
  const int32_t* valoff;
  const uint8_t* huffval;
};

// Entropy-decodes one whole image into zero-initialised coefficient blocks.  `tables[0..1]` = DC,
// `tables[2..3]` = AC.  Returns 0, or non-zero if the data was damaged / ran out.
OAKE_HD int decode_scan(const oake_jpeg_desc& d, const uint8_t* bytes, const HuffView* tables, uint8_t* scratch) {
  BitReader br;
  br.src = bytes + d.scan_off;
  br.pos = 0;
  br.end = d.scan_len;
  br.reset();
  bool bad = false;
  int32_t pred[3] = {0, 0, 0};
  const uint32_t n_mcus = d.mcus_x * d.mcus_y;
  uint32_t until_restart = d.restart_interval;
  uint32_t mx = 0, my = 0;
  for (uint32_t m = 0; m < n_mcus; ++m) {
    if (d.restart_interval != 0 && until_restart == 0) {
      if (!br.restart()) bad = true;
      pred[0] = pred[1] = pred[2] = 0;
      until_restart = d.restart_interval;
    }
    for (uint32_t c = 0; c < d.ncomp; ++c) {
      const oake_jpeg_comp& k = d.comp[c];
      const HuffView dc = tables[k.dc_tbl], ac = tables[2 + k.ac_tbl];
      int16_t* plane = reinterpret_cast<int16_t*>(scratch + k.coef_off);
      for (uint32_t v = 0; v < k.v; ++v) {
        for (uint32_t h = 0; h < k.h; ++h) {
          int16_t* blk = plane + (static_cast<uint64_t>(my * k.v + v) * k.blocks_w + (mx * k.h + h)) * 64;
          br.refill();
          const int s = decode_symbol(br, dc.look, dc.maxcode, dc.valoff, dc.huffval, &bad) & 15;
          pred[c] += br.receive_extend(s);
          blk[0] = static_cast<int16_t>(pred[c]);
          for (int i = 1; i < 64;) {
            br.refill();
            const int rs = decode_symbol(br, ac.look, ac.maxcode, ac.valoff, ac.huffval, &bad);
            const int r = rs >> 4, sz = rs & 15;
            if (sz == 0) {
              if (r != 15) break;  // end of block
              i += 16;
              continue;
            }
            i += r;
            const int32_t val = br.receive_extend(sz);
            if (i < 64) blk[i] = static_cast<int16_t>(val);
            ++i;
          }
        }
      }
    }
    --until_restart;
    if (++mx == d.mcus_x) {
      mx = 0;
      ++my;
    }
  }
  return (bad || br.overran()) ? 1 : 0;
}

// ----------------------------------------------------------------------------------------- IDCT
constexpr int kConstBits = 13;
constexpr int kPass1Bits = 2;

OAKE_HD int32_t descale(int32_t x, int n) { return (x + (1 << (n - 1))) >> n; }

// the sample range-limit table of libjpeg behind the IDCT, as arithmetic: index = (x & 1023)
OAKE_HD uint8_t idct_range_limit(int32_t x) {
  const int32_t i = x & 1023;
  if (i < 128) return static_cast<uint8_t>(i + 128);
  if (i < 512) return 255;
  if (i < 896) return 0;
  return static_cast<uint8_t>(i - 896);
}

// One 1-D pass over eight values (already dequantised / from the workspace); results left unscaled.
OAKE_HD void idct_1d(const int32_t in[8], int32_t out[8]) {
  // even part
  int32_t z2 = in[2], z3 = in[6];
  int32_t z1 = (z2 + z3) * 4433;
  const int32_t e2 = z1 + z3 * (-15137);
  const int32_t e3 = z1 + z2 * 6270;
  z2 = in[0];
  z3 = in[4];
  const int32_t e0 = (z2 + z3) * (1 << kConstBits);
  const int32_t e1 = (z2 - z3) * (1 << kConstBits);
  const int32_t t10 = e0 + e3, t13 = e0 - e3, t11 = e1 + e2, t12 = e1 - e2;
  // odd part
  int32_t o0 = in[7], o1 = in[5], o2 = in[3], o3 = in[1];
  z1 = o0 + o3;
  z2 = o1 + o2;
  z3 = o0 + o2;
  int32_t z4 = o1 + o3;
  const int32_t z5 = (z3 + z4) * 9633;
  o0 *= 2446;
  o1 *= 16819;
  o2 *= 25172;
  o3 *= 12299;
  z1 *= -7373;
  z2 *= -20995;
  z3 *= -16069;
  z4 *= -3196;
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
  int32_t ws[64];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    int32_t in[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) in[r] = static_cast<int32_t>(coef[zigzag_pos(r * 8 + c)]) * static_cast<int32_t>(quant[r * 8 + c]);
    idct_1d(in, o);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[r * 8 + c] = descale(o[r], kConstBits - kPass1Bits);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int32_t o[8];
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
