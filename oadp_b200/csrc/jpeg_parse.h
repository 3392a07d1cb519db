// Host-side JPEG header parser: markers -> oake_jpeg_desc (geometry, quantisation tables, Huffman
// lookup tables).  Plain C++ (no CUDA) so that the CPU test harness can compile it too.  The
// envelope is what oake_jpeg_decode implements; everything else reports "unsupported" and the caller
// decodes that file with Pillow, as oadp/oake/base.py:53 does for every file.
#pragma once

#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/oake_b200.h"

namespace oake {
namespace jpeg {

enum ParseResult { kOk = 0, kMalformed = 1, kUnsupported = OAKE_JPEG_UNSUPPORTED };

inline uint64_t align256(uint64_t v) { return (v + 255u) & ~static_cast<uint64_t>(255u); }

// T.81 Annex C code assignment + the lookup tables of oake_jpeg_huff.
inline bool build_huff(const uint8_t counts[16], const uint8_t* symbols, int n_symbols, oake_jpeg_huff* t) {
  memset(t, 0, sizeof(*t));
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    t->valoff[l] = k - code;
    for (int i = 0; i < counts[l - 1]; ++i, ++k, ++code) {
      if (k >= n_symbols || code >= (1 << l)) return false;
      t->huffval[k] = symbols[k];
      if (l <= 9) {
        const int lo = code << (9 - l);
        for (int j = 0; j < (1 << (9 - l)); ++j) t->look[lo + j] = static_cast<uint16_t>((l << 8) | symbols[k]);
      }
    }
    t->maxcode[l] = counts[l - 1] ? code - 1 : -1;
    code <<= 1;
  }
  t->maxcode[0] = -1;
  t->maxcode[17] = 0x7FFFFFFF;
  return true;
}

inline int parse(const uint8_t* p, size_t len, oake_jpeg_desc* d, std::string* why) {
  auto bad = [&](const char* m) {
    *why = m;
    return static_cast<int>(kMalformed);
  };
  auto unsupported = [&](const char* m) {
    *why = m;
    return static_cast<int>(kUnsupported);
  };
  memset(d, 0, sizeof(*d));
  if (len < 4 || p[0] != 0xFF || p[1] != 0xD8) return bad("not a JPEG file (no SOI)");
  size_t pos = 2;
  bool have_sof = false, jfif = false, adobe = false;
  int adobe_transform = -1;
  bool have_quant[4] = {false, false, false, false};
  bool have_dc[4] = {false, false, false, false}, have_ac[4] = {false, false, false, false};
  uint8_t comp_id[3] = {0, 0, 0};
  uint32_t frame_h[3] = {1, 1, 1}, frame_v[3] = {1, 1, 1};
  while (true) {
    // next marker (any number of 0xFF fill bytes in front of it)
    while (pos < len && p[pos] != 0xFF) ++pos;
    while (pos < len && p[pos] == 0xFF) ++pos;
    if (pos >= len) return bad("file ends before the scan");
    const uint8_t m = p[pos++];
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;  // stand-alone markers
    if (m == 0xD9) return bad("EOI before the scan");
    if (pos + 2 > len) return bad("truncated marker segment");
    const size_t seg = (static_cast<size_t>(p[pos]) << 8) | p[pos + 1];
    if (seg < 2 || pos + seg > len) return bad("marker segment runs past the end of the file");
    const uint8_t* s = p + pos + 2;
    const size_t n = seg - 2;
    if (m == 0xE0) {
      if (n >= 5 && memcmp(s, "JFIF\0", 5) == 0) jfif = true;
    } else if (m == 0xEE) {
      if (n >= 12 && memcmp(s, "Adobe", 5) == 0) {
        adobe = true;
        adobe_transform = s[11];
      }
    } else if (m == 0xDB) {  // DQT
      size_t i = 0;
      while (i < n) {
        const int pq = s[i] >> 4, tq = s[i] & 15;
        ++i;
        if (tq > 3 || pq > 1) return bad("bad quantisation table header");
        if (i + (pq ? 128u : 64u) > n) return bad("truncated quantisation table");
        static const uint8_t zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                       12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                       35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                       58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        for (int k = 0; k < 64; ++k) {
          const uint16_t q = pq ? static_cast<uint16_t>((s[i] << 8) | s[i + 1]) : s[i];
          i += pq ? 2 : 1;
          d->quant[tq][zz[k]] = q;
        }
        have_quant[tq] = true;
      }
    } else if (m == 0xC4) {  // DHT
      size_t i = 0;
      while (i < n) {
        if (i + 17 > n) return bad("truncated Huffman table");
        const int tc = s[i] >> 4, th = s[i] & 15;
        const uint8_t* counts = s + i + 1;
        int total = 0;
        for (int k = 0; k < 16; ++k) total += counts[k];
        if (tc > 1 || th > 3 || total > 256 || i + 17 + total > n) return bad("bad Huffman table");
        if (th > 1) return unsupported("Huffman table id above 1");
        if (!build_huff(counts, s + i + 17, total, tc ? &d->ac[th] : &d->dc[th])) return bad("bad Huffman table");
        (tc ? have_ac : have_dc)[th] = true;
        i += 17 + total;
      }
    } else if (m == 0xDD) {  // DRI
      if (n < 2) return bad("truncated DRI");
      d->restart_interval = (static_cast<uint32_t>(s[0]) << 8) | s[1];
    } else if (m == 0xC0 || m == 0xC1) {  // SOF0 / SOF1: baseline / extended sequential, Huffman
      if (have_sof) return bad("two frame headers");
      if (n < 6) return bad("truncated frame header");
      const int precision = s[0];
      d->height = (static_cast<uint32_t>(s[1]) << 8) | s[2];
      d->width = (static_cast<uint32_t>(s[3]) << 8) | s[4];
      d->ncomp = s[5];
      if (n < 6 + 3 * static_cast<size_t>(d->ncomp)) return bad("truncated frame header");
      have_sof = true;
      if (precision != 8) return unsupported("sample precision other than 8 bits");
      if (d->height == 0 || d->width == 0) return unsupported("frame size given by a DNL marker");
      if (d->ncomp != 1 && d->ncomp != 3) return unsupported("component count other than 1 or 3");
      for (uint32_t c = 0; c < d->ncomp; ++c) {
        comp_id[c] = s[6 + 3 * c];
        frame_h[c] = s[7 + 3 * c] >> 4;
        frame_v[c] = s[7 + 3 * c] & 15;
        d->comp[c].quant = s[8 + 3 * c];
        if (frame_h[c] < 1 || frame_h[c] > 4 || frame_v[c] < 1 || frame_v[c] > 4 || d->comp[c].quant > 3)
          return bad("bad component parameters");
      }
    } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
      // progressive, lossless, differential or arithmetic-coded frame; keep the size for the caller
      if (n >= 5) {
        d->height = (static_cast<uint32_t>(s[1]) << 8) | s[2];
        d->width = (static_cast<uint32_t>(s[3]) << 8) | s[4];
      }
      return unsupported("not a baseline / extended-sequential Huffman JPEG");
    } else if (m == 0xDA) {  // SOS
      if (!have_sof) return bad("scan before the frame header");
      if (n < 1 || n < 4 + 2 * static_cast<size_t>(s[0])) return bad("truncated scan header");
      const uint32_t ns = s[0];
      if (ns != d->ncomp) return unsupported("more than one scan");
      for (uint32_t c = 0; c < ns; ++c) {
        if (s[1 + 2 * c] != comp_id[c]) return unsupported("scan components out of frame order");
        d->comp[c].dc_tbl = s[2 + 2 * c] >> 4;
        d->comp[c].ac_tbl = s[2 + 2 * c] & 15;
        if (d->comp[c].dc_tbl > 1 || d->comp[c].ac_tbl > 1) return unsupported("Huffman table id above 1");
        if (!have_dc[d->comp[c].dc_tbl] || !have_ac[d->comp[c].ac_tbl]) return bad("scan uses an undefined Huffman table");
        if (!have_quant[d->comp[c].quant]) return bad("component uses an undefined quantisation table");
      }
      if (s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63 || s[3 + 2 * ns] != 0) return bad("bad spectral selection in a sequential scan");
      pos += seg;
      break;
    }
    pos += seg;
  }

  // colour space as libjpeg decides it (jdapimin.c default_decompress_parms)
  if (d->ncomp == 3) {
    bool ycc = true;
    if (jfif) ycc = true;
    else if (adobe) ycc = adobe_transform == 1;
    else if (comp_id[0] == 'R' && comp_id[1] == 'G' && comp_id[2] == 'B') ycc = false;
    if (!ycc) return unsupported("3-component file that is not YCbCr");
    if (frame_h[1] != 1 || frame_v[1] != 1 || frame_h[2] != 1 || frame_v[2] != 1)
      return unsupported("subsampled component other than by a luma factor");
    const uint32_t h = frame_h[0], v = frame_v[0];
    if (!((h == 1 && v == 1) || (h == 2 && v == 1) || (h == 2 && v == 2))) return unsupported("sampling other than 4:4:4, 4:2:2, 4:2:0");
    d->hmax = h;
    d->vmax = v;
    d->mcus_x = (d->width + 8 * h - 1) / (8 * h);
    d->mcus_y = (d->height + 8 * v - 1) / (8 * v);
    for (int c = 0; c < 3; ++c) {
      oake_jpeg_comp& k = d->comp[c];
      k.h = c == 0 ? h : 1;
      k.v = c == 0 ? v : 1;
      k.blocks_w = d->mcus_x * k.h;
      k.blocks_h = d->mcus_y * k.v;
      k.width = (d->width * k.h + h - 1) / h;
      k.height = (d->height * k.v + v - 1) / v;
    }
    // libjpeg only uses the triangle filter when the subsampled row has more than two samples
    if (h == 2 && d->comp[1].width <= 2) return unsupported("image too narrow for fancy upsampling");
  } else {
    // a single-component scan is never interleaved: one block per MCU whatever the frame header says
    d->hmax = d->vmax = 1;
    oake_jpeg_comp& k = d->comp[0];
    k.h = k.v = 1;
    k.blocks_w = d->mcus_x = (d->width + 7) / 8;
    k.blocks_h = d->mcus_y = (d->height + 7) / 8;
    k.width = d->width;
    k.height = d->height;
  }
  uint64_t off = 0;
  d->total_blocks = 0;
  for (uint32_t c = 0; c < d->ncomp; ++c) {
    oake_jpeg_comp& k = d->comp[c];
    const uint64_t blocks = static_cast<uint64_t>(k.blocks_w) * k.blocks_h;
    k.coef_off = off;
    off = align256(off + blocks * 128);
    d->total_blocks += static_cast<uint32_t>(blocks);
  }
  for (uint32_t c = 0; c < d->ncomp; ++c) {
    oake_jpeg_comp& k = d->comp[c];
    k.plane_off = off;
    off = align256(off + static_cast<uint64_t>(k.blocks_w) * k.blocks_h * 64);
  }
  d->scan_off = pos;
  d->scan_len = len - pos;
  // one slot per 1024 bits of entropy-coded data (jpeg_core.cuh kSubBits): entry state, exit state,
  // block count, first block -- 24 bytes
  d->restart_count = d->restart_interval ? (d->mcus_x * d->mcus_y + d->restart_interval - 1) / d->restart_interval : 0;
  d->sync_slots = static_cast<uint32_t>((d->scan_len * 8 + 1023) / 1024 + 1);
  d->sync_off = off;
  off = align256(off + static_cast<uint64_t>(d->sync_slots) * 24);
  d->scratch_bytes = off;
  d->out_off = 0;
  return kOk;
}

// Upper bound of what stage() writes for a file whose parsed descriptor is `d`.
inline uint64_t stream_bound(const oake_jpeg_desc& d) {
  return ((d.scan_len + 3) & ~static_cast<uint64_t>(3)) + 16 + 4 * static_cast<uint64_t>(d.restart_count);
}

// Copies the entropy-coded segment of `file` to `dst` WITHOUT the 0x00 stuffed after each 0xFF data byte,
// cut at the first marker that is not RSTn and followed by >= 16 zero bytes (what jpeg_core.cuh's
// BitReader expects) and then by the restart table (byte offset of the start of every restart interval,
// 0xFFFFFFFF where the marker is missing), and rebases the descriptor: `placed` = `parsed` with scan_off = stream_off (the
// offset `dst` will have in the device byte arena; a multiple of 4), scan_len = clean length, out_off,
// and the coefficient / plane offsets moved behind *scratch_off (which is advanced).  Returns the bytes
// written (a multiple of 4, <= stream_bound()).
inline uint64_t stage(const oake_jpeg_desc& parsed, const uint8_t* file, uint8_t* dst, uint64_t stream_off,
                      uint64_t out_off, uint64_t* scratch_off, oake_jpeg_desc* placed) {
  const uint8_t* s = file + parsed.scan_off;
  const uint8_t* end = s + parsed.scan_len;
  uint8_t* o = dst;
  // interval starts are collected here first: their place in `dst` depends on the clean length
  std::vector<uint32_t> starts;
  if (parsed.restart_count) {
    starts.assign(parsed.restart_count, 0xFFFFFFFFu);
    starts[0] = 0;
  }
  uint32_t next_interval = 1;
  while (s < end) {
    const uint8_t* ff = static_cast<const uint8_t*>(memchr(s, 0xFF, static_cast<size_t>(end - s)));
    const size_t run = static_cast<size_t>((ff ? ff : end) - s);
    memcpy(o, s, run);
    o += run;
    s += run;
    if (!ff) break;
    const uint8_t next = s + 1 < end ? s[1] : 0xD9;
    if (next == 0x00) {  // a data byte 0xFF
      *o++ = 0xFF;
      s += 2;
    } else if ((next & 0xF8) == 0xD0) {  // RSTn stays in the stream; the decoder knows where to expect it
      *o++ = 0xFF;
      *o++ = next;
      s += 2;
      if (next_interval < parsed.restart_count) starts[next_interval++] = static_cast<uint32_t>(o - dst);
    } else if (next == 0xFF) {  // fill byte in front of a marker
      s += 1;
    } else {
      break;  // EOI or any other marker: the scan ends here
    }
  }
  const uint64_t clean = static_cast<uint64_t>(o - dst);
  uint64_t total = ((clean + 3) & ~static_cast<uint64_t>(3)) + 16;
  memset(o, 0, static_cast<size_t>(total - clean));
  if (parsed.restart_count) {
    memcpy(dst + total, starts.data(), 4 * starts.size());
    total += 4 * starts.size();
  }
  if (placed != &parsed) *placed = parsed;
  const uint64_t base = align256(*scratch_off);
  for (uint32_t c = 0; c < placed->ncomp; ++c) {
    placed->comp[c].coef_off += base;
    placed->comp[c].plane_off += base;
  }
  placed->sync_off += base;
  placed->scan_off = stream_off;
  placed->scan_len = clean;
  placed->out_off = out_off;
  *scratch_off = base + placed->scratch_bytes;
  return total;
}

}  // namespace jpeg
}  // namespace oake
