// CPU test harness (tests only; never linked into liboake_b200.so): runs the per-item functions of
// oadp_b200/csrc/jpeg_core.cuh -- the ones the CUDA kernels wrap -- over a whole file with plain
// loops, so that tests/test_jpeg_core_host.py can check them against Pillow without a GPU.
#include <stddef.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../oadp_b200/csrc/jpeg_core.cuh"
#include "../oadp_b200/csrc/jpeg_parse.h"

using namespace oake;

extern "C" {

// Returns the parser's code (0 ok, 1 malformed, 2 unsupported), or 3 if the entropy data was damaged.
// parallel != 0: the subsequence-parallel entropy decode (files without restart markers; files with them
// are decoded interval by interval either way), else the serial one; *sync_rounds (optional) = rounds step 2 needed.
// `out` must hold width * height * 3 bytes (call with out == NULL first to get the size).
int harness_decode(const uint8_t* data, size_t len, uint8_t* out, int* width, int* height, int parallel, int* sync_rounds) {
  oake_jpeg_desc d;
  std::string why;
  const int rc = jpeg::parse(data, len, &d, &why);
  *width = static_cast<int>(d.width);
  *height = static_cast<int>(d.height);
  if (rc != 0 || out == nullptr) return rc;
  uint64_t scratch_off = 0;
  std::vector<uint8_t> stream(jpeg::stream_bound(d));
  oake_jpeg_desc parsed = d;
  jpeg::stage(parsed, data, stream.data(), /*stream_off=*/0, /*out_off=*/0, &scratch_off, &d);
  std::vector<uint8_t> scratch(scratch_off, 0);
  static_assert(offsetof(oake_jpeg_desc, ac) == offsetof(oake_jpeg_desc, dc) + 2 * sizeof(oake_jpeg_huff), "dc[2], ac[2] contiguous");
  const oake_jpeg_huff* views = &d.dc[0];
  if (parallel && d.restart_interval == 0) {
    // the five steps of the subsequence-parallel decode (jpeg_core.cuh), one "thread" after the other
    const jpeg::McuMap map = jpeg::make_mcu_map(d);
    const uint32_t bits = static_cast<uint32_t>(d.scan_len * 8);
    const uint32_t n = (bits + jpeg::kSubBits - 1) / jpeg::kSubBits;
    if (n == 0 || n > d.sync_slots) return 3;
    std::vector<uint64_t> entry(n), exit_(n);
    std::vector<uint32_t> count(n), first(n);
    for (uint32_t i = 0; i < n; ++i) {
      entry[i] = jpeg::pack_state(i * jpeg::kSubBits, 0, 0);
      const jpeg::SubResult r = jpeg::decode_subsequence<false>(d, map, stream.data(), views, entry[i], (i + 1) * jpeg::kSubBits, 0, nullptr);
      exit_[i] = r.exit;
      count[i] = r.count;
    }
    int rounds = 0;
    for (bool changed = true; changed; ++rounds) {
      changed = false;
      const std::vector<uint64_t> seen = exit_;  // every thread of a round reads the previous round's states
      for (uint32_t i = 1; i < n; ++i) {
        if (seen[i - 1] == entry[i]) continue;
        entry[i] = seen[i - 1];
        const jpeg::SubResult r = jpeg::decode_subsequence<false>(d, map, stream.data(), views, entry[i], (i + 1) * jpeg::kSubBits, 0, nullptr);
        exit_[i] = r.exit;
        count[i] = r.count;
        changed = true;
      }
    }
    if (sync_rounds) *sync_rounds = rounds;
    uint32_t total = 0;
    for (uint32_t i = 0; i < n; ++i) {
      first[i] = total;
      total += count[i];
    }
    if (total < d.total_blocks) return 3;
    bool bad = false;
    for (uint32_t i = 0; i < n; ++i)
      bad |= jpeg::decode_subsequence<true>(d, map, stream.data(), views, entry[i], (i + 1) * jpeg::kSubBits, first[i], scratch.data()).bad;
    if (bad) return 3;
    int32_t pred[3] = {0, 0, 0};
    for (uint32_t B = 0; B < d.total_blocks; ++B) {
      int16_t* blk = jpeg::scan_block(d, map, scratch.data(), B);
      const int c = static_cast<int>(map.comp(B % map.bpm));
      pred[c] += blk[0];
      blk[0] = static_cast<int16_t>(pred[c]);
    }
  } else if (jpeg::decode_scan(d, stream.data(), views, scratch.data()) != 0) {
    return 3;
  }
  for (uint32_t c = 0; c < d.ncomp; ++c) {
    const oake_jpeg_comp& k = d.comp[c];
    for (uint32_t b = 0; b < k.blocks_w * k.blocks_h; ++b) {
      const uint32_t by = b / k.blocks_w, bx = b % k.blocks_w;
      const int16_t* coef = reinterpret_cast<const int16_t*>(scratch.data() + k.coef_off) + static_cast<size_t>(b) * 64;
      jpeg::idct_block(coef, d.quant[k.quant], scratch.data() + k.plane_off + static_cast<size_t>(by) * 8 * k.blocks_w * 8 + bx * 8,
                       k.blocks_w * 8);
    }
  }
  for (uint32_t y = 0; y < d.height; ++y)
    for (uint32_t x = 0; x < d.width; ++x) jpeg::pixel_rgb(d, scratch.data(), x, y, out + (static_cast<size_t>(y) * d.width + x) * 3);
  return 0;
}

}  // extern "C"
