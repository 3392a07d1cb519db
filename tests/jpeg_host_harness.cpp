// CPU test harness (tests only; never linked into liboake_b200.so): runs the per-item functions of
// oadp_b200/csrc/jpeg_core.cuh -- the ones the CUDA kernels wrap -- over a whole file with plain
// loops, so that tests/test_jpeg_core_host.py can check them against Pillow without a GPU.
#include <stdlib.h>

#include <string>
#include <vector>

#include "../oadp_b200/csrc/jpeg_core.cuh"
#include "../oadp_b200/csrc/jpeg_parse.h"

using namespace oake;

extern "C" {

// Returns the parser's code (0 ok, 1 malformed, 2 unsupported), or 3 if the entropy data was damaged.
// `out` must hold width * height * 3 bytes (call with out == NULL first to get the size).
int harness_decode(const uint8_t* data, size_t len, uint8_t* out, int* width, int* height) {
  oake_jpeg_desc d;
  std::string why;
  const int rc = jpeg::parse(data, len, &d, &why);
  *width = static_cast<int>(d.width);
  *height = static_cast<int>(d.height);
  if (rc != 0 || out == nullptr) return rc;
  uint64_t scratch_off = 0;
  jpeg::place(&d, /*file_off=*/0, /*out_off=*/0, &scratch_off);
  std::vector<uint8_t> scratch(scratch_off, 0);
  jpeg::HuffView views[4];
  const oake_jpeg_huff* t[4] = {&d.dc[0], &d.dc[1], &d.ac[0], &d.ac[1]};
  for (int i = 0; i < 4; ++i) views[i] = {t[i]->look, t[i]->maxcode, t[i]->valoff, t[i]->huffval};
  if (jpeg::decode_scan(d, data, views, scratch.data()) != 0) return 3;
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
