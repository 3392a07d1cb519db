// Sanitizer fuzz driver (tests only): runs the JPEG header parser, the staging pass and both entropy
// decoders + IDCT + colour code of oadp_b200/csrc (through tests/jpeg_host_harness.cpp) over a file of
// [u32 length][bytes] records.  Built with -fsanitize=address,undefined by tests/test_jpeg_core_host.py:
// damaged input must be rejected or decoded to garbage, never read or write out of bounds.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "jpeg_host_harness.cpp"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  uint32_t len;
  long records = 0, decoded = 0;
  std::vector<uint8_t> out;
  while (fread(&len, 4, 1, f) == 1) {
    uint8_t* data = static_cast<uint8_t*>(malloc(len ? len : 1));  // exact size: any over-read is caught
    if (fread(data, 1, len, f) != len) return 2;
    for (int parallel = 0; parallel < 2; ++parallel) {
      int w = 0, h = 0, rounds = 0;
      int rc = harness_decode(data, len, nullptr, &w, &h, 0, nullptr);
      if (rc == 0 && static_cast<size_t>(w) * h <= 4000000) {
        out.assign(static_cast<size_t>(w) * h * 3, 0);
        rc = harness_decode(data, len, out.data(), &w, &h, parallel, &rounds);
        decoded += rc == 0;
      }
    }
    free(data);
    ++records;
  }
  printf("records %ld decoded %ld\n", records, decoded);
  return 0;
}
