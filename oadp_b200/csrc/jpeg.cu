// GPU JPEG decode in front of the OAKE resize: compressed files in, uint8 HWC RGB in
// the same arena `oake_resize_u8` reads -- the step oadp/oake/base.py:53 does with Pillow on the host.
// Integer / byte work, HBM- and latency-bound; results are bit-identical to Pillow's (the arithmetic
// is in jpeg_core.cuh).
//
//   entropy kernel   one warp per image: the Huffman tables go to shared memory, the 32 lanes clear
//                    the image's coefficient blocks, then lane 0 walks the bit stream (the code is
//                    inherently serial inside a restart-free scan; the batch supplies the
//                    parallelism -- one image per warp, up to 32 resident warps per SM).
//   idct kernel      one thread per 8x8 block: dequantise, islow IDCT, range limit -> uint8 planes.
//   colour kernel    one thread per pixel pair: fancy chroma upsampling + YCbCr -> RGB, HWC stores.
#include <cuda_runtime.h>
#include <stddef.h>

#include <string>

#include "../../include/oake_b200.h"
#include "jpeg_core.cuh"
#include "jpeg_parse.h"

namespace oake {
int fail_msg(const char* fmt, ...);  // encoder.cu
}

using namespace oake;

namespace {

__global__ void __launch_bounds__(32) jpeg_entropy_kernel(const uint8_t* __restrict__ bytes,
                                                          const oake_jpeg_desc* __restrict__ descs,
                                                          uint8_t* __restrict__ scratch, int32_t* __restrict__ status) {
  __shared__ oake_jpeg_huff tables[4];  // dc0 dc1 ac0 ac1
  // geometry part of the descriptor (everything in front of the quantisation tables): the only part
  // decode_scan touches, kept next to the tables so that the serial loop never waits on global memory
  // for it
  constexpr int kHeadBytes = offsetof(oake_jpeg_desc, quant);
  static_assert(kHeadBytes % 8 == 0, "descriptor layout");
  __shared__ __align__(16) uint8_t head[kHeadBytes];
  const int lane = threadIdx.x;
  {
    const oake_jpeg_desc& g = descs[blockIdx.x];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&g.dc[0]);  // dc[2] and ac[2] are contiguous
    uint32_t* dst = reinterpret_cast<uint32_t*>(tables);
    for (int i = lane; i < static_cast<int>(sizeof(tables) / 4); i += 32) dst[i] = src[i];
    const uint32_t* hs = reinterpret_cast<const uint32_t*>(&g);
    for (int i = lane; i < kHeadBytes / 4; i += 32) reinterpret_cast<uint32_t*>(head)[i] = hs[i];
  }
  __syncwarp();
  const oake_jpeg_desc& d = *reinterpret_cast<const oake_jpeg_desc*>(head);
  for (uint32_t c = 0; c < d.ncomp; ++c) {
    const oake_jpeg_comp& k = d.comp[c];
    uint4* p = reinterpret_cast<uint4*>(scratch + k.coef_off);
    const uint32_t n16 = k.blocks_w * k.blocks_h * 8;  // 128 bytes per block
    for (uint32_t i = lane; i < n16; i += 32) p[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncwarp();
  if (lane == 0) {
    jpeg::HuffView views[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) views[t] = {tables[t].look, tables[t].maxcode, tables[t].valoff, tables[t].huffval};
    status[blockIdx.x] = jpeg::decode_scan(d, bytes, views, scratch);
  }
}

__global__ void __launch_bounds__(128) jpeg_idct_kernel(const oake_jpeg_desc* __restrict__ descs,
                                                        uint8_t* __restrict__ scratch) {
  const oake_jpeg_desc& d = descs[blockIdx.y];
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.total_blocks) return;
  uint32_t c = 0;
  while (c + 1 < d.ncomp && b >= d.comp[c].blocks_w * d.comp[c].blocks_h) {
    b -= d.comp[c].blocks_w * d.comp[c].blocks_h;
    ++c;
  }
  const oake_jpeg_comp& k = d.comp[c];
  const uint32_t by = b / k.blocks_w, bx = b - by * k.blocks_w;
  // the block's 64 coefficients as eight 16-byte loads
  __align__(16) int16_t coef[64];
  const uint4* src = reinterpret_cast<const uint4*>(scratch + k.coef_off + static_cast<uint64_t>(b) * 128);
#pragma unroll
  for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(coef)[i] = src[i];
  __align__(8) uint8_t px[64];
  jpeg::idct_block(coef, d.quant[k.quant], px, 8);
  const uint32_t pitch = k.blocks_w * 8;
  uint8_t* dst = scratch + k.plane_off + static_cast<uint64_t>(by * 8) * pitch + bx * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) *reinterpret_cast<uint2*>(dst + static_cast<uint64_t>(r) * pitch) = reinterpret_cast<const uint2*>(px)[r];
}

__global__ void __launch_bounds__(256) jpeg_colour_kernel(const oake_jpeg_desc* __restrict__ descs,
                                                          const uint8_t* __restrict__ scratch,
                                                          uint8_t* __restrict__ out) {
  const oake_jpeg_desc& d = descs[blockIdx.y];
  const uint32_t pairs_w = (d.width + 1) / 2;
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pairs_w * d.height) return;
  const uint32_t Y = p / pairs_w, X = (p - Y * pairs_w) * 2;
  uint8_t* dst = out + d.out_off + (static_cast<uint64_t>(Y) * d.width + X) * 3;
  uint8_t rgb[6];
  jpeg::pixel_rgb(d, scratch, X, Y, rgb);
  const bool two = X + 1 < d.width;
  if (two) jpeg::pixel_rgb(d, scratch, X + 1, Y, rgb + 3);
#pragma unroll
  for (int i = 0; i < 3; ++i) dst[i] = rgb[i];
  if (two) {
#pragma unroll
    for (int i = 3; i < 6; ++i) dst[i] = rgb[i];
  }
}

}  // namespace

extern "C" {

size_t oake_jpeg_desc_bytes(void) { return sizeof(oake_jpeg_desc); }

int oake_jpeg_parse(const uint8_t* data, size_t len, oake_jpeg_desc* desc) {
  if (!data || !desc) return fail_msg("NULL argument");
  std::string why;
  const int rc = jpeg::parse(data, len, desc, &why);
  if (rc != 0) fail_msg("oake_jpeg_parse: %s", why.c_str());
  return rc;
}

int oake_jpeg_place(oake_jpeg_desc* desc, uint64_t file_off, uint64_t out_off, uint64_t* scratch_off) {
  if (!desc || !scratch_off) return fail_msg("NULL argument");
  if (desc->ncomp != 1 && desc->ncomp != 3) return fail_msg("descriptor was not produced by oake_jpeg_parse");
  jpeg::place(desc, file_off, out_off, scratch_off);
  return 0;
}

int oake_jpeg_decode(const uint8_t* bytes, const oake_jpeg_desc* descs_host, const oake_jpeg_desc* descs_dev, int n,
                     void* scratch, uint8_t* out, int32_t* status, void* stream) {
  if (n < 0) return fail_msg("negative count");
  if (n == 0) return 0;
  if (n > 65535) return fail_msg("at most 65535 images per call");
  if (!bytes || !descs_host || !descs_dev || !scratch || !out || !status) return fail_msg("NULL buffer");
  uint32_t max_blocks = 0, max_pairs = 0;
  for (int i = 0; i < n; ++i) {
    const oake_jpeg_desc& d = descs_host[i];
    if ((d.ncomp != 1 && d.ncomp != 3) || d.width == 0 || d.height == 0 || d.total_blocks == 0)
      return fail_msg("descriptor %d was not produced by oake_jpeg_parse", i);
    max_blocks = d.total_blocks > max_blocks ? d.total_blocks : max_blocks;
    const uint32_t pairs = ((d.width + 1) / 2) * d.height;
    max_pairs = pairs > max_pairs ? pairs : max_pairs;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* scr = static_cast<uint8_t*>(scratch);
  jpeg_entropy_kernel<<<n, 32, 0, st>>>(bytes, descs_dev, scr, status);
  jpeg_idct_kernel<<<dim3((max_blocks + 127) / 128, n), 128, 0, st>>>(descs_dev, scr);
  jpeg_colour_kernel<<<dim3((max_pairs + 255) / 256, n), 256, 0, st>>>(descs_dev, scr, out);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("oake_jpeg_decode launch: %s", cudaGetErrorString(e));
}

}  // extern "C"
