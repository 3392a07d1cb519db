// Front end of the OAKE tower: turns crops into the conv1 patch matrix (im2col) that the tcgen05
// GEMM consumes.
//
//  * im2col_pixels: (B,3,224,224) fp32 CLIP-normalised crops, the tensor the reference feeds to
//    `model.encode_image` / `model.visual` (oadp/oake/globals.py:54-57, blocks.py:126-129,
//    objects.py:319-330) -> act_t [B*P, 3072].  stride 32 / pad 0 (P = 49) is CLIP's conv1;
//    stride 16 / pad 15 (P = 196) is the objects surgery of objects.py:299-301.
#include "kernels.cuh"

namespace oake {

namespace {

constexpr int kImg = 224;
constexpr int kPatch = 32;
constexpr int kCols = 3 * kPatch * kPatch;  // 3072
constexpr int kChunks = kCols / 8;          // 16-byte output chunks per patch row

__global__ void __launch_bounds__(256)
im2col_pixels_kernel(const float* __restrict__ pixels, act_t* __restrict__ patches, long long total,
                     int stride, int pad, int grid) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int chunk = static_cast<int>(idx % kChunks);
  const long long prow = idx / kChunks;  // b * P + gy * grid + gx
  const int P = grid * grid;
  const int b = static_cast<int>(prow / P);
  const int g = static_cast<int>(prow - static_cast<long long>(b) * P);
  const int gy = g / grid, gx = g - gy * grid;
  const int c = chunk / (kPatch * kPatch / 8);
  const int rem = chunk - c * (kPatch * kPatch / 8);
  const int ky = rem / (kPatch / 8);
  const int kx0 = (rem - ky * (kPatch / 8)) * 8;
  const int y = gy * stride - pad + ky;
  const int x0 = gx * stride - pad + kx0;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (y >= 0 && y < kImg) {
    const float* src = pixels + (static_cast<size_t>(b) * 3 + c) * kImg * kImg + static_cast<size_t>(y) * kImg;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int x = x0 + j;
      if (x >= 0 && x < kImg) v[j] = __ldg(src + x);
    }
  }
  uint4 u;
  u.x = pack2(v[0], v[1]);
  u.y = pack2(v[2], v[3]);
  u.z = pack2(v[4], v[5]);
  u.w = pack2(v[6], v[7]);
  *reinterpret_cast<uint4*>(patches + prow * kCols + chunk * 8) = u;
}

}  // namespace

cudaError_t launch_im2col_pixels(cudaStream_t st, const float* pixels, act_t* patches, int B,
                                 int stride, int pad, int grid) {
  if (B <= 0) return cudaSuccess;
  const long long total = static_cast<long long>(B) * grid * grid * kChunks;
  const long long blocks = (total + 255) / 256;
  im2col_pixels_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(pixels, patches, total, stride,
                                                                      pad, grid);
  return cudaGetLastError();
}

}  // namespace oake
