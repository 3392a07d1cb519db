// Row-wise kernels of the OAKE tower: token assembly + ln_pre (the tail of SURVEY 2.2 K1), the
// stand-alone LayerNorm used for ln_post (K2; ln_1 / ln_2 are folded into the GEMM epilogues),
// final L2-normalise + fp16 cast (K8; reference: F.normalize(...).half() at
// oadp/oake/globals.py:58-59, blocks.py:130-133, objects.py:331-334).
// One warp per 768-wide row, statistics in fp32 registers, 128-bit global accesses.
#include "kernels.cuh"

namespace oake {

namespace {

constexpr int kWidth = 768;
constexpr float kLnEps = 1e-5f;

// v: 24 values of one row held by this lane as 6 float4 at float4 index lane + 32*j.
__device__ __forceinline__ void ln_normalize(float4 (&v)[6], const float* __restrict__ w,
                                             const float* __restrict__ b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) s += v[j].x + v[j].y + v[j].z + v[j].w;
  const float mean = warp_sum(s) * (1.0f / kWidth);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    v[j].x -= mean;
    v[j].y -= mean;
    v[j].z -= mean;
    v[j].w -= mean;
    ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / kWidth) + kLnEps);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float4 g = __ldg(w4 + lane + 32 * j);
    const float4 be = __ldg(b4 + lane + 32 * j);
    v[j].x = v[j].x * rstd * g.x + be.x;
    v[j].y = v[j].y * rstd * g.y + be.y;
    v[j].z = v[j].z * rstd * g.z + be.z;
    v[j].w = v[j].w * rstd * g.w + be.w;
  }
}

// Packs a normalised row to act_t, stores it, and returns (sum, sum of squares) of the STORED values.
__device__ __forceinline__ float2 store_row_act(act_t* dst, const float4 (&v)[6], int lane) {
  uint2* o2 = reinterpret_cast<uint2*>(dst);
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    uint2 u;
    u.x = pack2(v[j].x, v[j].y);
    u.y = pack2(v[j].z, v[j].w);
    o2[lane + 32 * j] = u;
    const float2 a = unpack2(u.x), b = unpack2(u.y);
    s += (a.x + a.y) + (b.x + b.y);
    ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
  }
  return make_float2(warp_sum(s), warp_sum(ss));
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const act_t* __restrict__ x, const float* __restrict__ w,
                 const float* __restrict__ b, act_t* __restrict__ out, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint2* x2 = reinterpret_cast<const uint2*>(x + static_cast<size_t>(row) * kWidth);
  float4 v[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const uint2 u = x2[lane + 32 * j];
    const float2 a = unpack2(u.x), c = unpack2(u.y);
    v[j] = make_float4(a.x, a.y, c.x, c.y);
  }
  ln_normalize(v, w, b, lane);
  store_row_act(out + static_cast<size_t>(row) * kWidth, v, lane);
}

__global__ void __launch_bounds__(256)
assemble_ln_pre_kernel(const float* __restrict__ patch_out, const float* __restrict__ class_emb,
                       const float* __restrict__ pos, const float* __restrict__ w,
                       const float* __restrict__ b, act_t* __restrict__ x, float2* __restrict__ stats,
                       int B, int P, int with_y, int src_grid, int pgrid) {
  const int rows = B * P + B;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bool is_cls = row >= B * P;
  size_t src_row = row;
  if (src_grid > 0 && !is_cls) {
    const int b = row / P, i = row - b * P;
    const int gy = i / pgrid, gx = i - gy * pgrid;
    src_row = static_cast<size_t>(b) * src_grid * src_grid + gy * src_grid + gx;
  }
  const float4* src = reinterpret_cast<const float4*>(is_cls ? class_emb : patch_out + src_row * kWidth);
  const int tok = is_cls ? 0 : 1 + row % P;
  const float4* p4 = reinterpret_cast<const float4*>(pos + static_cast<size_t>(tok) * kWidth);
  float4 v[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float4 a = src[lane + 32 * j];
    const float4 p = __ldg(p4 + lane + 32 * j);
    v[j] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
  ln_normalize(v, w, b, lane);
  const float2 st = store_row_act(x + static_cast<size_t>(row) * kWidth, v, lane);
  if (lane < kStatSlots) stats[static_cast<size_t>(row) * kStatSlots + lane] = lane == 0 ? st : make_float2(0.f, 0.f);
  if (is_cls && with_y) {  // side token y0 = x[CLS] right after ln_pre (objects.py:216-222)
    store_row_act(x + static_cast<size_t>(row + B) * kWidth, v, lane);
    if (lane < kStatSlots)
      stats[static_cast<size_t>(row + B) * kStatSlots + lane] = lane == 0 ? st : make_float2(0.f, 0.f);
  }
}

// F.normalize(e, dim=1, eps=1e-12) then .half(); dim == 512 -> 16 floats per lane.
__global__ void __launch_bounds__(256)
l2norm_half_kernel(const float* __restrict__ e, __half* __restrict__ out, int rows, int dim) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* r = e + static_cast<size_t>(row) * dim;
  float ss = 0.f;
  for (int i = lane; i < dim; i += 32) ss += r[i] * r[i];
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  for (int i = lane; i < dim; i += 32) out[static_cast<size_t>(row) * dim + i] = __float2half_rn(r[i] * inv);
}

}  // namespace

cudaError_t launch_layernorm(cudaStream_t st, const act_t* x, const float* w, const float* b,
                             act_t* out, int rows, int width) {
  if (width != kWidth) return cudaErrorInvalidValue;
  if (rows <= 0) return cudaSuccess;
  layernorm_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, w, b, out, rows);
  return cudaGetLastError();
}

cudaError_t launch_assemble_ln_pre(cudaStream_t st, const float* patch_out, const float* class_emb,
                                   const float* pos, const float* w, const float* b, act_t* x,
                                   float2* stats, int B, int P, int width, int with_y, int src_grid) {
  if (width != kWidth) return cudaErrorInvalidValue;
  if (B <= 0) return cudaSuccess;
  const int rows = B * P + B;
  int pgrid = 1;
  while (pgrid * pgrid < P) ++pgrid;
  assemble_ln_pre_kernel<<<(rows + 7) / 8, 256, 0, st>>>(patch_out, class_emb, pos, w, b, x, stats, B,
                                                         P, with_y, src_grid, pgrid);
  return cudaGetLastError();
}

cudaError_t launch_l2norm_half(cudaStream_t st, const float* e, __half* out, int rows, int dim) {
  if (rows <= 0) return cudaSuccess;
  l2norm_half_kernel<<<(rows + 7) / 8, 256, 0, st>>>(e, out, rows, dim);
  return cudaGetLastError();
}

}  // namespace oake
