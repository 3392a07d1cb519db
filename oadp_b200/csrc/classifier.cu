// Cosine classifier of oadp.dp, forward and backward.
//
//   h      = F.normalize(x W^T + b)                    NormalizedLinear      oadp/dp/utils.py:47-51
//   E      = [text rows as stored ; F.normalize(bg)]   BaseClassifier.embeddings  classifiers.py:49-57
//   logits = alpha * (h E^T) - shift, novel columns -inf while training          classifiers.py:59-68,
//            (Classifier: alpha = scaler, shift = bias; ViLD: alpha = 1/scaler)   82-83, 105-112
//
// Split exactly where the reference splits it: `_linear` is its own module whose OUTPUT (the
// normalised (N,512) tensor) is captured by the todd distiller hooks
// ('.roi_head._object_head.fc_cls._linear', configs/dp/models/*.py), so h is materialised and each
// half has its own backward; autograd adds the distillation gradient and the logits gradient on h.
// The dense parts run on the tower's tcgen05 GEMM (fp16 operands, fp32 accumulate and outputs);
// gradients are rescaled by a power of two derived from their max-abs before the fp16 cast so that
// neither loss-scaled nor tiny gradients leave the fp16 range.
#include <stdarg.h>
#include <stdio.h>

#include <string>

#include "kernels.cuh"

using namespace oake;

namespace oake {
int fail_msg(const char* fmt, ...);  // encoder.cu
}

namespace {

constexpr int kDim = 512;

// power-of-two scale that brings max|v| to ~2^13 (fp16 max is 2^16)
__device__ __forceinline__ float scale_from_maxabs(float m) {
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  return exp2f(floorf(13.f - log2f(m)));
}

__global__ void cast_kernel(const float* __restrict__ in, act_t* __restrict__ out, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    uint2 u;
    u.x = pack2(v.x, v.y);
    u.y = pack2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + i) = u;
  } else {
    for (long long j = i; j < n; ++j) out[j] = to_act(in[j]);
  }
}

// F.normalize(h_raw, dim=1, eps=1e-12): h = h_raw / max(||h_raw||, eps); one warp per 512-wide row
__global__ void __launch_bounds__(256)
l2norm_rows_kernel(const float* __restrict__ h_raw, float* __restrict__ h, float* __restrict__ inv_norm, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* r4 = reinterpret_cast<const float4*>(h_raw + static_cast<size_t>(row) * kDim);
  float4 v[4];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j] = r4[lane + 32 * j];
    ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
  }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  float4* o4 = reinterpret_cast<float4*>(h + static_cast<size_t>(row) * kDim);
#pragma unroll
  for (int j = 0; j < 4; ++j) o4[lane + 32 * j] = make_float4(v[j].x * inv, v[j].y * inv, v[j].z * inv, v[j].w * inv);
  if (lane == 0) inv_norm[row] = inv;
}

// Inference form of the same: also writes the normalised row in the tensor-core type, the A operand of the
// logits GEMM that follows (one pass over h_raw instead of l2norm + cast).
__global__ void __launch_bounds__(256)
l2norm_rows_act_kernel(const float* __restrict__ h_raw, float* __restrict__ h, act_t* __restrict__ h_act, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* r4 = reinterpret_cast<const float4*>(h_raw + static_cast<size_t>(row) * kDim);
  float4 v[4];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j] = r4[lane + 32 * j];
    ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
  }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  float4* o4 = reinterpret_cast<float4*>(h + static_cast<size_t>(row) * kDim);
  uint2* a2 = reinterpret_cast<uint2*>(h_act + static_cast<size_t>(row) * kDim);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 o = make_float4(v[j].x * inv, v[j].y * inv, v[j].z * inv, v[j].w * inv);
    o4[lane + 32 * j] = o;
    a2[lane + 32 * j] = make_uint2(pack2(o.x, o.y), pack2(o.z, o.w));
  }
}

// x in fp16 / bf16 -> act_t (a copy when the types agree is never launched: the caller passes x through)
template <typename T>
__global__ void cast_from_kernel(const T* __restrict__ in, act_t* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = to_act(static_cast<float>(in[i]));
}

// dh_raw = (dh - h (h . dh)) * inv_norm   -- backward of F.normalize
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ h, const float* __restrict__ inv_norm,
                  float* __restrict__ dh_raw, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* d4 = reinterpret_cast<const float4*>(dh + static_cast<size_t>(row) * kDim);
  const float4* h4 = reinterpret_cast<const float4*>(h + static_cast<size_t>(row) * kDim);
  float4 d[4], hh[4];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    d[j] = d4[lane + 32 * j];
    hh[j] = h4[lane + 32 * j];
    dot += d[j].x * hh[j].x + d[j].y * hh[j].y + d[j].z * hh[j].z + d[j].w * hh[j].w;
  }
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
  float4* o4 = reinterpret_cast<float4*>(dh_raw + static_cast<size_t>(row) * kDim);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o4[lane + 32 * j] = make_float4((d[j].x - hh[j].x * dot) * inv, (d[j].y - hh[j].y * dot) * inv,
                                    (d[j].z - hh[j].z * dot) * inv, (d[j].w - hh[j].w * dot) * inv);
}

// e_act [k_pad,512]: text rows as stored, normalised bg row, zero padding; et_act [512,k_pad] optional
__global__ void __launch_bounds__(256)
pack_embeddings_kernel(const float* __restrict__ text, const float* __restrict__ bg, int num_all, int k_pad,
                       act_t* __restrict__ e_act, act_t* __restrict__ et_act) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= k_pad) return;
  float v[16];
  float scale = 1.f;
  const float* src = nullptr;
  if (row < num_all) {
    src = text + static_cast<size_t>(row) * kDim;
  } else if (row == num_all && bg != nullptr) {
    src = bg;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = src ? src[lane + 32 * j] : 0.f;
  if (row == num_all && bg != nullptr) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) ss += v[j] * v[j];
    scale = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const act_t a = to_act(v[j] * scale);
    e_act[static_cast<size_t>(row) * kDim + lane + 32 * j] = a;
    if (et_act != nullptr) et_act[static_cast<size_t>(lane + 32 * j) * k_pad + row] = a;
  }
}

__global__ void maxabs_kernel(const float* __restrict__ in, long long n, float* __restrict__ out) {
  float m = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = fabsf(in[i]);
    if (isfinite(v)) m = fmaxf(m, v);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
}

// out[c][r] = act(in[r][c] * scale), r < R (zero for R <= r < r_pad); 32x32 tiles through smem
__global__ void __launch_bounds__(256)
transpose_cast_kernel(const float* __restrict__ in, act_t* __restrict__ out, int R, int Ccols, int r_pad,
                      const float* __restrict__ maxabs) {
  __shared__ float tile[32][33];
  const float scale = maxabs ? scale_from_maxabs(*maxabs) : 1.f;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < R && c < Ccols) ? in[static_cast<size_t>(r) * Ccols + c] * scale : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (c < Ccols && r < r_pad) out[static_cast<size_t>(c) * r_pad + r] = to_act(tile[tx][j]);
  }
}

__global__ void scaled_cast_kernel(const float* __restrict__ in, act_t* __restrict__ out, long long n,
                                   const float* __restrict__ maxabs) {
  const float scale = scale_from_maxabs(*maxabs);
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = to_act(in[i] * scale);
}

// dy_act[n][k] = act(dY[n][k] * alpha * scale) for live columns, 0 for -inf / padding columns
__global__ void grad_logits_kernel(const float* __restrict__ dy, act_t* __restrict__ out, int N, int K, int k_pad,
                                   float alpha, int ninf_lo, int ninf_hi, const float* __restrict__ maxabs) {
  const float scale = scale_from_maxabs(*maxabs) * alpha;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(N) * k_pad) return;
  const int k = static_cast<int>(i % k_pad);
  float v = 0.f;
  if (k < K && !(k >= ninf_lo && k < ninf_hi)) {
    v = dy[i];
    if (!isfinite(v)) v = 0.f;
  }
  out[i] = to_act(v * scale);
}

__global__ void unscale_kernel(float* __restrict__ buf, long long n, const float* __restrict__ maxabs, float extra) {
  const float inv = extra / scale_from_maxabs(*maxabs);
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) buf[i] *= inv;
}

// db[c] = sum_r m[r][c]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ m, float* __restrict__ out, int R, int Ccols) {
  __shared__ float part[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
  float s = 0.f;
  if (c < Ccols)
    for (int r = ty; r < R; r += 8) s += m[static_cast<size_t>(r) * Ccols + c];
  part[ty][threadIdx.x & 31] = s;
  __syncthreads();
  if (ty == 0 && c < Ccols) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += part[j][threadIdx.x & 31];
    out[c] = t;
  }
}

// gradient of the learnable background row: dE = sum_n dlogits[n][bg] * alpha * h[n][:], then back
// through F.normalize(bg).  One CTA of 512 threads (thread = feature).
__global__ void __launch_bounds__(512)
bg_grad_kernel(const float* __restrict__ dy, const float* __restrict__ h, const float* __restrict__ bg, int N,
               int k_pad, int col, float alpha, float* __restrict__ dbg) {
  __shared__ float red[16];
  const int f = threadIdx.x;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) {
    float g = dy[static_cast<size_t>(n) * k_pad + col];
    if (!isfinite(g)) g = 0.f;
    acc += g * h[static_cast<size_t>(n) * kDim + f];
  }
  acc *= alpha;
  const float b = bg[f];
  float ss = warp_sum(b * b);
  float dot = warp_sum(b * acc);
  __shared__ float red2[16];
  if ((f & 31) == 0) {
    red[f >> 5] = ss;
    red2[f >> 5] = dot;
  }
  __syncthreads();
  float tss = 0.f, tdot = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    tss += red[j];
    tdot += red2[j];
  }
  const float norm = fmaxf(sqrtf(tss), 1e-12f);
  const float e = b / norm;
  dbg[f] = (acc - e * (tdot / norm)) / norm;
}

size_t up(size_t v) { return (v + 1023) / 1024 * 1024; }

struct Carver {
  uint8_t* base;
  size_t off = 0, cap;
  Carver(void* ws, size_t bytes) : base(static_cast<uint8_t*>(ws)), cap(bytes) {
    off = (1024 - (reinterpret_cast<uintptr_t>(ws) & 1023)) & 1023;
  }
  void* take(size_t bytes) {
    void* p = base + off;
    off += up(bytes);
    return p;
  }
  bool ok() const { return off <= cap; }
};

int num_sms() {
  int dev = 0, ns = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&ns, cudaDevAttrMultiProcessorCount, dev);
  return ns;
}

#define CK(call)                                                                    \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) return fail_msg("%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

int pad64(int n) { return (n + 63) / 64 * 64; }

}  // namespace

extern "C" {

int oake_classifier_workspace_bytes(int N, int in_features, int k_pad, size_t* out_bytes) {
  if (!out_bytes || N < 0 || in_features <= 0 || k_pad <= 0) return fail_msg("bad argument");
  const size_t n = static_cast<size_t>(N), np = pad64(N);
  // the largest of the four calls, with slack for alignment
  size_t nl_fwd = up(n * in_features * 2) + up(512ull * in_features * 2) + up(n * kDim * 4);
  size_t nl_bwd = up(n * kDim * 4) + up(n * kDim * 2) + up(512ull * np * 2) + up(static_cast<size_t>(in_features) * np * 2) +
                  up(static_cast<size_t>(in_features) * 512 * 2) + 1024;
  size_t cl_fwd = up(n * kDim * 2) + up(static_cast<size_t>(k_pad) * kDim * 2);
  size_t cl_bwd = up(n * k_pad * 2) + 2 * up(static_cast<size_t>(k_pad) * kDim * 2) + 1024;
  size_t fused = up(n * in_features * 2) + up(n * kDim * 4) + up(n * kDim * 2);
  size_t m = nl_fwd;
  if (fused > m) m = fused;
  if (nl_bwd > m) m = nl_bwd;
  if (cl_fwd > m) m = cl_fwd;
  if (cl_bwd > m) m = cl_bwd;
  *out_bytes = m + 8192;
  return 0;
}

int oake_normalized_linear_fwd(const float* x, const float* w, const float* b, int N, int in_features, float* h,
                               float* inv_norm, void* ws, size_t ws_bytes, void* stream) {
  if (N == 0) return 0;
  if (!x || !w || !b || !h || !inv_norm || !ws) return fail_msg("NULL buffer");
  if (in_features % 64 != 0) return fail_msg("in_features must be a multiple of 64, got %d", in_features);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(ws, ws_bytes);
  act_t* x_act = static_cast<act_t*>(c.take(static_cast<size_t>(N) * in_features * 2));
  act_t* w_act = static_cast<act_t*>(c.take(512ull * in_features * 2));
  float* h_raw = static_cast<float*>(c.take(static_cast<size_t>(N) * kDim * 4));
  if (!c.ok()) return fail_msg("workspace too small");
  const long long nx = static_cast<long long>(N) * in_features, nw = 512ll * in_features;
  cast_kernel<<<static_cast<unsigned>((nx / 4 + 256) / 256), 256, 0, st>>>(x, x_act, nx);
  cast_kernel<<<static_cast<unsigned>((nw / 4 + 256) / 256), 256, 0, st>>>(w, w_act, nw);
  CUtensorMap tmA, tmW;
  if (make_tmap_act_2d(&tmA, x_act, N, in_features, 128) || make_tmap_act_2d(&tmW, w_act, kDim, in_features, gemm_block_n(kDim)))
    return fail_msg("cuTensorMapEncodeTiled failed");
  GemmEpilogue ep{b, nullptr, nullptr, nullptr, nullptr, h_raw, kDim, 0, 1, 0};
  CK(launch_gemm(st, tmA, tmW, N, kDim, in_features, ep, num_sms()));
  l2norm_rows_kernel<<<(N + 7) / 8, 256, 0, st>>>(h_raw, h, inv_norm, N);
  CK(cudaGetLastError());
  return 0;
}

// ---- inference fast path: NormalizedLinear + cosine logits in ONE call, prepared operands ------------------
int oake_classifier_prepare(const float* w, int in_features, const float* text, const float* bg, int num_all,
                            int k_pad, void* w_act, void* e_act, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (w != nullptr) {
    if (!w_act || in_features <= 0) return fail_msg("bad argument");
    const long long nw = 512ll * in_features;
    cast_kernel<<<static_cast<unsigned>((nw / 4 + 256) / 256), 256, 0, st>>>(w, static_cast<act_t*>(w_act), nw);
  }
  if (text != nullptr) {
    const int K = num_all + (bg ? 1 : 0);
    if (!e_act || k_pad % 128 != 0 || k_pad < K) return fail_msg("k_pad must be a multiple of 128 and >= %d", K);
    pack_embeddings_kernel<<<(k_pad + 7) / 8, 256, 0, st>>>(text, bg, num_all, k_pad, static_cast<act_t*>(e_act), nullptr);
  }
  CK(cudaGetLastError());
  return 0;
}

int oake_classifier_fwd(const void* x, int x_dtype, const void* w_act, const float* bias, const void* e_act, int N,
                        int in_features, int k_pad, float alpha, float shift, int ninf_lo, int ninf_hi, float* h,
                        float* logits, void* ws, size_t ws_bytes, void* stream) {
  if (N == 0) return 0;
  if (!x || !w_act || !bias || !e_act || !h || !logits || !ws) return fail_msg("NULL buffer");
  if (in_features % 64 != 0) return fail_msg("in_features must be a multiple of 64, got %d", in_features);
  if (k_pad % 128 != 0) return fail_msg("k_pad must be a multiple of 128");
#ifdef OAKE_USE_BF16
  const int act_code = OAKE_DTYPE_BF16;
#else
  const int act_code = OAKE_DTYPE_F16;
#endif
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(ws, ws_bytes);
  act_t* x_act = static_cast<act_t*>(c.take(static_cast<size_t>(N) * in_features * 2));
  float* h_raw = static_cast<float*>(c.take(static_cast<size_t>(N) * kDim * 4));
  act_t* h_act = static_cast<act_t*>(c.take(static_cast<size_t>(N) * kDim * 2));
  if (!c.ok()) return fail_msg("workspace too small");
  const long long nx = static_cast<long long>(N) * in_features;
  const act_t* a_ptr = x_act;
  if (x_dtype == act_code) {
    a_ptr = static_cast<const act_t*>(x);  // already in the tensor-core type (mmcv fp16 mode): no copy at all
  } else if (x_dtype == OAKE_DTYPE_F32) {
    cast_kernel<<<static_cast<unsigned>((nx / 4 + 256) / 256), 256, 0, st>>>(static_cast<const float*>(x), x_act, nx);
  } else if (x_dtype == OAKE_DTYPE_F16) {
    cast_from_kernel<<<static_cast<unsigned>((nx + 255) / 256), 256, 0, st>>>(static_cast<const __half*>(x), x_act, nx);
  } else if (x_dtype == OAKE_DTYPE_BF16) {
    cast_from_kernel<<<static_cast<unsigned>((nx + 255) / 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), x_act, nx);
  } else {
    return fail_msg("x_dtype %d: expected OAKE_DTYPE_F32 / F16 / BF16", x_dtype);
  }
  if ((reinterpret_cast<uintptr_t>(a_ptr) & 15) != 0) return fail_msg("x must be 16-byte aligned");
  CUtensorMap tmA, tmW, tmH, tmE;
  if (make_tmap_act_2d(&tmA, a_ptr, N, in_features, 128) || make_tmap_act_2d(&tmW, w_act, kDim, in_features, gemm_block_n(kDim)) ||
      make_tmap_act_2d(&tmH, h_act, N, kDim, 128) || make_tmap_act_2d(&tmE, e_act, k_pad, kDim, gemm_block_n(k_pad)))
    return fail_msg("cuTensorMapEncodeTiled failed");
  const int ns = num_sms();
  GemmEpilogue ep1{bias, nullptr, nullptr, nullptr, nullptr, h_raw, kDim, 0, 1, 0};
  CK(launch_gemm(st, tmA, tmW, N, kDim, in_features, ep1, ns));
  l2norm_rows_act_kernel<<<(N + 7) / 8, 256, 0, st>>>(h_raw, h, h_act, N);
  GemmEpilogue ep2{nullptr, nullptr, nullptr, nullptr, nullptr, logits, k_pad, 0, 1, 0};
  ep2.alpha = alpha;
  ep2.shift = shift;
  ep2.ninf_lo = ninf_lo;
  ep2.ninf_hi = ninf_hi;
  CK(launch_gemm(st, tmH, tmE, N, k_pad, kDim, ep2, ns));
  CK(cudaGetLastError());
  return 0;
}

int oake_normalized_linear_bwd(const float* x, const float* w, const float* h, const float* inv_norm,
                               const float* dh, int N, int in_features, float* dx, float* dw, float* db, void* ws,
                               size_t ws_bytes, void* stream) {
  if (N == 0) return 0;
  if (!x || !w || !h || !inv_norm || !dh || !ws) return fail_msg("NULL buffer");
  // dx / dw are GEMMs whose OUTPUT width is in_features (tiles of 128 columns)
  if (in_features % 128 != 0) return fail_msg("the backward pass needs in_features to be a multiple of 128");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int np = pad64(N);
  Carver c(ws, ws_bytes);
  float* dh_raw = static_cast<float*>(c.take(static_cast<size_t>(N) * kDim * 4));
  act_t* dhr_act = static_cast<act_t*>(c.take(static_cast<size_t>(N) * kDim * 2));
  act_t* dhr_t = static_cast<act_t*>(c.take(512ull * np * 2));
  act_t* x_t = static_cast<act_t*>(c.take(static_cast<size_t>(in_features) * np * 2));
  act_t* w_t = static_cast<act_t*>(c.take(static_cast<size_t>(in_features) * 512 * 2));
  float* maxabs = static_cast<float*>(c.take(16));
  if (!c.ok()) return fail_msg("workspace too small");
  const int ns = num_sms();
  l2norm_bwd_kernel<<<(N + 7) / 8, 256, 0, st>>>(dh, h, inv_norm, dh_raw, N);
  CK(cudaMemsetAsync(maxabs, 0, 4, st));
  const long long nd = static_cast<long long>(N) * kDim;
  maxabs_kernel<<<256, 256, 0, st>>>(dh_raw, nd, maxabs);
  if (db) colsum_kernel<<<kDim / 32, 256, 0, st>>>(dh_raw, db, N, kDim);
  if (dx) {
    scaled_cast_kernel<<<static_cast<unsigned>((nd + 255) / 256), 256, 0, st>>>(dh_raw, dhr_act, nd, maxabs);
    transpose_cast_kernel<<<dim3((in_features + 31) / 32, kDim / 32), 256, 0, st>>>(w, w_t, kDim, in_features, kDim, nullptr);
    CUtensorMap tmA, tmB;
    if (make_tmap_act_2d(&tmA, dhr_act, N, kDim, 128) ||
        make_tmap_act_2d(&tmB, w_t, in_features, kDim, gemm_block_n(in_features)))
      return fail_msg("cuTensorMapEncodeTiled failed");
    GemmEpilogue ep{nullptr, nullptr, nullptr, nullptr, nullptr, dx, in_features, 0, 1, 0};
    CK(launch_gemm(st, tmA, tmB, N, in_features, kDim, ep, ns));
    const long long n = static_cast<long long>(N) * in_features;
    unscale_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(dx, n, maxabs, 1.f);
  }
  if (dw) {
    transpose_cast_kernel<<<dim3(kDim / 32, (np + 31) / 32), 256, 0, st>>>(dh_raw, dhr_t, N, kDim, np, maxabs);
    transpose_cast_kernel<<<dim3((in_features + 31) / 32, (np + 31) / 32), 256, 0, st>>>(x, x_t, N, in_features, np, nullptr);
    CUtensorMap tmA, tmB;
    if (make_tmap_act_2d(&tmA, dhr_t, kDim, np, 128) ||
        make_tmap_act_2d(&tmB, x_t, in_features, np, gemm_block_n(in_features)))
      return fail_msg("cuTensorMapEncodeTiled failed");
    GemmEpilogue ep{nullptr, nullptr, nullptr, nullptr, nullptr, dw, in_features, 0, 1, 0};
    CK(launch_gemm(st, tmA, tmB, kDim, in_features, np, ep, ns));
    const long long n = 512ll * in_features;
    unscale_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(dw, n, maxabs, 1.f);
  }
  CK(cudaGetLastError());
  return 0;
}

int oake_cosine_logits_fwd(const float* h, const float* text, const float* bg, int N, int num_all, int k_pad,
                           float alpha, float shift, int ninf_lo, int ninf_hi, float* logits, void* ws,
                           size_t ws_bytes, void* stream) {
  if (N == 0) return 0;
  if (!h || !text || !logits || !ws) return fail_msg("NULL buffer");
  const int K = num_all + (bg ? 1 : 0);
  if (k_pad % 128 != 0 || k_pad < K) return fail_msg("k_pad must be a multiple of 128 and >= %d", K);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(ws, ws_bytes);
  act_t* h_act = static_cast<act_t*>(c.take(static_cast<size_t>(N) * kDim * 2));
  act_t* e_act = static_cast<act_t*>(c.take(static_cast<size_t>(k_pad) * kDim * 2));
  if (!c.ok()) return fail_msg("workspace too small");
  const long long nh = static_cast<long long>(N) * kDim;
  cast_kernel<<<static_cast<unsigned>((nh / 4 + 256) / 256), 256, 0, st>>>(h, h_act, nh);
  pack_embeddings_kernel<<<(k_pad + 7) / 8, 256, 0, st>>>(text, bg, num_all, k_pad, e_act, nullptr);
  CUtensorMap tmA, tmB;
  if (make_tmap_act_2d(&tmA, h_act, N, kDim, 128) || make_tmap_act_2d(&tmB, e_act, k_pad, kDim, gemm_block_n(k_pad)))
    return fail_msg("cuTensorMapEncodeTiled failed");
  GemmEpilogue ep{nullptr, nullptr, nullptr, nullptr, nullptr, logits, k_pad, 0, 1, 0};
  ep.alpha = alpha;
  ep.shift = shift;
  ep.ninf_lo = ninf_lo;
  ep.ninf_hi = ninf_hi;
  CK(launch_gemm(st, tmA, tmB, N, k_pad, kDim, ep, num_sms()));
  return 0;
}

int oake_cosine_logits_bwd(const float* h, const float* text, const float* bg, const float* dlogits, int N,
                           int num_all, int k_pad, float alpha, int ninf_lo, int ninf_hi, float* dh, float* dbg,
                           void* ws, size_t ws_bytes, void* stream) {
  if (N == 0) return 0;
  if (!h || !text || !dlogits || !dh || !ws) return fail_msg("NULL buffer");
  const int K = num_all + (bg ? 1 : 0);
  if (k_pad % 128 != 0 || k_pad < K) return fail_msg("k_pad must be a multiple of 128 and >= %d", K);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(ws, ws_bytes);
  act_t* dy_act = static_cast<act_t*>(c.take(static_cast<size_t>(N) * k_pad * 2));
  act_t* e_act = static_cast<act_t*>(c.take(static_cast<size_t>(k_pad) * kDim * 2));
  act_t* et_act = static_cast<act_t*>(c.take(static_cast<size_t>(k_pad) * kDim * 2));
  float* maxabs = static_cast<float*>(c.take(16));
  if (!c.ok()) return fail_msg("workspace too small");
  const long long n = static_cast<long long>(N) * k_pad;
  CK(cudaMemsetAsync(maxabs, 0, 4, st));
  maxabs_kernel<<<256, 256, 0, st>>>(dlogits, n, maxabs);
  grad_logits_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(dlogits, dy_act, N, K, k_pad, 1.f, ninf_lo,
                                                                             ninf_hi, maxabs);
  pack_embeddings_kernel<<<(k_pad + 7) / 8, 256, 0, st>>>(text, bg, num_all, k_pad, e_act, et_act);
  CUtensorMap tmA, tmB;
  if (make_tmap_act_2d(&tmA, dy_act, N, k_pad, 128) || make_tmap_act_2d(&tmB, et_act, kDim, k_pad, gemm_block_n(kDim)))
    return fail_msg("cuTensorMapEncodeTiled failed");
  GemmEpilogue ep{nullptr, nullptr, nullptr, nullptr, nullptr, dh, kDim, 0, 1, 0};
  CK(launch_gemm(st, tmA, tmB, N, kDim, k_pad, ep, num_sms()));
  const long long nh = static_cast<long long>(N) * kDim;
  unscale_kernel<<<static_cast<unsigned>((nh + 255) / 256), 256, 0, st>>>(dh, nh, maxabs, alpha);
  if (bg && dbg) bg_grad_kernel<<<1, 512, 0, st>>>(dlogits, h, bg, N, k_pad, num_all, alpha, dbg);
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"
