// Distillation-side losses, value AND gradient in one pass:
//   * L1 / MSE between the hooked `fc_cls._linear` rows and the cached CLIP rows (todd L1Loss /
//     MSELoss; configs/dp/models/{vild_ensemble_faster_rcnn_r50_fpn,block,global_}.py),
//   * RKD: MSE between the Gram matrices of student and teacher rows (oadp/base/losses.py:68-108),
//   * AsymmetricLoss on probabilities (oadp/base/losses.py:10-65; block / global heads).
// The reference runs each as 5-15 elementwise / reduction launches plus autograd's mirror image; here
// one call writes the scalar loss and d loss / d input (the backward pass only scales it).  Sums
// are two-stage and fixed-order (per-block partials, then one block): deterministic, no atomics.
#include <math.h>

#include "kernels.cuh"

namespace oake {
int fail_msg(const char* fmt, ...);  // encoder.cu

namespace {

constexpr int kPartials = 256;  // blocks of the first reduction stage

__device__ __forceinline__ float block_sum(float v, float* red) {  // 256 threads
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = l < 8 ? red[l] : 0.f;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

__global__ void __launch_bounds__(256) final_sum_kernel(const float* __restrict__ partial, int n, float scale,
                                                        float* __restrict__ loss) {
  __shared__ float red[8];
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) v += partial[i];
  const float t = block_sum(v, red);
  if (threadIdx.x == 0) *loss = t * scale;
}

// kind 0: |a - b|, kind 1: (a - b)^2.  grad = scale * d/da.
template <int KIND>
__global__ void __launch_bounds__(256) pair_loss_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        long long n, float scale, float* __restrict__ grad,
                                                        float* __restrict__ partial) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const float d = a[i] - b[i];
    if (KIND == 0) {
      acc += fabsf(d);
      if (grad) grad[i] = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
    } else {
      acc = fmaf(d, d, acc);
      if (grad) grad[i] = 2.f * scale * d;
    }
  }
  const float t = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) asl_kernel(const float* __restrict__ x, const uint8_t* __restrict__ y, long long n,
                                                  float gamma_neg, float gamma_pos, float clip, float eps, float scale,
                                                  float* __restrict__ grad, float* __restrict__ partial) {
  __shared__ float red[8];
  const bool focus = gamma_neg > 0.f || gamma_pos > 0.f;
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const float xi = x[i];
    const bool pos = y[i] != 0;
    float comp = 1.f - xi;
    bool comp_passes = true;  // gradient through clamp(max=1) (inclusive, like torch)
    if (clip > 0.f) {
      comp += clip;
      comp_passes = comp <= 1.f;
      comp = fminf(comp, 1.f);
    }
    const float p = pos ? xi : comp;                  // the probability of the true outcome
    const float lg = logf(fmaxf(p, eps));
    const float w = focus ? powf(1.f - p, pos ? gamma_pos : gamma_neg) : 1.f;  // no gradient (torch.no_grad)
    acc -= lg * w;
    if (grad) {
      float g = 0.f;
      if (p >= eps) g = pos ? -w / p : (comp_passes ? w / p : 0.f);  // d(-log p)/dx, dp/dx = +1 / -1
      grad[i] = g * scale;
    }
  }
  const float t = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// D[i][j] = s_i . s_j - t_i . t_j, one warp per (i, j); partial sums of D^2 per block of 8 pairs.
__global__ void __launch_bounds__(256) gram_diff_kernel(const float* __restrict__ s, const float* __restrict__ t, int N,
                                                        int dim, float* __restrict__ D) {
  const long long pair = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (pair >= static_cast<long long>(N) * N) return;
  const int i = static_cast<int>(pair / N), j = static_cast<int>(pair - static_cast<long long>(i) * N);
  const int lane = threadIdx.x & 31;
  float a = 0.f, b = 0.f;
  for (int c = lane; c < dim; c += 32) {
    a = fmaf(s[static_cast<size_t>(i) * dim + c], s[static_cast<size_t>(j) * dim + c], a);
    b = fmaf(t[static_cast<size_t>(i) * dim + c], t[static_cast<size_t>(j) * dim + c], b);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) D[pair] = a - b;
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ v, long long n, float* __restrict__ partial) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += 256ll * gridDim.x)
    acc = fmaf(v[i], v[i], acc);
  const float t = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// grad_s[i][c] = 4 * scale * sum_j D[i][j] s[j][c]   (D is symmetric)
__global__ void __launch_bounds__(256) rkd_grad_kernel(const float* __restrict__ D, const float* __restrict__ s, int N,
                                                       int dim, float scale4, float* __restrict__ grad) {
  const int i = blockIdx.x;
  for (int c = threadIdx.x; c < dim; c += 256) {
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(D[static_cast<size_t>(i) * N + j], s[static_cast<size_t>(j) * dim + c], acc);
    grad[static_cast<size_t>(i) * dim + c] = acc * scale4;
  }
}

int blocks_for(long long n) {
  const long long b = (n + 1023) / 1024;
  return static_cast<int>(b < 1 ? 1 : (b > kPartials ? kPartials : b));
}

}  // namespace
}  // namespace oake

using namespace oake;

extern "C" int oake_loss_workspace_bytes(int rkd_rows, size_t* out_bytes) {
  if (!out_bytes || rkd_rows < 0) return fail_msg("bad argument");
  *out_bytes = kPartials * sizeof(float) + static_cast<size_t>(rkd_rows) * rkd_rows * sizeof(float);
  return 0;
}

extern "C" int oake_pair_loss(const float* pred, const float* target, long long n, int kind, float scale, float* loss,
                              float* grad, void* ws, size_t ws_bytes, void* stream) {
  if (!pred || !target || !loss || !ws) return fail_msg("NULL buffer");
  if (n <= 0) return fail_msg("n must be positive");
  if (kind != 0 && kind != 1) return fail_msg("kind must be 0 (L1) or 1 (MSE)");
  if (ws_bytes < kPartials * sizeof(float)) return fail_msg("workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(ws);
  const int nb = blocks_for(n);
  if (kind == 0)
    pair_loss_kernel<0><<<nb, 256, 0, st>>>(pred, target, n, scale, grad, partial);
  else
    pair_loss_kernel<1><<<nb, 256, 0, st>>>(pred, target, n, scale, grad, partial);
  final_sum_kernel<<<1, 256, 0, st>>>(partial, nb, scale, loss);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("pair_loss launch: %s", cudaGetErrorString(e));
}

extern "C" int oake_asymmetric_loss(const float* x, const uint8_t* y, long long n, float gamma_neg, float gamma_pos,
                                    float clip, float eps, float scale, float* loss, float* grad, void* ws,
                                    size_t ws_bytes, void* stream) {
  if (!x || !y || !loss || !ws) return fail_msg("NULL buffer");
  if (n <= 0) return fail_msg("n must be positive");
  if (ws_bytes < kPartials * sizeof(float)) return fail_msg("workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(ws);
  const int nb = blocks_for(n);
  asl_kernel<<<nb, 256, 0, st>>>(x, y, n, gamma_neg, gamma_pos, clip, eps, scale, grad, partial);
  final_sum_kernel<<<1, 256, 0, st>>>(partial, nb, scale, loss);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("asymmetric_loss launch: %s", cudaGetErrorString(e));
}

extern "C" int oake_rkd_loss(const float* s, const float* t, int N, int dim, float scale, float* loss, float* grad_s,
                             void* ws, size_t ws_bytes, void* stream) {
  if (!s || !t || !loss || !ws) return fail_msg("NULL buffer");
  if (N <= 0 || dim <= 0) return fail_msg("N and dim must be positive");
  const size_t need = kPartials * sizeof(float) + static_cast<size_t>(N) * N * sizeof(float);
  if (ws_bytes < need) return fail_msg("workspace too small: %zu < %zu", ws_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(ws);
  float* D = partial + kPartials;
  const long long pairs = static_cast<long long>(N) * N;
  gram_diff_kernel<<<static_cast<unsigned>((pairs + 7) / 8), 256, 0, st>>>(s, t, N, dim, D);
  const int nb = blocks_for(pairs);
  sumsq_kernel<<<nb, 256, 0, st>>>(D, pairs, partial);
  final_sum_kernel<<<1, 256, 0, st>>>(partial, nb, scale, loss);
  if (grad_s) rkd_grad_kernel<<<N, 256, 0, st>>>(D, s, N, dim, 4.f * scale, grad_s);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("rkd_loss launch: %s", cudaGetErrorString(e));
}
