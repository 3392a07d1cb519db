// ViLD ensemble scoring: the epilogue that sits directly behind the two cosine
// classifier calls at inference, oadp/dp/roi_heads.py:93-112:
//
//     bbox_scores   = softmax(bbox_logits)   ** lambda
//     object_scores = softmax(object_logits) ** (1 - lambda)
//     cls_score     = bbox_scores * object_scores;  cls_score[:, -1] = 1 - cls_score[:, :-1].sum(-1)
//     return cls_score.log()
//
// lambda = 2/3 for base categories, 1/3 for novel categories and the background column
// (roi_heads.py:55-59).  The reference runs ~10 elementwise / reduction kernels over (N, K+1) fp32
// (N ~ 1000 RoIs per image, K+1 = 66 or 1204); here one warp owns one RoI row, both logit rows are
// read once into registers, and the log-score row is written once: HBM-bound, 12 bytes per element.
#include <math.h>

#include "kernels.cuh"

namespace oake {
int fail_msg(const char* fmt, ...);  // encoder.cu

namespace {

constexpr int kMaxPerLane = 40;  // K + 1 <= 1280

__global__ void __launch_bounds__(256) vild_ensemble_kernel(const float* __restrict__ bbox_logits,
                                                            const float* __restrict__ object_logits,
                                                            const float* __restrict__ lambda, float* __restrict__ out,
                                                            int N, int K1, int ld_bbox, int ld_object, int ld_out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* b = bbox_logits + static_cast<size_t>(row) * ld_bbox;
  const float* o = object_logits + static_cast<size_t>(row) * ld_object;
  float vb[kMaxPerLane], vo[kMaxPerLane];
  float mb = -INFINITY, mo = -INFINITY;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    vb[i] = c < K1 ? __ldg(b + c) : -INFINITY;
    vo[i] = c < K1 ? __ldg(o + c) : -INFINITY;
    mb = fmaxf(mb, vb[i]);
    mo = fmaxf(mo, vo[i]);
  }
  mb = warp_max(mb);
  mo = warp_max(mo);
  float sb = 0.f, so = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    vb[i] -= mb;  // -inf (masked column, padding) stays -inf
    vo[i] -= mo;
    sb += expf(vb[i]);
    so += expf(vo[i]);
  }
  // log-domain: log(softmax(b)^l * softmax(o)^(1-l)) = l (b - mb - log sb) + (1-l) (o - mo - log so);
  // one exponential per element is left, for the foreground sum the background column needs
  const float lb = logf(warp_sum(sb));
  const float lo = logf(warp_sum(so));
  float fg = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    float ls = -INFINITY;
    if (c < K1) {
      const float l = __ldg(lambda + c);
      ls = l * (vb[i] - lb) + (1.0f - l) * (vo[i] - lo);  // lambda in (0, 1): -inf terms give -inf, never NaN
    }
    vb[i] = ls;
    if (c < K1 - 1) fg += expf(ls);
  }
  fg = warp_sum(fg);
  float* dst = out + static_cast<size_t>(row) * ld_out;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    if (c < K1) dst[c] = c == K1 - 1 ? logf(1.0f - fg) : vb[i];
  }
}

}  // namespace
}  // namespace oake

using namespace oake;

extern "C" int oake_vild_ensemble(const float* bbox_logits, const float* object_logits, const float* lambda,
                                  int N, int K1, int ld_bbox, int ld_object, float* out, int ld_out, void* stream) {
  if (N == 0) return 0;
  if (!bbox_logits || !object_logits || !lambda || !out) return fail_msg("NULL buffer");
  if (K1 < 2 || K1 > 32 * kMaxPerLane) return fail_msg("K + 1 = %d outside [2, %d]", K1, 32 * kMaxPerLane);
  if (ld_bbox < K1 || ld_object < K1 || ld_out < K1) return fail_msg("row pitch smaller than K + 1");
  const int rows_per_block = 8;
  vild_ensemble_kernel<<<(N + rows_per_block - 1) / rows_per_block, 32 * rows_per_block, 0,
                         static_cast<cudaStream_t>(stream)>>>(bbox_logits, object_logits, lambda, out, N, K1,
                                                              ld_bbox, ld_object, ld_out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("vild_ensemble launch: %s", cudaGetErrorString(e));
}
