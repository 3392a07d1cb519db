// Re-softmax + multiclass NMS -- the tail of the ViLD / OADP inference path (second half):
//
//   mmdet BBoxHead.get_bboxes:  scores = softmax(cls_score)            (cls_score = the ensemble's log scores,
//                               multiclass_nms(bboxes, scores, score_thr, nms, max_per_img)   roi_heads.py:93-112)
//   oadp/dp/test_nni.py:55-92:  the same multiclass_nms on the re-weighted ensemble scores
//   configs: score_thr 0.0, nms iou 0.5, max_per_img 300 (vild_ensemble_faster_rcnn_r50_fpn.py:41-44)
//
// mmdet expands every RoI into one candidate per foreground class ((N x K) candidates, K = 65 / 1203), offsets
// the boxes class by class and runs one NMS over the lot (or K separate ones beyond 10 000 candidates).  With
// class-agnostic box regression (`reg_class_agnostic=True`, the reference's configs) all K problems share the
// SAME N boxes, so the N x N overlap relation is computed ONCE, as a bit matrix; each class then sorts its
// score column (one CTA per class, bitonic sort in shared memory) and walks it greedily against that matrix.
//   nms_iou_mask_kernel   mask[i][w] bit b = IoU(box i, box 64 w + b) > thr     (N x ceil(N/64) words)
//   nms_classes_kernel    per class: candidates with score > score_thr, descending scores, greedy keep
// The suppression rule and the IoU expression are torchvision / mmcv `nms` (inter / (area_i + area_j - inter) >
// thr, areas (x2 - x1) (y2 - y1), fp32).  The final top-`max_num` over the kept candidates is a library top-k
// in the Python wrapper (oadp_b200/dp/nms.py).
#include <math.h>

#include "kernels.cuh"

namespace oake {
int fail_msg(const char* fmt, ...);  // encoder.cu

namespace {

constexpr int kMaxBoxes = 4096;  // per image; the reference keeps 1000 proposals (faster_rcnn_r50_fpn.py:122-127)

// softmax over K1 columns of a row, one warp per row (K1 <= 1280 like the ensemble kernel)
constexpr int kMaxPerLane = 40;
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int N,
                                                           int K1, int ld_in, int ld_out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* src = in + static_cast<size_t>(row) * ld_in;
  float v[kMaxPerLane];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < K1 ? __ldg(src + c) : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    v[i] = expf(v[i] - m);
    s += v[i];
  }
  const float inv = 1.0f / warp_sum(s);
  float* dst = out + static_cast<size_t>(row) * ld_out;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    if (c < K1) dst[c] = v[i] * inv;
  }
}

__device__ __forceinline__ bool overlaps(const float4 a, const float4 b, float thr) {
  const float iw = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.f);
  const float ih = fmaxf(fminf(a.w, b.w) - fmaxf(a.y, b.y), 0.f);
  const float inter = iw * ih;
  const float area_a = (a.z - a.x) * (a.w - a.y), area_b = (b.z - b.x) * (b.w - b.y);
  return inter / (area_a + area_b - inter) > thr;
}

// grid (ceil(N/64), ceil(N/64)), 64 threads: thread t of block (bx, by) owns box i = 64 by + t and the 64 boxes
// of column block bx, staged in shared memory
__global__ void __launch_bounds__(64) nms_iou_mask_kernel(const float4* __restrict__ boxes, int N, float thr,
                                                          unsigned long long* __restrict__ mask, int words) {
  __shared__ float4 col[64];
  const int t = threadIdx.x;
  const int j0 = blockIdx.x * 64, i = blockIdx.y * 64 + t;
  if (j0 + t < N) col[t] = boxes[j0 + t];
  __syncthreads();
  if (i >= N) return;
  const float4 a = boxes[i];
  unsigned long long bits = 0ull;
  const int nj = min(64, N - j0);
  for (int b = 0; b < nj; ++b)
    if (overlaps(a, col[b], thr)) bits |= 1ull << b;
  mask[static_cast<size_t>(i) * words + blockIdx.x] = bits;
}

// One CTA (256 threads) per class.  scores: column `cls` of a row-major (N, ld) matrix.
__global__ void __launch_bounds__(256) nms_classes_kernel(const float* __restrict__ scores, int N, int ld, float score_thr,
                                                          const unsigned long long* __restrict__ mask, int words,
                                                          uint8_t* __restrict__ keep, int n_pad) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* key = reinterpret_cast<float*>(smem);               // [n_pad] scores, -inf for non-candidates / padding
  int* idx = reinterpret_cast<int*>(key + n_pad);            // [n_pad] box index
  unsigned long long* removed = reinterpret_cast<unsigned long long*>(idx + n_pad);  // [words]
  const int cls = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < n_pad; i += blockDim.x) {
    float s = -INFINITY;
    if (i < N) {
      s = scores[static_cast<size_t>(i) * ld + cls];
      if (!(s > score_thr)) s = -INFINITY;  // mmdet: valid_mask = scores > score_thr (NaN is not a candidate)
    }
    key[i] = s;
    idx[i] = i;
  }
  for (int w = tid; w < words; w += blockDim.x) removed[w] = 0ull;
  __syncthreads();
  // bitonic sort, descending by score; ties by ascending box index (a total order: the result does not depend
  // on the thread schedule)
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n_pad; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const float a = key[i], b = key[p];
          const int ia = idx[i], ib = idx[p];
          const bool a_first = a > b || (a == b && ia < ib);  // a belongs before b in descending order
          const bool up = (i & k) == 0;
          if (up ? !a_first : a_first) {
            key[i] = b;
            key[p] = a;
            idx[i] = ib;
            idx[p] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  // greedy scan (one warp: the scan is sequential; each lane owns words lane, lane + 32, ... of `removed`)
  uint8_t* out = keep + static_cast<size_t>(cls) * N;
  for (int i = tid; i < N; i += blockDim.x) out[i] = 0;
  __syncthreads();
  if (tid < 32) {
    for (int r = 0; r < N; ++r) {
      const float s = key[r];
      if (s == -INFINITY) break;  // the candidates are exhausted (sorted: the rest is not valid either)
      const int i = idx[r];
      const bool dead = (removed[i >> 6] >> (i & 63)) & 1ull;
      if (!dead) {  // warp-uniform: every lane reads the same word
        if (tid == 0) out[i] = 1;
        const unsigned long long* row = mask + static_cast<size_t>(i) * words;
        for (int w = tid; w < words; w += 32) removed[w] |= row[w];
      }
      __syncwarp();
    }
  }
}

}  // namespace
}  // namespace oake

using namespace oake;

extern "C" {

int oake_softmax_rows(const float* in, int N, int K1, int ld_in, float* out, int ld_out, void* stream) {
  if (N == 0) return 0;
  if (!in || !out) return fail_msg("NULL buffer");
  if (K1 < 1 || K1 > 32 * kMaxPerLane) return fail_msg("K + 1 = %d outside [1, %d]", K1, 32 * kMaxPerLane);
  if (ld_in < K1 || ld_out < K1) return fail_msg("row pitch smaller than K + 1");
  softmax_rows_kernel<<<(N + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, N, K1, ld_in, ld_out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("softmax_rows launch: %s", cudaGetErrorString(e));
}

int oake_nms_workspace_bytes(int N, size_t* out_bytes) {
  if (!out_bytes || N < 0 || N > kMaxBoxes) return fail_msg("N must be in [0, %d]", kMaxBoxes);
  const size_t words = (static_cast<size_t>(N) + 63) / 64;
  *out_bytes = static_cast<size_t>(N) * words * 8 + 256;
  return 0;
}

int oake_multiclass_nms(const float* boxes_xyxy, const float* scores, int N, int K, int ld_scores, float score_thr,
                        float iou_thr, uint8_t* keep, void* ws, size_t ws_bytes, void* stream) {
  if (N == 0 || K == 0) return 0;
  if (!boxes_xyxy || !scores || !keep || !ws) return fail_msg("NULL buffer");
  if (N < 0 || N > kMaxBoxes) return fail_msg("N = %d boxes per call, at most %d", N, kMaxBoxes);
  if (K < 0 || ld_scores < K) return fail_msg("bad K / row pitch");
  if ((reinterpret_cast<uintptr_t>(boxes_xyxy) & 15) != 0) return fail_msg("boxes must be 16-byte aligned");
  const int words = (N + 63) / 64;
  size_t need = 0;
  oake_nms_workspace_bytes(N, &need);
  if (ws_bytes < need) return fail_msg("workspace too small: need %zu bytes", need);
  unsigned long long* mask =
      reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  nms_iou_mask_kernel<<<dim3(words, words), 64, 0, st>>>(reinterpret_cast<const float4*>(boxes_xyxy), N, iou_thr, mask, words);
  int n_pad = 64;
  while (n_pad < N) n_pad <<= 1;
  const size_t smem = static_cast<size_t>(n_pad) * 8 + static_cast<size_t>(words) * 8;
  if (cudaError_t e = ensure_dynamic_smem<nms_classes_kernel>(static_cast<int>(4096 * 8 + 64 * 8)); e != cudaSuccess)
    return fail_msg("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  nms_classes_kernel<<<K, 256, smem, st>>>(scores, N, ld_scores, score_thr, mask, words, keep, n_pad);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("multiclass_nms launch: %s", cudaGetErrorString(e));
}

}  // extern "C"
