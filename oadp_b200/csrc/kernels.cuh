// Internal launcher declarations shared by the OAKE translation units.
#pragma once

#include "../../include/oake_b200.h"
#include "common.cuh"

namespace oake {

// Raises a kernel's dynamic shared-memory limit once per (kernel, device): the attribute belongs to
// the device's context, so a process that drives several GPUs needs it on each of them.
template <auto Kernel>
inline cudaError_t ensure_dynamic_smem(int bytes) {
  static unsigned long long done = 0;  // one bit per device ordinal
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && ((done >> dev) & 1ull)) return cudaSuccess;
  e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev < 64) done |= 1ull << dev;
  return e;
}

// ---------------------------------------------------------------- gemm.cu
// Row statistics travel as kStatSlots partial (sum, sum of squares) pairs per row: the producing
// GEMM writes one slot per 128 output columns (N <= 1024), the consumer adds the slots in a fixed
// order -- deterministic, no atomics (unused slots must hold zeros).
constexpr int kStatSlots = 8;

struct GemmEpilogue {
  const float* bias;        // [N] fp32 (plain bias, or the folded c_n; required with colsum) or nullptr
  const float* colsum;      // [N] fp32 s_n = sum_k W'[n,k]: enables the LayerNorm fold, or nullptr
  const float2* ln_stats;   // [M][kStatSlots] statistics of the A rows over K; required with colsum
  const act_t* residual;    // act_t [M, ld_res] added after the activation (act_t output only)
  float2* out_stats;        // [M][kStatSlots] statistics of the output rows (needs residual, N % 256 == 0), or nullptr
  void* out;                // act_t or fp32 [M, ldo]; may alias `residual`
  int ldo;
  int ld_res;
  int out_f32;  // 1: fp32 output (bias / activation only), 0: act_t output
  int act;      // 0: identity, 1: QuickGELU  u * sigmoid(1.702 u)
  // fp32 output only (cosine classifier): y = alpha * y - shift; columns [ninf_lo, ninf_hi) = -inf
  float alpha = 1.f;
  float shift = 0.f;
  int ninf_lo = 0;
  int ninf_hi = 0;
  // Shifted-A mode (patch embedding of the objects tower over the block matrix, frontend.cu): K is walked in
  // segments of a_seg_kb 64-wide k-blocks; segment s reads A columns [0, 64 a_seg_kb) again, at rows + a_shift[s]
  // (rows past the end of A read as zeros).  0 = off.
  int a_seg_kb = 0;
  int a_shift[4] = {0, 0, 0, 0};
};

// Encodes a 2D row-major [rows, cols] act_t tensor as a TMA map with a (box_rows x 64) box and
// 128-byte swizzle.  Returns 0 on success.
int make_tmap_act_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                     uint32_t box_rows);

// out[M,N] = epilogue(A[M,K] * W[N,K]^T).  A and W are act_t, K-contiguous.  N % 128 == 0,
// K % 64 == 0.  tmA must have a 128-row box, tmW a `gemm_block_n(N)`-row box.
int gemm_block_n(int N);
cudaError_t launch_gemm(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmW, int M,
                        int N, int K, const GemmEpilogue& ep, int num_sms);
// CUDA-core reference of the same contract (debug / unit tests only; never on the product path).
cudaError_t launch_gemm_simt(cudaStream_t st, const act_t* A, const act_t* W, int M, int N, int K,
                             const GemmEpilogue& ep);

// -------------------------------------------------------------- rowops.cu
// LayerNorm of act_t rows (only ln_post uses a stand-alone LayerNorm; ln_1 / ln_2 are folded into
// the consuming GEMM's epilogue).
cudaError_t launch_layernorm(cudaStream_t st, const act_t* x, const float* w, const float* b,
                             act_t* out, int rows, int width);
// Builds the residual stream x (act_t) and its row statistics ([rows][kStatSlots], slot 0 = sum /
// sum of squares of the stored values, other slots zero): rows [0, B*P) = LN(patch_out + pos[1 + i % P]); rows [B*P, B*P+B) = LN(class_emb +
// pos[0]); if with_y, rows [B*P+B, B*P+2B) = copy of the class rows.  src_grid > 0: patch (gy, gx) of crop b is row
// b * src_grid^2 + gy * src_grid + gx of patch_out (the block-matrix GEMM's output grid) instead of row b * P + i.
cudaError_t launch_assemble_ln_pre(cudaStream_t st, const float* patch_out, const float* class_emb,
                                   const float* pos, const float* w, const float* b, act_t* x,
                                   float2* stats, int B, int P, int width, int with_y, int src_grid = 0);
cudaError_t launch_l2norm_half(cudaStream_t st, const float* e, __half* out, int rows, int dim);

// ----------------------------------------------------------- attention.cu
// qkv: act [R, 3*W] (q | k | v, head h = columns 64h..64h+63 of each third); rows ordered
// [B*P patch rows | B class rows | (B side rows)].  Writes act [R, W].
//   with_side = 0: main stream only (patch + class rows).
//   with_side = 1: objects; the side token y (one query over the P patch keys + itself, additive
//                  bias -100 * mask, mask fp32 [B, P] with 1 = background) rides in the same tile.
//                  side_only = 1 writes just the B side rows (last block of the objects tower).
cudaError_t launch_attention(cudaStream_t st, const act_t* qkv, const float* mask, act_t* out, int B, int P,
                             int heads, int with_side, int side_only);
// attention_cs.cu: persistent tcgen05 implementation of the same contract for the 197-token tower
// (whole tiles; default there).  OAKE_ATTN=mma selects the mma.sync kernel of attention.cu instead.
bool attention_use_tc(int P, int side_only);
cudaError_t launch_attention_cs(cudaStream_t st, const act_t* qkv, const float* mask, act_t* out, int B, int P,
                                int heads, int with_side, int side_only);

// ------------------------------------------------------------ frontend.cu
// pixels fp32 NCHW [B,3,224,224] (already CLIP-normalised) -> act [B*P, 3*32*32] conv1 patches,
// column order (c, ky, kx).  stride 32 / pad 0 (P=49) or stride 16 / pad 15 (P=196).
cudaError_t launch_im2col_pixels(cudaStream_t st, const float* pixels, act_t* patches, int B,
                                 int stride, int pad, int grid);
// Same patch matrix from uint8 HWC 224x224 crops living anywhere in `arena`; lut[c*256+v] is the
// ToTensor+Normalize value of byte v in channel c, already rounded to act_t.
cudaError_t launch_im2col_u8(cudaStream_t st, const uint8_t* arena, const oake_crop_src* crops,
                             const act_t* lut, act_t* patches, int B, int stride, int pad, int grid);
// Objects tower: the 15 x 15 block matrix [B*225, 768] (16 x 16 blocks of the zero-padded crop, column order
// (c, ky, kx)) that replaces im2col there, and conv1's weight regrouped to [O, (dy, dx), c, ky, kx] for it.
constexpr int kBlockGrid = 15;
cudaError_t launch_blockcol_pixels(cudaStream_t st, const float* pixels, act_t* blocks, int B);
cudaError_t launch_blockcol_u8(cudaStream_t st, const uint8_t* arena, const oake_crop_src* crops, const act_t* lut,
                               act_t* blocks, int B);
cudaError_t launch_conv1_regroup(cudaStream_t st, const act_t* w, act_t* out, int out_ch);
// Pillow-exact crop + antialiased bicubic resize, one job per (source rectangle -> output window).
cudaError_t launch_resize_u8(cudaStream_t st, const uint8_t* src, uint8_t* dst, const oake_resize_job* jobs,
                             int n_jobs, int max_tiles, int* err_flag);
// The same resize with the pixels written straight into the tower's front-end matrix (act_t, through the
// ToTensor + Normalize table): job i = crop i, windows must be whole 224 x 224 crops.  blocks16 = 0: conv1's im2col
// [n*49, 3072] (T50); 1: the block matrix [n*225, 768] of the objects tower, zero border included.
cudaError_t launch_resize_to_matrix(cudaStream_t st, const uint8_t* src, const oake_resize_job* jobs, int n_jobs,
                                    int* err_flag, const act_t* lut, act_t* matrix, int blocks16);
cudaError_t launch_object_masks(cudaStream_t st, const float* fg, const float* box, float* masks, int B,
                                int grid);
int resize_max_taps();

}  // namespace oake
