// Front end of the OAKE tower: turns images + crop rectangles into the conv1
// patch matrix the tcgen05 GEMM consumes, entirely on the GPU.
//
//  * resize_u8: Pillow's antialiased separable bicubic (`Image.crop(box).resize(size, BICUBIC)`,
//    libImaging/Resample.c: double-precision coefficients, 22-bit fixed point, horizontal pass then
//    vertical pass with a uint8 intermediate), bit-exact on uint8.  It stands where the reference's
//    DataLoader workers call PIL through torchvision `Resize(224, BICUBIC)` + `CenterCrop(224)`
//    (clip `_transform`; oadp/oake/globals.py:32, blocks.py:73-81,95, objects.py:116-127), and the
//    pyramid `image.resize((int(w / 1.5), int(h / 1.5)))` of blocks.py:73-77.
//  * im2col_u8: uint8 224x224 crops (anywhere in a byte arena) -> ToTensor + Normalize (exact fp32
//    table) -> act_t [B*P, 3072] patch rows, column order (c, ky, kx).  stride 32 / pad 0 (P = 49)
//    is CLIP's conv1; stride 16 / pad 15 (P = 196) is the objects surgery of objects.py:299-301.
//  * im2col_pixels: the same patch matrix from (B,3,224,224) fp32 crops that were preprocessed on
//    the host, i.e. the tensor the reference hands to `encode_image` / `visual`.
//  * object_masks: the 14x14 foreground masks of objects.py:129-155.
#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "../../include/oake_b200.h"
#include "kernels.cuh"

namespace oake {

namespace {

constexpr int kImg = 224;
constexpr int kPatch = 32;
constexpr int kCols = 3 * kPatch * kPatch;  // 3072
constexpr int kChunks = kCols / 8;          // 16-byte output chunks per patch row

// ------------------------------------------------------------------------------------------------
// im2col from fp32 NCHW crops
// ------------------------------------------------------------------------------------------------
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

// ------------------------------------------------------------------------------------------------
// im2col from uint8 HWC crops + ToTensor/Normalize table
// ------------------------------------------------------------------------------------------------
// One thread = 8 consecutive kx of one (crop, patch, ky) for all three channels: reads 24
// contiguous bytes, writes three 16-byte chunks.  lut[c * 256 + v] = act_t((v / 255 - mean_c) / std_c);
// padded (out-of-crop) pixels are exact zeros, as conv2d's zero padding acts AFTER Normalize.
__global__ void __launch_bounds__(256)
im2col_u8_kernel(const uint8_t* __restrict__ arena, const oake_crop_src* __restrict__ crops,
                 const act_t* __restrict__ lut, act_t* __restrict__ patches, long long total, int stride,
                 int pad, int grid) {
  __shared__ act_t s_lut[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int per_row = kPatch * kPatch / 8;  // (ky, kx0) pairs per patch
  const int sub = static_cast<int>(idx % per_row);
  const long long prow = idx / per_row;
  const int P = grid * grid;
  const int b = static_cast<int>(prow / P);
  const int g = static_cast<int>(prow - static_cast<long long>(b) * P);
  const int gy = g / grid, gx = g - gy * grid;
  const int ky = sub / (kPatch / 8);
  const int kx0 = (sub - ky * (kPatch / 8)) * 8;
  const int y = gy * stride - pad + ky;
  const int x0 = gx * stride - pad + kx0;
  const oake_crop_src cs = crops[b];
  uint32_t out[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[c][j] = 0u;
  if (y >= 0 && y < kImg) {
    const uint8_t* src = arena + cs.off + (static_cast<long long>(y) * cs.pitch_px + x0) * 3;
    unsigned short h[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int x = x0 + j;
      const bool ok = x >= 0 && x < kImg;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        act_t val = to_act(0.f);
        if (ok) val = s_lut[c * 256 + src[j * 3 + c]];
        h[c][j] = *reinterpret_cast<unsigned short*>(&val);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        out[c][j] = static_cast<uint32_t>(h[c][2 * j]) | (static_cast<uint32_t>(h[c][2 * j + 1]) << 16);
  }
  act_t* dst = patches + prow * kCols + ky * kPatch + kx0;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    *reinterpret_cast<uint4*>(dst + c * kPatch * kPatch) = make_uint4(out[c][0], out[c][1], out[c][2], out[c][3]);
}

// ------------------------------------------------------------------------------------------------
// Objects tower (stride 16, pad 15, 32 x 32 kernel; objects.py:299-301): block matrix instead of im2col.
// The padded 254 x 254 crop is cut into 15 x 15 NON-overlapping 16 x 16 blocks (rows 0..239: the last 14 padded
// rows / columns are never touched by a patch); row (b, by, bx) of the block matrix holds one block, column
// order (c, ky, kx), 768 wide.  A 32 x 32 patch at (gy, gx) is the four blocks (gy + dy, gx + dx), i.e. rows
// r, r + 1, r + 15, r + 16 of the matrix, so the patch embedding is ONE GEMM whose A operand is read at four row
// shifts (GemmEpilogue::a_seg_kb) against conv1's weight regrouped by (dy, dx) -- 0.35 MB per crop written and
// re-read (from L2) instead of the 1.2 MB of an explicit im2col in which every pixel appears four times.
// ------------------------------------------------------------------------------------------------
constexpr int kBlk = 16;                       // block edge
constexpr int kBlkGrid = 15;                   // blocks per crop edge
constexpr int kBlkCols = 3 * kBlk * kBlk;      // 768
constexpr int kBlkPad = 15;

// One thread = 8 consecutive kx of one (crop, block, ky), three channels: 24 contiguous source bytes, three
// 16-byte chunks out; a warp writes one whole block (3 x 512 contiguous bytes).
__global__ void __launch_bounds__(256)
blockcol_u8_kernel(const uint8_t* __restrict__ arena, const oake_crop_src* __restrict__ crops,
                   const act_t* __restrict__ lut, act_t* __restrict__ blocks, long long total) {
  __shared__ act_t s_lut[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int sub = static_cast<int>(idx & 31);
  const long long brow = idx >> 5;  // b * 225 + by * 15 + bx
  const int b = static_cast<int>(brow / (kBlkGrid * kBlkGrid));
  const int g = static_cast<int>(brow - static_cast<long long>(b) * (kBlkGrid * kBlkGrid));
  const int by = g / kBlkGrid, bx = g - by * kBlkGrid;
  const int ky = sub >> 1, kx0 = (sub & 1) * 8;
  const int y = by * kBlk + ky - kBlkPad;
  const int x0 = bx * kBlk + kx0 - kBlkPad;
  uint32_t out[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[c][j] = 0u;
  if (y >= 0 && y < kImg) {
    const oake_crop_src cs = crops[b];
    const uint8_t* src = arena + cs.off + (static_cast<long long>(y) * cs.pitch_px + x0) * 3;
    unsigned short h[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int x = x0 + j;
      const bool ok = x >= 0 && x < kImg;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        act_t val = to_act(0.f);
        if (ok) val = s_lut[c * 256 + src[j * 3 + c]];
        h[c][j] = *reinterpret_cast<unsigned short*>(&val);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        out[c][j] = static_cast<uint32_t>(h[c][2 * j]) | (static_cast<uint32_t>(h[c][2 * j + 1]) << 16);
  }
  act_t* dst = blocks + brow * kBlkCols + ky * kBlk + kx0;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    *reinterpret_cast<uint4*>(dst + c * kBlk * kBlk) = make_uint4(out[c][0], out[c][1], out[c][2], out[c][3]);
}

// The same block matrix from fp32 NCHW crops; one thread = one 16-byte output chunk.
__global__ void __launch_bounds__(256)
blockcol_pixels_kernel(const float* __restrict__ pixels, act_t* __restrict__ blocks, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int chunk = static_cast<int>(idx % (kBlkCols / 8));  // (c, ky, kx0)
  const long long brow = idx / (kBlkCols / 8);
  const int b = static_cast<int>(brow / (kBlkGrid * kBlkGrid));
  const int g = static_cast<int>(brow - static_cast<long long>(b) * (kBlkGrid * kBlkGrid));
  const int by = g / kBlkGrid, bx = g - by * kBlkGrid;
  const int c = chunk >> 5, ky = (chunk >> 1) & 15, kx0 = (chunk & 1) * 8;
  const int y = by * kBlk + ky - kBlkPad;
  const int x0 = bx * kBlk + kx0 - kBlkPad;
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
  *reinterpret_cast<uint4*>(blocks + brow * kBlkCols + chunk * 8) = u;
}

// conv1 weight [O, 3, 32, 32] -> [O, (dy, dx), 3, 16, 16]: the K order the shifted-A patch GEMM walks.
__global__ void conv1_regroup_kernel(const act_t* __restrict__ w, act_t* __restrict__ out, int total) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int o = idx / kCols, k = idx - o * kCols;
  const int s = k / kBlkCols, r = k - s * kBlkCols;
  const int c = r >> 8, ky = (r >> 4) & 15, kx = r & 15;
  const int dy = s >> 1, dx = s & 1;
  out[idx] = w[static_cast<size_t>(o) * kCols + c * kPatch * kPatch + (dy * kBlk + ky) * kPatch + dx * kBlk + kx];
}

// ------------------------------------------------------------------------------------------------
// Pillow-exact antialiased bicubic resize of a crop rectangle, uint8 HWC
// ------------------------------------------------------------------------------------------------
constexpr int kTile = 32;       // output tile (pixels)
constexpr int kMaxTaps = 48;    // filter taps per output sample: scale factors up to ~11.5
constexpr int kMaxRows = 448;   // intermediate rows per tile kept in shared memory
constexpr int kPrecBits = 32 - 8 - 2;

// libImaging bicubic_filter, a = -0.5, evaluated with explicit roundings (no FMA contraction).
__device__ __forceinline__ double pil_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) {
    const double t = __dadd_rn(__dmul_rn(a + 2.0, x), -(a + 3.0));
    return __dadd_rn(__dmul_rn(__dmul_rn(t, x), x), 1.0);
  }
  if (x < 2.0) {
    double t = __dmul_rn(__dadd_rn(x, -5.0), x);
    t = __dmul_rn(__dadd_rn(t, 8.0), x);
    return __dmul_rn(__dadd_rn(t, -4.0), a);
  }
  return 0.0;
}

// libImaging precompute_coeffs + normalize_coeffs_8bpc, split so that a whole CTA shares the work:
// pil_bounds (one thread per output sample) fixes the tap window, pil_weight (one thread per tap)
// evaluates the filter in double precision, pil_weight_sum (one thread per sample) performs the
// sequential weight sum in libImaging's order; the divide / fixed-point rounding of every tap is
// independent again (pil_fixed, one thread per tap).
struct PilAxis {
  double scale, ss, support;
};
__device__ __forceinline__ PilAxis pil_axis(int in_size, int out_size) {
  PilAxis a;
  a.scale = static_cast<double>(in_size) / static_cast<double>(out_size);
  const double filterscale = a.scale < 1.0 ? 1.0 : a.scale;
  a.support = __dmul_rn(2.0, filterscale);
  a.ss = 1.0 / filterscale;
  return a;
}
__device__ __forceinline__ void pil_bounds(const PilAxis& a, int in_size, int xx, int* first, int* count) {
  const double center = __dmul_rn(static_cast<double>(xx) + 0.5, a.scale);
  int xmin = static_cast<int>(__dadd_rn(__dadd_rn(center, -a.support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, a.support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > kMaxTaps) xmax = kMaxTaps;  // guarded by the host-side scale check
  *first = xmin;
  *count = xmax;
}
__device__ __forceinline__ double pil_weight(const PilAxis& a, int xx, int xmin, int x) {
  const double center = __dmul_rn(static_cast<double>(xx) + 0.5, a.scale);
  const double arg = __dmul_rn(__dadd_rn(__dadd_rn(static_cast<double>(x + xmin), -center), 0.5), a.ss);
  return pil_bicubic(arg);
}
__device__ __forceinline__ double pil_weight_sum(const double* w, int count) {  // libImaging's order
  double ww = 0.0;
  for (int x = 0; x < count; ++x) ww = __dadd_rn(ww, w[x]);
  return ww;
}
__device__ __forceinline__ int pil_fixed(double v, double ww) {  // normalise + 22-bit fixed point
  if (ww != 0.0) v = __ddiv_rn(v, ww);
  const double f = __dmul_rn(v, static_cast<double>(1 << kPrecBits));
  return v < 0.0 ? static_cast<int>(__dadd_rn(-0.5, f)) : static_cast<int>(__dadd_rn(0.5, f));
}

__device__ __forceinline__ uint8_t pil_clip8(int ss) {
  const int v = ss >> kPrecBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// ---- Coefficients once per JOB, pixels per tile.
// A tile used to regenerate its 64 rows of filter coefficients itself: in fp64, behind five barriers, as many
// instructions as the pixel work of an upsampled crop (and 40-tap rows for the 5x downscales that the expanded
// boxes of large proposals need, objects.py:76-99).  resize_prepare_kernel now computes the tables of both axes once
// per job (libImaging precompute_coeffs + normalize_coeffs_8bpc, the same device functions as before: same bits)
// into a stream-ordered scratch, and sorts the jobs into two classes:
//   FAST  at most kFastTaps taps per axis (scale factors up to 3.5): resize_fast_kernel, one CTA per tile, pure
//         integer pixel work in 17 KB of shared memory -- eight CTAs = 64 warps per SM instead of four (the kernel
//         is bound by the latency of its byte loads, profiles/r2_07);
//   BIG   everything else: a persistent grid (56 KB per CTA) walks the (job, tile) list the prepare kernel collected.
//         Jobs whose tables found no room in the scratch fall back to per-tile generation there.
// Both kernels compute only the part of a tile's source footprint that lies inside the image: PIL `crop` pads with
// zeros, a zero row / column of the intermediate image adds exactly 0 to every sum, and a tile that misses the
// image altogether is written as zeros at once.
#ifndef OAKE_RESIZE_HFIXED
#define OAKE_RESIZE_HFIXED 1  // 0: the horizontal pass always takes the general tap loop (A/B builds)
#endif
// (Measured and rejected on one box, ms per 2350 crops, tools/bench_resize.py: FAST up to 24 taps / 200 rows 2.16 vs
// 2.09 -- heavy tiles in a plain grid; the BIG kernel on a side stream next to the FAST one, two CTAs per SM, 2.19
// vs 2.08.)
constexpr int kFastTaps = 16;
constexpr int kFastRows = 128;             // 31 * 3.5 + 2 * 7 + 2 source rows under one 32-row tile
constexpr int kJobTabInts = 12288;         // scratch budget per job (48 KB); a 224 x 224 FAST window needs 8064 ints
constexpr int kTabSlackInts = 1 << 20;     // + 4 MB per pass: a few whole-image pyramid levels (640 + 480 samples each)

struct ResizeState {  // header of the scratch of one oake_resize_u8 pass
  int n_big;          // jobs on the BIG work list
  int big_tiles;      // sum of their tile counts
  int next;           // work counter of the persistent kernel
  int tab_used;       // ints handed out of the table arena
};
struct ResizeJobInfo {
  int cls;      // 0 = FAST, 1 = BIG
  int tab_off;  // first int of the job's tables in the arena, -1 = none.  Per axis (horizontal, then vertical), with
                // n = the window size rounded up to whole tiles: first[n], count[n], taps[n][pitch]
  int pitch;    // ints per row of taps: kFastTaps + 1 or kMaxTaps + 1
  int pad;
};

// Where the resized pixels go.  mode 0: uint8 HWC at the job's dst_off (the C ABI's oake_resize_u8).  mode 1: straight
// into the tower's front-end matrix as act_t through the ToTensor + Normalize table -- job i of the launch is crop
// crop0 + i, its window is the whole 224 x 224 crop, and pixel (y, x, c) lands at
//   row (crop * grid + (y + pad) >> shift) * grid + ((x + pad) >> shift),  column (c << 2 shift) + ky << shift + kx
// (T50: 32 x 32 patches = conv1's im2col, shift 5, pad 0, grid 7; T197: the 16 x 16 block matrix, shift 4, pad 15,
// grid 15).  Every pixel has exactly one place in either matrix, so the uint8 crop and the kernel that re-read it
// disappear.
struct ResizeOut {
  int mode;
  act_t* matrix;
  const act_t* lut;
  int shift, pad, grid;
  int crop0;
};
__device__ __forceinline__ int pil_ksize(int in_size, int out_size) {
  const PilAxis a = pil_axis(in_size, out_size);
  return static_cast<int>(ceil(a.support)) * 2 + 1;  // libImaging precompute_coeffs
}

// grid = jobs.  Classifies the job, reserves and fills its tables, appends BIG jobs to the work list.
__global__ void __launch_bounds__(256)
resize_prepare_kernel(const oake_resize_job* __restrict__ jobs, int n_jobs, ResizeState* __restrict__ state,
                      ResizeJobInfo* __restrict__ info, int4* __restrict__ big_list, int* __restrict__ tab,
                      int* __restrict__ err, ResizeOut out) {
  __shared__ ResizeJobInfo s_info;
  const int jb = blockIdx.x;
  const oake_resize_job job = jobs[jb];
  if (out.mode == 1 && (job.win_w != kImg || job.win_h != kImg)) {  // the matrix holds whole crops only
    if (threadIdx.x == 0) {
      atomicExch(err, 1);
      info[jb] = ResizeJobInfo{0, -1, 0, 1};  // (pad = 1: skipped by both tile kernels)
    }
    return;
  }
  if (threadIdx.x == 0) {
    const int kmax = max(pil_ksize(job.box_w, job.out_w), pil_ksize(job.box_h, job.out_h));
    ResizeJobInfo ji;
    ji.pitch = 1 + (kmax <= kFastTaps ? kFastTaps : kMaxTaps);
    ji.pad = 0;
    ji.tab_off = -1;
    if (kmax <= kMaxTaps) {
      const int n_h = (job.win_w + kTile - 1) / kTile * kTile, n_v = (job.win_h + kTile - 1) / kTile * kTile;
      const long long need = static_cast<long long>(n_h + n_v) * (2 + ji.pitch);
      const long long cap = n_jobs > 0 ? static_cast<long long>(n_jobs) * kJobTabInts + kTabSlackInts : 0;  // (0: test hook)
      if (need <= cap) {
        const int off = atomicAdd(&state->tab_used, static_cast<int>(need));
        if (off + need <= cap) ji.tab_off = off;
      }
    }
    ji.cls = (kmax <= kFastTaps && ji.tab_off >= 0) ? 0 : 1;
    if (ji.cls == 1) {
      const int tiles = ((job.win_w + kTile - 1) / kTile) * ((job.win_h + kTile - 1) / kTile);
      const int slot = atomicAdd(&state->n_big, 1);
      const int start = atomicAdd(&state->big_tiles, tiles);
      big_list[slot] = make_int4(jb, start, tiles, 0);
    }
    s_info = ji;
    info[jb] = ji;
  }
  __syncthreads();
  if (out.mode == 1 && out.pad > 0) {
    // the zero border of the padded crop: every (block, channel, ky) row of the matrix that holds a padding element
    // is cleared here; the tile kernels (later in the stream) then write the pixels that share those rows
    const int crop = out.crop0 + jb, g = out.grid, bs = 1 << out.shift;  // (one row = bs elements = 32 bytes at bs = 16)
    const int last = 224 + out.pad - 1;                                  // last padded coordinate that holds a pixel
    act_t* base = out.matrix + static_cast<size_t>(crop) * g * g * 3 * bs * bs;
    for (int row = threadIdx.x; row < g * g * 3 * bs; row += blockDim.x) {
      const int ky = row & (bs - 1), blk = row / (3 * bs);
      const int by = blk / g, bx = blk - by * g;
      const int py = by * bs + ky;
      const bool padded = py < out.pad || py > last || bx * bs < out.pad || bx * bs + bs - 1 > last;
      if (padded) {
        uint4* o = reinterpret_cast<uint4*>(base + static_cast<size_t>(row) * bs);
        for (int i = 0; i < bs / 8; ++i) o[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  const ResizeJobInfo ji = s_info;
  if (ji.tab_off < 0) return;
  const PilAxis ax_h = pil_axis(job.box_w, job.out_w);
  const PilAxis ax_v = pil_axis(job.box_h, job.out_h);
  const int n_h = (job.win_w + kTile - 1) / kTile * kTile, n_v = (job.win_h + kTile - 1) / kTile * kTile;
  for (int sidx = threadIdx.x; sidx < n_h + n_v; sidx += blockDim.x) {
    const bool is_h = sidx < n_h;
    const int i = is_h ? sidx : sidx - n_h, n = is_h ? n_h : n_v;
    int* axis = tab + ji.tab_off + (is_h ? 0 : n_h * (2 + ji.pitch));
    int first = 0, count = 0;
    if (i < (is_h ? job.win_w : job.win_h)) {
      const PilAxis& ax = is_h ? ax_h : ax_v;
      const int xx = (is_h ? job.win_x : job.win_y) + i;
      pil_bounds(ax, is_h ? job.box_w : job.box_h, xx, &first, &count);
      count = min(count, ji.pitch - 1);  // (count <= ksize <= taps already)
      double w[kMaxTaps];
      for (int x = 0; x < count; ++x) w[x] = pil_weight(ax, xx, first, x);
      const double ww = pil_weight_sum(w, count);
      int* k = axis + 2 * n + static_cast<size_t>(i) * ji.pitch;
      for (int x = 0; x < count; ++x) k[x] = pil_fixed(w[x], ww);
    }
    axis[i] = first;  // (samples beyond the window: empty entries, so that tiles copy whole rows)
    axis[n + i] = count;
  }
}

template <int TAPS, int ROWS, bool GEN>
struct TileSmem {
  static constexpr int kPitch = TAPS + 1;  // odd row pitch: thread = column reads are conflict-free
  alignas(16) int kh[kTile][TAPS + 1];
  alignas(16) int kv[kTile][TAPS + 1];
  alignas(16) int h_first[kTile];
  alignas(16) int h_count[kTile];
  alignas(16) int v_first[kTile];
  alignas(16) int v_count[kTile];
  int work_job, work_tile;
  alignas(16) act_t lut[768];  // ToTensor + Normalize table (matrix output only)
  double ww[GEN ? 2 * kTile : 1];
  union {
    double w[GEN ? 2 * kTile : 1][GEN ? kMaxTaps : 1];  // raw filter weights (per-tile generation: BIG jobs without a table)
    uint8_t tmp[ROWS][kTile][3];    // horizontally resampled rows (pixel phase)
  };
};
using FastSmem = TileSmem<kFastTaps, kFastRows, false>;
using BigSmem = TileSmem<kMaxTaps, kMaxRows, true>;

// The tile's 32 + 32 table rows -> shared memory, as 16-byte copies (the table has the shared arrays' layout).
// (tx, ty): tile coordinates inside the window.
template <class SM>
__device__ __forceinline__ void tile_load_tables(SM& sm, const int* __restrict__ tj, int win_w, int win_h, int tx, int ty) {
  constexpr int P = SM::kPitch;
  constexpr int kK4 = kTile * P / 4;  // 16-byte pieces of one axis' taps
  const int n_h = (win_w + kTile - 1) / kTile * kTile, n_v = (win_h + kTile - 1) / kTile * kTile;
  const int* ah = tj;
  const int* av = tj + n_h * (2 + P);
  const int4* k_h = reinterpret_cast<const int4*>(ah + 2 * n_h + tx * kTile * P);
  const int4* k_v = reinterpret_cast<const int4*>(av + 2 * n_v + ty * kTile * P);
  int4* s_kh = reinterpret_cast<int4*>(&sm.kh[0][0]);
  int4* s_kv = reinterpret_cast<int4*>(&sm.kv[0][0]);
  for (int i = threadIdx.x; i < kK4; i += blockDim.x) {
    s_kh[i] = __ldg(k_h + i);
    s_kv[i] = __ldg(k_v + i);
  }
  if (threadIdx.x < 4 * (kTile / 4)) {  // first / count of both axes: four rows of 32 ints
    const int a = threadIdx.x / (kTile / 4), o = threadIdx.x % (kTile / 4);
    const int* g = a == 0 ? ah + tx * kTile : a == 1 ? ah + n_h + tx * kTile : a == 2 ? av + ty * kTile : av + n_v + ty * kTile;
    int* d = a == 0 ? sm.h_first : a == 1 ? sm.h_count : a == 2 ? sm.v_first : sm.v_count;
    reinterpret_cast<int4*>(d)[o] = __ldg(reinterpret_cast<const int4*>(g) + o);
  }
}

// Source footprint of the tile (valid after the tile's bounds are in shared memory and a barrier):
// tmp row r holds source row y_lo + r; rows [rv0, rv1) lie inside the image.  Returns false for a tile that
// misses the image altogether.  Bounds are monotonic in the sample index (libImaging: xmin, xmax from a centre
// that grows with the sample).
struct Footprint {
  int r_lo, nr, y_lo, rv0, rv1;
};
template <class SM>
__device__ __forceinline__ bool tile_footprint(const SM& sm, const oake_resize_job& job, int tw, int th, int max_rows,
                                               int* __restrict__ err, Footprint* fp) {
  fp->r_lo = sm.v_first[0];
  fp->nr = sm.v_first[th - 1] + sm.v_count[th - 1] - fp->r_lo;
  if (fp->nr > max_rows) {
    if (threadIdx.x == 0) atomicExch(err, 1);
    fp->nr = max_rows;
  }
  fp->y_lo = job.box_y0 + fp->r_lo;
  fp->rv0 = max(0, -fp->y_lo);
  fp->rv1 = min(fp->nr, job.src_h - fp->y_lo);
  const int x_lo = job.box_x0 + sm.h_first[0], x_hi = job.box_x0 + sm.h_first[tw - 1] + sm.h_count[tw - 1];
  return fp->rv0 < fp->rv1 && x_hi > 0 && x_lo < job.src_w;
}

// Output of one thread's column `ox` of the window.  MODE 0: uint8 HWC at the job's dst_off.  MODE 1 / 2: the T50 /
// T197 front-end matrix (ResizeOut): the column part of the address is fixed per thread, a row adds a 32-bit offset,
// and the ToTensor + Normalize table sits in shared memory.
template <int MODE>
struct PixelSink {
  static constexpr int kShift = MODE == 2 ? 4 : 5, kPad = MODE == 2 ? kBlkPad : 0, kGrid = MODE == 2 ? kBlkGrid : 7;
  static constexpr int kCs = 1 << (2 * kShift);  // channel stride inside a matrix row
  uint8_t* u8;
  long long pitch3;
  act_t* col;
  const act_t* lut;
  __device__ __forceinline__ PixelSink(const oake_resize_job& job, uint8_t* __restrict__ dst_arena, const ResizeOut& out, int jb,
                                       int ox, const act_t* s_lut) {
    if (MODE == 0) {
      u8 = dst_arena + job.dst_off + static_cast<long long>(ox) * 3;
      pitch3 = static_cast<long long>(job.dst_pitch_px) * 3;
    } else {
      const int px = ox + kPad;
      col = out.matrix + (static_cast<size_t>(out.crop0 + jb) * kGrid * kGrid + (px >> kShift)) * 3 * kCs + (px & ((1 << kShift) - 1));
      lut = s_lut;
    }
  }
  __device__ __forceinline__ void put(int oy, int v0, int v1, int v2) const {
    if (MODE == 0) {
      uint8_t* o = u8 + oy * pitch3;
      o[0] = static_cast<uint8_t>(v0);
      o[1] = static_cast<uint8_t>(v1);
      o[2] = static_cast<uint8_t>(v2);
    } else {
      const int py = oy + kPad;
      act_t* o = col + ((py >> kShift) * (kGrid * 3 * kCs) + ((py & ((1 << kShift) - 1)) << kShift));
      o[0] = lut[v0];
      o[kCs] = lut[256 + v1];
      o[2 * kCs] = lut[512 + v2];
    }
  }
};

// (matrix output) the table -> shared memory; the caller's next barrier publishes it
template <int MODE, class SM>
__device__ __forceinline__ void tile_load_lut(SM& sm, const ResizeOut& out) {
  if (MODE != 0 && threadIdx.x < 768 / 8)
    reinterpret_cast<uint4*>(sm.lut)[threadIdx.x] = __ldg(reinterpret_cast<const uint4*>(out.lut) + threadIdx.x);
}

template <int MODE, class SM>
__device__ __forceinline__ void tile_zero(const SM& sm, const oake_resize_job& job, uint8_t* __restrict__ dst_arena,
                                          const ResizeOut& out, int jb, int ox0, int oy0, int tw, int th) {
  const int j = threadIdx.x & (kTile - 1);
  if (MODE != 0) __syncthreads();  // the table (block-uniform path: every thread of the CTA is here)
  if (j >= tw) return;
  const PixelSink<MODE> sink(job, dst_arena, out, jb, ox0 + j, sm.lut);
  for (int r = threadIdx.x / kTile; r < th; r += blockDim.x / kTile) sink.put(oy0 + r, 0, 0, 0);
}

// N taps of one output sample, three channels: `px` advances by `step` bytes per tap.
template <int N>
__device__ __forceinline__ void taps_fixed(const int* __restrict__ k, const uint8_t* __restrict__ px, int step, int& s0, int& s1,
                                           int& s2) {
#pragma unroll
  for (int t = 0; t < N; ++t) {
    const int kk = k[t];
    s0 += static_cast<int>(px[t * step + 0]) * kk;
    s1 += static_cast<int>(px[t * step + 1]) * kk;
    s2 += static_cast<int>(px[t * step + 2]) * kk;
  }
}

// Horizontal pass (crop rows y_lo + [rv0, rv1) -> tmp, uint8 like Pillow's intermediate image), barrier, vertical
// pass.  One thread per (row, output column): the three channels share the tap loop, and the taps that fall outside
// the source image are cut off once, outside the loop.
template <int MODE, class SM>
__device__ __forceinline__ void tile_pixels(SM& sm, const oake_resize_job& job, const uint8_t* __restrict__ src_arena,
                                            uint8_t* __restrict__ dst_arena, const ResizeOut& out, int jb, int ox0, int oy0,
                                            int tw, int th, const Footprint& fp) {
  const uint8_t* src = src_arena + job.src_off;
  const int tid = threadIdx.x;
  const int j = tid & (kTile - 1);  // output column of the tile; rows go tid / 32, + 8, + 16, ...
  {
    const bool live = j < tw;
    const int first = live ? job.box_x0 + sm.h_first[j] : 0;
    const int t0 = live ? max(0, -first) : 0;
    const int t1 = live ? min(sm.h_count[j], job.src_w - first) : 0;
    const int* k = sm.kh[live ? j : 0];
    const long long row_bytes = static_cast<long long>(job.src_pitch_px) * 3;
    const uint8_t* px_row = src + static_cast<long long>(fp.y_lo + fp.rv0 + tid / kTile) * row_bytes + (first + t0) * 3;
    // A warp = the 32 columns of one row.  When no column of the tile is clipped by the image's left / right edge, every
    // lane runs the warp's widest tap count N with its coefficients in registers (zero weights beyond its own count;
    // first + N stays inside the source row): no loop, no remainder, no shared-memory reads per row.
    const int nmax = __reduce_max_sync(0xffffffffu, t1 - t0);
    const bool fixed = OAKE_RESIZE_HFIXED && __all_sync(0xffffffffu, !live || (t0 == 0 && t1 == sm.h_count[j] && first + nmax <= job.src_w)) &&
                       nmax >= 4 && nmax <= 9;
    if (fixed) {
      int kr[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) kr[t] = t < t1 ? k[t] : 0;  // (t0 == 0: the lane's own taps, zero weights beyond)
      if (live) {
        for (int r = fp.rv0 + tid / kTile; r < fp.rv1; r += 256 / kTile, px_row += (256 / kTile) * row_bytes) {
          int s0 = 1 << (kPrecBits - 1), s1 = s0, s2 = s0;
          switch (nmax) {
            case 4: taps_fixed<4>(kr, px_row, 3, s0, s1, s2); break;
            case 5: taps_fixed<5>(kr, px_row, 3, s0, s1, s2); break;
            case 6: taps_fixed<6>(kr, px_row, 3, s0, s1, s2); break;
            case 7: taps_fixed<7>(kr, px_row, 3, s0, s1, s2); break;
            case 8: taps_fixed<8>(kr, px_row, 3, s0, s1, s2); break;
            default: taps_fixed<9>(kr, px_row, 3, s0, s1, s2); break;
          }
          uint8_t* o = sm.tmp[r][j];
          o[0] = pil_clip8(s0);
          o[1] = pil_clip8(s1);
          o[2] = pil_clip8(s2);
        }
      }
    } else if (live) {
      for (int r = fp.rv0 + tid / kTile; r < fp.rv1; r += 256 / kTile, px_row += (256 / kTile) * row_bytes) {
        int s0 = 1 << (kPrecBits - 1), s1 = s0, s2 = s0;
        const uint8_t* px = px_row;
#pragma unroll 4
        for (int t = t0; t < t1; ++t, px += 3) {
          const int kk = k[t];
          s0 += static_cast<int>(px[0]) * kk;
          s1 += static_cast<int>(px[1]) * kk;
          s2 += static_cast<int>(px[2]) * kk;
        }
        uint8_t* o = sm.tmp[r][j];
        o[0] = pil_clip8(s0);
        o[1] = pil_clip8(s1);
        o[2] = pil_clip8(s2);
      }
    }
  }
  __syncthreads();
  if (j >= tw) return;
  const PixelSink<MODE> sink(job, dst_arena, out, jb, ox0 + j, sm.lut);
  for (int r = tid / kTile; r < th; r += 256 / kTile) {
    const int first = sm.v_first[r] - fp.r_lo;
    const int t0 = max(0, fp.rv0 - first);
    const int t1 = min(sm.v_count[r], fp.rv1 - first);
    const int* k = sm.kv[r] + t0;
    const uint8_t* px = sm.tmp[first + t0][j];
    int s0 = 1 << (kPrecBits - 1), s1 = s0, s2 = s0;
    // a warp shares its row, hence the tap count: the common counts run without loop or remainder code
    switch (t1 - t0) {
      case 4: taps_fixed<4>(k, px, kTile * 3, s0, s1, s2); break;
      case 5: taps_fixed<5>(k, px, kTile * 3, s0, s1, s2); break;
      case 6: taps_fixed<6>(k, px, kTile * 3, s0, s1, s2); break;
      case 7: taps_fixed<7>(k, px, kTile * 3, s0, s1, s2); break;
      case 8: taps_fixed<8>(k, px, kTile * 3, s0, s1, s2); break;
      case 9: taps_fixed<9>(k, px, kTile * 3, s0, s1, s2); break;
      default:
#pragma unroll 4
        for (int t = t0; t < t1; ++t, px += kTile * 3, ++k) {
          const int kk = k[0];
          s0 += static_cast<int>(px[0]) * kk;
          s1 += static_cast<int>(px[1]) * kk;
          s2 += static_cast<int>(px[2]) * kk;
        }
    }
    sink.put(oy0 + r, pil_clip8(s0), pil_clip8(s1), pil_clip8(s2));
  }
}

template <int MODE>
__global__ void __launch_bounds__(256)
resize_fast_kernel(const uint8_t* __restrict__ src_arena, uint8_t* __restrict__ dst_arena,
                   const oake_resize_job* __restrict__ jobs, const ResizeJobInfo* __restrict__ info,
                   const int* __restrict__ tab, int* __restrict__ err, ResizeOut out) {
  __shared__ FastSmem sm;
  const int jb = blockIdx.y;
  const ResizeJobInfo ji = info[jb];
  if (ji.cls != 0 || ji.pad != 0) return;
  const oake_resize_job job = jobs[jb];
  const int tiles_x = (job.win_w + kTile - 1) / kTile;
  const int tiles_y = (job.win_h + kTile - 1) / kTile;
  if (static_cast<int>(blockIdx.x) >= tiles_x * tiles_y) return;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int ox0 = tx * kTile, oy0 = ty * kTile;  // relative to the window
  const int tw = min(kTile, job.win_w - ox0);
  const int th = min(kTile, job.win_h - oy0);
  tile_load_tables(sm, tab + ji.tab_off, job.win_w, job.win_h, tx, ty);
  tile_load_lut<MODE>(sm, out);
  __syncthreads();
  Footprint fp;
  if (!tile_footprint(sm, job, tw, th, kFastRows, err, &fp)) {
    tile_zero<MODE>(sm, job, dst_arena, out, jb, ox0, oy0, tw, th);
    return;
  }
  tile_pixels<MODE>(sm, job, src_arena, dst_arena, out, jb, ox0, oy0, tw, th, fp);
}

// Per-tile coefficient generation (BIG jobs without a table): the CTA shares the work -- pil_bounds one thread per
// output sample, pil_weight one thread per tap, pil_weight_sum one thread per sample (libImaging's order), the
// divide / fixed-point rounding of every tap independent again.
__device__ __forceinline__ void tile_generate(BigSmem& sm, const oake_resize_job& job, int ox0, int oy0, int tw, int th) {
  const int tid = threadIdx.x;
  const PilAxis ax_h = pil_axis(job.box_w, job.out_w);
  const PilAxis ax_v = pil_axis(job.box_h, job.out_h);
  const int gx0 = job.win_x + ox0, gy0 = job.win_y + oy0;
  if (tid < kTile) {
    if (tid < tw) pil_bounds(ax_h, job.box_w, gx0 + tid, &sm.h_first[tid], &sm.h_count[tid]);
  } else if (tid < 2 * kTile) {
    const int r = tid - kTile;
    if (r < th) pil_bounds(ax_v, job.box_h, gy0 + r, &sm.v_first[r], &sm.v_count[r]);
  }
  __syncthreads();
  // thread = (sample sidx = tid % 64, tap x = tid / 64, + 4, + 8, ...): a warp holds 32 samples at the same tap and
  // the loop stops at the widest tap window of the tile
  const int sidx = tid & (2 * kTile - 1);
  const bool s_h = sidx < kTile;
  const int s_r = sidx - kTile;
  const bool s_live = s_h ? sidx < tw : s_r < th;
  const int s_count = !s_live ? 0 : (s_h ? sm.h_count[sidx] : sm.v_count[s_r]);
  const int s_first = !s_live ? 0 : (s_h ? sm.h_first[sidx] : sm.v_first[s_r]);
  const int max_count = __reduce_max_sync(0xffffffffu, s_count);
  for (int x = tid / (2 * kTile); x < max_count; x += blockDim.x / (2 * kTile))
    if (x < s_count) sm.w[sidx][x] = pil_weight(s_h ? ax_h : ax_v, s_h ? gx0 + sidx : gy0 + s_r, s_first, x);
  __syncthreads();
  if (tid < kTile) {
    if (tid < tw) sm.ww[tid] = pil_weight_sum(sm.w[tid], sm.h_count[tid]);
  } else if (tid < 2 * kTile) {
    const int r = tid - kTile;
    if (r < th) sm.ww[tid] = pil_weight_sum(sm.w[tid], sm.v_count[r]);
  }
  __syncthreads();
  for (int x = tid / (2 * kTile); x < max_count; x += blockDim.x / (2 * kTile))
    if (x < s_count) (s_h ? sm.kh[sidx] : sm.kv[s_r])[x] = pil_fixed(sm.w[sidx][x], sm.ww[sidx]);
  __syncthreads();  // (the weights' storage is the pixel phase's tmp)
}

// Persistent grid over the (job, tile) pairs of the BIG jobs; nothing to do (and gone at once) when there are none.
template <int MODE>
__global__ void __launch_bounds__(256)
resize_big_kernel(const uint8_t* __restrict__ src_arena, uint8_t* __restrict__ dst_arena,
                  const oake_resize_job* __restrict__ jobs, ResizeState* __restrict__ state,
                  const ResizeJobInfo* __restrict__ info, const int4* __restrict__ big_list,
                  const int* __restrict__ tab, int* __restrict__ err, ResizeOut out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BigSmem& sm = *reinterpret_cast<BigSmem*>(smem_raw);
  const int total = state->big_tiles, n_big = state->n_big;
  tile_load_lut<MODE>(sm, out);  // (published by the loop's barriers)
  for (;;) {
    __syncthreads();  // the previous tile's shared memory is free
    if (threadIdx.x == 0) sm.work_tile = atomicAdd(&state->next, 1);
    __syncthreads();
    const int w = sm.work_tile;
    if (w >= total) return;
    __syncthreads();  // everyone has read the work index before the search overwrites it
    for (int e = threadIdx.x; e < n_big; e += blockDim.x) {
      const int4 it = big_list[e];
      if (w >= it.y && w < it.y + it.z) {
        sm.work_job = it.x;
        sm.work_tile = w - it.y;
      }
    }
    __syncthreads();
    const int jb = sm.work_job, tile = sm.work_tile;
    const oake_resize_job job = jobs[jb];
    const ResizeJobInfo ji = info[jb];
    const int tiles_x = (job.win_w + kTile - 1) / kTile;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int ox0 = tx * kTile, oy0 = ty * kTile;
    const int tw = min(kTile, job.win_w - ox0);
    const int th = min(kTile, job.win_h - oy0);
    if (ji.tab_off >= 0 && ji.pitch == BigSmem::kPitch) {
      tile_load_tables(sm, tab + ji.tab_off, job.win_w, job.win_h, tx, ty);
      __syncthreads();
    } else {  // the scratch had no room for this job's tables
      tile_generate(sm, job, ox0, oy0, tw, th);
    }
    Footprint fp;
    if (!tile_footprint(sm, job, tw, th, kMaxRows, err, &fp)) {
      tile_zero<MODE>(sm, job, dst_arena, out, jb, ox0, oy0, tw, th);
      continue;
    }
    tile_pixels<MODE>(sm, job, src_arena, dst_arena, out, jb, ox0, oy0, tw, th, fp);
  }
}

// ------------------------------------------------------------------------------------------------
// objects.py:129-155 masks.  box = the expanded square (float xyxy), fg = proposal - box.lt.
// x = arange(x2 - x1): the reference iterates todd BBoxes, which yields PYTHON floats (PIL's `crop` needs
// them: round(Tensor) raises), so x2 - x1 is a double difference of two fp32 values and arange takes its
// ceiling -- ceil(double width) samples, not ceilf of an fp32 difference.  Inside iff fg0 <= x <= fg2; the bool mask
// is resampled to 14x14 by F.interpolate(mode='nearest'): src = min(floor(dst * (n / 14.f)), n - 1).
// Output 1 = background.
// ------------------------------------------------------------------------------------------------
__global__ void object_masks_kernel(const float* __restrict__ fg, const float* __restrict__ box,
                                    float* __restrict__ masks, int B, int grid) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * grid * grid) return;
  const int b = idx / (grid * grid);
  const int cell = idx - b * grid * grid;
  const int gy = cell / grid, gx = cell - gy * grid;
  const float* bb = box + b * 4;
  const float* f = fg + b * 4;
  const int nx = static_cast<int>(ceil(static_cast<double>(bb[2]) - static_cast<double>(bb[0])));
  const int ny = static_cast<int>(ceil(static_cast<double>(bb[3]) - static_cast<double>(bb[1])));
  float bg = 1.f;
  if (nx > 0 && ny > 0) {
    const float sx = static_cast<float>(nx) / static_cast<float>(grid);
    const float sy = static_cast<float>(ny) / static_cast<float>(grid);
    const int ix = min(static_cast<int>(floorf(gx * sx)), nx - 1);
    const int iy = min(static_cast<int>(floorf(gy * sy)), ny - 1);
    const float x = static_cast<float>(ix), y = static_cast<float>(iy);
    const bool inside = (f[0] <= x) && (x <= f[2]) && (f[1] <= y) && (y <= f[3]);
    bg = inside ? 0.f : 1.f;
  }
  masks[idx] = bg;
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

cudaError_t launch_im2col_u8(cudaStream_t st, const uint8_t* arena, const oake_crop_src* crops,
                             const act_t* lut, act_t* patches, int B, int stride, int pad, int grid) {
  if (B <= 0) return cudaSuccess;
  const long long total = static_cast<long long>(B) * grid * grid * (kPatch * kPatch / 8);
  const long long blocks = (total + 255) / 256;
  im2col_u8_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(arena, crops, lut, patches, total, stride, pad,
                                                                  grid);
  return cudaGetLastError();
}

cudaError_t launch_blockcol_pixels(cudaStream_t st, const float* pixels, act_t* blocks, int B) {
  if (B <= 0) return cudaSuccess;
  const long long total = static_cast<long long>(B) * kBlkGrid * kBlkGrid * (kBlkCols / 8);
  blockcol_pixels_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(pixels, blocks, total);
  return cudaGetLastError();
}

cudaError_t launch_blockcol_u8(cudaStream_t st, const uint8_t* arena, const oake_crop_src* crops, const act_t* lut,
                               act_t* blocks, int B) {
  if (B <= 0) return cudaSuccess;
  const long long total = static_cast<long long>(B) * kBlkGrid * kBlkGrid * 32;
  blockcol_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(arena, crops, lut, blocks, total);
  return cudaGetLastError();
}

cudaError_t launch_conv1_regroup(cudaStream_t st, const act_t* w, act_t* out, int out_ch) {
  const int total = out_ch * kCols;
  conv1_regroup_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, out, total);
  return cudaGetLastError();
}

namespace {
// Stream-ordered scratch of the resize calls: a pool of its own per device that keeps what it has been given
// (the default pool hands memory back at every synchronisation).
cudaError_t resize_pool(cudaMemPool_t* out) {
  static cudaMemPool_t pools[64] = {};
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (pools[dev] == nullptr) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool;
    if (cudaError_t e = cudaMemPoolCreate(&pool, &props); e != cudaSuccess) return e;
    unsigned long long keep = ~0ull;
    if (cudaError_t e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); e != cudaSuccess) return e;
    pools[dev] = pool;
  }
  *out = pools[dev];
  return cudaSuccess;
}
constexpr int kResizeSlice = 8192;  // jobs per pass: bounds the table scratch at 8192 x 48 KB = 403 MB
// Test hooks (tests/test_gpu_frontend.py): $OAKE_RESIZE_SLICE = jobs per pass, $OAKE_RESIZE_NO_TABLES=1 = no room for
// any table, i.e. every job takes the per-tile coefficient generation of the BIG kernel.
int env_int(const char* name, int fallback) {
  const char* e = getenv(name);
  return (e != nullptr && e[0] != 0) ? atoi(e) : fallback;
}
}  // namespace

namespace {
cudaError_t launch_resize(cudaStream_t st, const uint8_t* src, uint8_t* dst, const oake_resize_job* jobs, int n_jobs,
                          int max_tiles, int* err_flag, ResizeOut out) {
  if (n_jobs <= 0 || max_tiles <= 0) return cudaSuccess;
  const int mode = out.mode == 0 ? 0 : (out.shift == 5 ? 1 : 2);
  {
    cudaError_t e = mode == 0   ? ensure_dynamic_smem<resize_big_kernel<0>>(static_cast<int>(sizeof(BigSmem)))
                    : mode == 1 ? ensure_dynamic_smem<resize_big_kernel<1>>(static_cast<int>(sizeof(BigSmem)))
                                : ensure_dynamic_smem<resize_big_kernel<2>>(static_cast<int>(sizeof(BigSmem)));
    if (e != cudaSuccess) return e;
  }
  int dev = 0, num_sms = 0;
  if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return e;
  if (cudaError_t e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); e != cudaSuccess) return e;
  cudaMemPool_t pool;
  if (cudaError_t e = resize_pool(&pool); e != cudaSuccess) return e;
  const int max_slice = std::max(1, env_int("OAKE_RESIZE_SLICE", kResizeSlice));
  const bool no_tables = env_int("OAKE_RESIZE_NO_TABLES", 0) != 0;
  const int slice = n_jobs < max_slice ? n_jobs : max_slice;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t off_info = 256;
  const size_t off_list = off_info + up(static_cast<size_t>(slice) * sizeof(ResizeJobInfo));
  const size_t off_tab = off_list + up(static_cast<size_t>(slice) * sizeof(int4));
  const size_t bytes = off_tab + (static_cast<size_t>(slice) * kJobTabInts + kTabSlackInts) * sizeof(int);
  uint8_t* scratch = nullptr;
  if (cudaError_t e = cudaMallocFromPoolAsync(reinterpret_cast<void**>(&scratch), bytes, pool, st); e != cudaSuccess) return e;
  ResizeState* state = reinterpret_cast<ResizeState*>(scratch);
  ResizeJobInfo* info = reinterpret_cast<ResizeJobInfo*>(scratch + off_info);
  int4* big_list = reinterpret_cast<int4*>(scratch + off_list);
  int* tab = reinterpret_cast<int*>(scratch + off_tab);
  cudaError_t err = cudaSuccess;
  for (int j0 = 0; j0 < n_jobs && err == cudaSuccess; j0 += slice) {
    const int n = n_jobs - j0 < slice ? n_jobs - j0 : slice;
    err = cudaMemsetAsync(state, 0, sizeof(ResizeState), st);
    if (err != cudaSuccess) break;
    ResizeOut o = out;
    o.crop0 = out.crop0 + j0;
    resize_prepare_kernel<<<n, 256, 0, st>>>(jobs + j0, no_tables ? 0 : n, state, info, big_list, tab, err_flag, o);
    const dim3 fg(max_tiles, n);
    const int bg = 4 * num_sms;
    const size_t bs = sizeof(BigSmem);
    if (mode == 0) {
      resize_fast_kernel<0><<<fg, 256, 0, st>>>(src, dst, jobs + j0, info, tab, err_flag, o);
      resize_big_kernel<0><<<bg, 256, bs, st>>>(src, dst, jobs + j0, state, info, big_list, tab, err_flag, o);
    } else if (mode == 1) {
      resize_fast_kernel<1><<<fg, 256, 0, st>>>(src, dst, jobs + j0, info, tab, err_flag, o);
      resize_big_kernel<1><<<bg, 256, bs, st>>>(src, dst, jobs + j0, state, info, big_list, tab, err_flag, o);
    } else {
      resize_fast_kernel<2><<<fg, 256, 0, st>>>(src, dst, jobs + j0, info, tab, err_flag, o);
      resize_big_kernel<2><<<bg, 256, bs, st>>>(src, dst, jobs + j0, state, info, big_list, tab, err_flag, o);
    }
    err = cudaGetLastError();
  }
  const cudaError_t e2 = cudaFreeAsync(scratch, st);
  return err != cudaSuccess ? err : e2;
}
}  // namespace

cudaError_t launch_resize_u8(cudaStream_t st, const uint8_t* src, uint8_t* dst, const oake_resize_job* jobs,
                             int n_jobs, int max_tiles, int* err_flag) {
  return launch_resize(st, src, dst, jobs, n_jobs, max_tiles, err_flag, ResizeOut{0, nullptr, nullptr, 0, 0, 0, 0});
}

cudaError_t launch_resize_to_matrix(cudaStream_t st, const uint8_t* src, const oake_resize_job* jobs, int n_jobs,
                                    int* err_flag, const act_t* lut, act_t* matrix, int blocks16) {
  // every job's window must be the whole 224 x 224 crop (checked by the prepare kernel: err_flag)
  return launch_resize(st, src, nullptr, jobs, n_jobs, 49, err_flag,
                       blocks16 ? ResizeOut{1, matrix, lut, 4, kBlkPad, kBlkGrid, 0} : ResizeOut{1, matrix, lut, 5, 0, 7, 0});
}

cudaError_t launch_object_masks(cudaStream_t st, const float* fg, const float* box, float* masks, int B,
                                int grid) {
  if (B <= 0) return cudaSuccess;
  const int total = B * grid * grid;
  object_masks_kernel<<<(total + 255) / 256, 256, 0, st>>>(fg, box, masks, B, grid);
  return cudaGetLastError();
}

int resize_max_taps() { return kMaxTaps; }

}  // namespace oake
