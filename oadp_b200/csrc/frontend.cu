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

struct ResizeSmem {
  int kh[kTile][kMaxTaps];
  int kv[kTile][kMaxTaps];
  int h_first[kTile], h_count[kTile], v_first[kTile], v_count[kTile];
  int r_lo, r_hi;
  double ww[2 * kTile];
  union {
    double w[2 * kTile][kMaxTaps];    // raw filter weights (coefficient phase): [0,32) horizontal, [32,64) vertical
    uint8_t tmp[kMaxRows][kTile][3];  // horizontally resampled rows (pixel phase)
  };
};

__global__ void __launch_bounds__(256)
resize_u8_kernel(const uint8_t* __restrict__ src_arena, uint8_t* __restrict__ dst_arena,
                 const oake_resize_job* __restrict__ jobs, int* __restrict__ err) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  ResizeSmem& sm = *reinterpret_cast<ResizeSmem*>(smem_raw);
  const oake_resize_job job = jobs[blockIdx.y];
  const int tiles_x = (job.win_w + kTile - 1) / kTile;
  const int tiles_y = (job.win_h + kTile - 1) / kTile;
  if (static_cast<int>(blockIdx.x) >= tiles_x * tiles_y) return;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int ox0 = job.win_x + tx * kTile, oy0 = job.win_y + ty * kTile;
  const int tw = min(kTile, job.win_x + job.win_w - ox0);
  const int th = min(kTile, job.win_y + job.win_h - oy0);
  const int tid = threadIdx.x;

  const PilAxis ax_h = pil_axis(job.box_w, job.out_w);
  const PilAxis ax_v = pil_axis(job.box_h, job.out_h);
  if (tid < kTile) {
    if (tid < tw) pil_bounds(ax_h, job.box_w, ox0 + tid, &sm.h_first[tid], &sm.h_count[tid]);
  } else if (tid < 2 * kTile) {
    const int r = tid - kTile;
    if (r < th) pil_bounds(ax_v, job.box_h, oy0 + r, &sm.v_first[r], &sm.v_count[r]);
  }
  __syncthreads();
  // One filter tap per thread: thread = (sample sidx = tid % 64, tap x = tid / 64, + 4, + 8, ...), so a
  // warp holds 32 samples at the same tap and the loop stops at the widest tap window of the tile
  // (typically 5..9 of the 48 slots).
  const int sidx = tid & (2 * kTile - 1);
  const bool s_h = sidx < kTile;
  const int s_r = sidx - kTile;
  const bool s_live = s_h ? sidx < tw : s_r < th;
  const int s_count = !s_live ? 0 : (s_h ? sm.h_count[sidx] : sm.v_count[s_r]);
  const int s_first = !s_live ? 0 : (s_h ? sm.h_first[sidx] : sm.v_first[s_r]);
  const int max_count = __reduce_max_sync(0xffffffffu, s_count);  // widest window among this warp's 32 samples
  for (int x = tid / (2 * kTile); x < max_count; x += blockDim.x / (2 * kTile))
    if (x < s_count) sm.w[sidx][x] = pil_weight(s_h ? ax_h : ax_v, s_h ? ox0 + sidx : oy0 + s_r, s_first, x);
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
  __syncthreads();
  if (tid == 0) {
    int lo = sm.v_first[0], hi = 0;
    for (int r = 0; r < th; ++r) hi = max(hi, sm.v_first[r] + sm.v_count[r]);
    sm.r_lo = lo;
    sm.r_hi = hi;
    if (hi - lo > kMaxRows) atomicExch(err, 1);
  }
  __syncthreads();
  const int r_lo = sm.r_lo;
  const int nr = min(sm.r_hi - r_lo, kMaxRows);
  const uint8_t* src = src_arena + job.src_off;

  // horizontal pass: crop rows [r_lo, r_lo + nr) -> tmp (uint8, like Pillow's intermediate image).
  // One thread per (row, output column): the three channels share the tap loop, and the taps that
  // fall outside the source image (zero padding of PIL `crop`) are cut off once, outside the loop.
  const int j = tid & (kTile - 1);  // output column of the tile; rows go tid / 32, + 8, + 16, ...
  for (int r = tid / kTile; r < nr && j < tw; r += blockDim.x / kTile) {
    const int y = job.box_y0 + r_lo + r;
    int s0 = 1 << (kPrecBits - 1), s1 = s0, s2 = s0;
    if (y >= 0 && y < job.src_h) {
      const int first = job.box_x0 + sm.h_first[j];
      const int t0 = max(0, -first);
      const int t1 = min(sm.h_count[j], job.src_w - first);
      const uint8_t* px = src + (static_cast<long long>(y) * job.src_pitch_px + first + t0) * 3;
      const int* k = sm.kh[j];
#pragma unroll 4
      for (int t = t0; t < t1; ++t, px += 3) {
        const int kk = k[t];
        s0 += static_cast<int>(px[0]) * kk;
        s1 += static_cast<int>(px[1]) * kk;
        s2 += static_cast<int>(px[2]) * kk;
      }
    }
    uint8_t* o = sm.tmp[r][j];
    o[0] = pil_clip8(s0);
    o[1] = pil_clip8(s1);
    o[2] = pil_clip8(s2);
  }
  __syncthreads();

  // vertical pass
  uint8_t* dst = dst_arena + job.dst_off;
  for (int r = tid / kTile; r < th && j < tw; r += blockDim.x / kTile) {
    const int first = sm.v_first[r] - r_lo;
    const int t1 = min(sm.v_count[r], nr - first);
    const int* k = sm.kv[r];
    const uint8_t* px = sm.tmp[first][j];
    int s0 = 1 << (kPrecBits - 1), s1 = s0, s2 = s0;
#pragma unroll 4
    for (int t = 0; t < t1; ++t, px += kTile * 3) {
      const int kk = k[t];
      s0 += static_cast<int>(px[0]) * kk;
      s1 += static_cast<int>(px[1]) * kk;
      s2 += static_cast<int>(px[2]) * kk;
    }
    const int oy = oy0 - job.win_y + r, ox = ox0 - job.win_x + j;
    uint8_t* o = dst + (static_cast<long long>(oy) * job.dst_pitch_px + ox) * 3;
    o[0] = pil_clip8(s0);
    o[1] = pil_clip8(s1);
    o[2] = pil_clip8(s2);
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

cudaError_t launch_resize_u8(cudaStream_t st, const uint8_t* src, uint8_t* dst, const oake_resize_job* jobs,
                             int n_jobs, int max_tiles, int* err_flag) {
  if (n_jobs <= 0 || max_tiles <= 0) return cudaSuccess;
  if (cudaError_t e = ensure_dynamic_smem<resize_u8_kernel>(static_cast<int>(sizeof(ResizeSmem))); e != cudaSuccess)
    return e;
  dim3 grid(max_tiles, n_jobs);
  resize_u8_kernel<<<grid, 256, sizeof(ResizeSmem), st>>>(src, dst, jobs, err_flag);
  return cudaGetLastError();
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
