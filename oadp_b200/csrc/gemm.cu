// Persistent warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   out[M,N] = epi( A[M,K] (act_t, K-major) x W[N,K]^T (act_t, K-major) ),  fp32 accumulate
//
// This one kernel carries every dense contraction of the OAKE tower (SURVEY 2.2 K1,K3,K5,K6,K8:
// patch embedding after im2col, QKV, attention out-proj + residual, c_fc + QuickGELU,
// c_proj + residual, final projection), replacing the cuDNN/cuBLAS calls the reference makes
// through clip.model.VisionTransformer (call sites oadp/oake/globals.py:57, blocks.py:129,
// objects.py:330).
//
// Structure (one CTA per SM, 256 threads):
//   warp 0    : TMA producer   -- cp.async.bulk.tensor 128x64 A tile + BNx64 W tile per stage,
//                                 128B-swizzled, completion on the stage's `full` mbarrier
//   warp 1    : MMA issuer     -- one elected thread issues 4 x tcgen05.mma (K=16) per stage into
//                                 a TMEM accumulator, tcgen05.commit releases the stage
//   warp 2    : TMEM allocator
//   warps 4-7 : epilogue       -- tcgen05.ld the finished 128xBN fp32 accumulator (double buffered
//                                 in TMEM so the next tile's MMAs overlap), bias / QuickGELU /
//                                 fp32 residual, store fp16 or fp32
// Tiles are walked n-fastest so the CTAs running concurrently share the A rows in L2; the
// weights (<= 4.7 MB) stay L2-resident.
#include <stdio.h>

#include "kernels.cuh"

namespace oake {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;

template <int BN>
struct Cfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kStgRowBytes = 128 + 16;              // 128 B of payload + 16 B bank skew
  static constexpr int kStgBytes = 32 * kStgRowBytes;        // one 32-row strip per epilogue warp
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 4 * kStgBytes + 1024;
  static constexpr int kTmemCols = 2 * BN;
};

// QuickGELU u * sigmoid(1.702 u) = h + h * tanh(0.851 u), h = u / 2: one MUFU per element.
__device__ __forceinline__ float quick_gelu(float u) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * u));
  const float h = 0.5f * u;
  return fmaf(h, t, h);
}

__device__ __forceinline__ void add_bias32(const float* __restrict__ bias, float (&v)[32]) {
  const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b = __ldg(b4 + j);
    v[4 * j + 0] += b.x;
    v[4 * j + 1] += b.y;
    v[4 * j + 2] += b.z;
    v[4 * j + 3] += b.w;
  }
}

// Epilogue of one warp on one 32-row strip of the accumulator.  Each thread owns a row in TMEM, but
// a thread-per-row global access pattern touches half sectors 32 rows apart; instead every
// 128-byte row segment goes through a per-warp shared-memory strip (144-byte row pitch:
// conflict-free both ways) so that each global instruction moves 4 rows x 128 contiguous bytes.
//
// act_t output (the tower's hot epilogues), chunks of 64 columns:
//   y = [LayerNorm fold] rstd_m * (acc - mean_m * s_n) + c_n   (W was pre-multiplied by gamma,
//        s_n = sum_k W'[n,k], c_n = sum_k beta_k W[n,k] + b_n; mean/rstd from the row statistics
//        the PRODUCER of the A rows accumulated) | acc + bias_n
//   y = QuickGELU(y)                       (c_fc)
//   y += residual (fp16, may alias out)    (out_proj, c_proj)
//   out = fp16(y);  out_stats[m] += (sum y, sum y^2) over this tile's columns of the rounded y
// fp32 output (patch embedding, final projection), chunks of 32 columns: bias / QuickGELU only.
struct RowLn {
  float rstd, nmr;  // y = rstd * acc + nmr * s_n + c_n,  nmr = -mean * rstd
};

__device__ __forceinline__ void load_residual_chunk(uint4 (&res)[8], const GemmEpilogue& ep, int row0,
                                                    int col0, int M, int sub_r, int sub_c) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = row0 + 4 * i + sub_r;
    res[i] = make_uint4(0, 0, 0, 0);
    if (r < M)
      res[i] = *reinterpret_cast<const uint4*>(
          reinterpret_cast<const uint8_t*>(ep.residual + static_cast<size_t>(r) * ep.ld_res + col0) + sub_c);
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_strip_act(const GemmEpilogue& ep, uint32_t t_strip, uint8_t* stg,
                                                   int row0, int col_base, int M, int lane, uint4 (&res)[8],
                                                   const RowLn ln) {
  constexpr int kPitch = Cfg<BN>::kStgRowBytes;
  constexpr int NC = BN / 64;
  uint8_t* my_row = stg + lane * kPitch;
  const int sub_r = lane >> 3;        // coalesced phase: 4 rows per instruction ...
  const int sub_c = (lane & 7) * 16;  // ... 8 lanes x 16 B per row
  const bool has_res = ep.residual != nullptr;
  float sum = 0.f, sumsq = 0.f;
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    const int col0 = col_base + c * 64;
    if (has_res) {
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(stg + (4 * i + sub_r) * kPitch + sub_c) = res[i];
      if (c + 1 < NC) load_residual_chunk(res, ep, row0, col0 + 64, M, sub_r, sub_c);  // in flight below
      __syncwarp();
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t r32[32];
      tmem_ld_32x32(t_strip + c * 64 + half * 32, r32);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r32[j]);
      const int cc = col0 + half * 32;
      if (ep.colsum != nullptr) {  // y = rstd * acc + (nmr * s_n + c_n)
        const float4* s4 = reinterpret_cast<const float4*>(ep.colsum + cc);
        const float4* c4 = reinterpret_cast<const float4*>(ep.bias + cc);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 sv = __ldg(s4 + j);
          const float4 cv = __ldg(c4 + j);
          v[4 * j + 0] = fmaf(ln.rstd, v[4 * j + 0], fmaf(ln.nmr, sv.x, cv.x));
          v[4 * j + 1] = fmaf(ln.rstd, v[4 * j + 1], fmaf(ln.nmr, sv.y, cv.y));
          v[4 * j + 2] = fmaf(ln.rstd, v[4 * j + 2], fmaf(ln.nmr, sv.z, cv.z));
          v[4 * j + 3] = fmaf(ln.rstd, v[4 * j + 3], fmaf(ln.nmr, sv.w, cv.w));
        }
      } else if (ep.bias != nullptr) {
        add_bias32(ep.bias + cc, v);
      }
      if (ep.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
      }
      uint4* seg = reinterpret_cast<uint4*>(my_row + half * 64);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (has_res) {
          const uint4 rv = seg[j];
          const float2 r0 = unpack2(rv.x), r1 = unpack2(rv.y), r2 = unpack2(rv.z), r3 = unpack2(rv.w);
          v[8 * j + 0] += r0.x;
          v[8 * j + 1] += r0.y;
          v[8 * j + 2] += r1.x;
          v[8 * j + 3] += r1.y;
          v[8 * j + 4] += r2.x;
          v[8 * j + 5] += r2.y;
          v[8 * j + 6] += r3.x;
          v[8 * j + 7] += r3.y;
        }
        uint4 u;
        u.x = pack2(v[8 * j + 0], v[8 * j + 1]);
        u.y = pack2(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack2(v[8 * j + 4], v[8 * j + 5]);
        u.w = pack2(v[8 * j + 6], v[8 * j + 7]);
        seg[j] = u;
        if (ep.out_stats != nullptr) {  // statistics of the values as stored (rounded)
          const float2 a = unpack2(u.x), b = unpack2(u.y), cdd = unpack2(u.z), d = unpack2(u.w);
          sum += (a.x + a.y) + (b.x + b.y) + (cdd.x + cdd.y) + (d.x + d.y);
          sumsq += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + cdd.x * cdd.x + cdd.y * cdd.y + d.x * d.x + d.y * d.y;
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + sub_r;
      if (row0 + r < M)
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(
            static_cast<act_t*>(ep.out) + static_cast<size_t>(row0 + r) * ep.ldo + col0) + sub_c) =
            *reinterpret_cast<const uint4*>(stg + r * kPitch + sub_c);
    }
    __syncwarp();
  }
  // One slot per 256-column block, summed in a fixed order by the consumer: deterministic (no atomics).
  if (ep.out_stats != nullptr && row0 + lane < M)
    ep.out_stats[static_cast<size_t>(row0 + lane) * kStatSlots + col_base / 256] = make_float2(sum, sumsq);
}

template <int BN>
__device__ __forceinline__ void epilogue_strip_f32(const GemmEpilogue& ep, uint32_t t_strip, uint8_t* stg,
                                                   int row0, int col_base, int M, int lane) {
  constexpr int kPitch = Cfg<BN>::kStgRowBytes;
  uint8_t* my_row = stg + lane * kPitch;
  const int sub_r = lane >> 3;
  const int sub_c = (lane & 7) * 16;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    const int col0 = col_base + c * 32;
    uint32_t r32[32];
    tmem_ld_32x32(t_strip + c * 32, r32);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r32[j]);
    if (ep.bias != nullptr) add_bias32(ep.bias + col0, v);
    if (ep.act == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(my_row + j * 16) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + sub_r;
      if (row0 + r < M)
        *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(
            static_cast<float*>(ep.out) + static_cast<size_t>(row0 + r) * ep.ldo + col0) + sub_c) =
            *reinterpret_cast<const float4*>(stg + r * kPitch + sub_c);
    }
    __syncwarp();
  }
}

template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    int M, int N, int K, GemmEpilogue ep) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::kStages * C::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full_bar = empty_bar + C::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint8_t* stg_base = smem + C::kStages * C::kStageBytes + C::kBarBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM;
  const int num_n = N / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = K / BK;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n;
        const int n_blk = tile - m_blk * num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
          tma_load_2d(sA + stage * C::kABytes, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(sB + stage * C::kBBytes, &tmW, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * C::kABytes);
          const uint32_t b_addr = smem_u32(sB + stage * C::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t a_desc = make_smem_desc_k_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t b_desc = make_smem_desc_k_sw128(b_addr + k * UMMA_K * 2);
            umma_f16(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------------- epilogue
    const int q = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n;
      const int n_blk = tile - m_blk * num_n;
      const int row0 = m_blk * BM + q * 32;
      uint8_t* stg = stg_base + q * C::kStgBytes;
      // Everything that does not depend on the accumulator is fetched before waiting for it.
      uint4 res[8];
      RowLn ln{1.f, 0.f};
      if (!ep.out_f32) {
        if (ep.residual != nullptr) load_residual_chunk(res, ep, row0, n_blk * BN, M, lane >> 3, (lane & 7) * 16);
        if (ep.colsum != nullptr && row0 + lane < M) {
          const float4* sp = reinterpret_cast<const float4*>(ep.ln_stats + static_cast<size_t>(row0 + lane) * kStatSlots);
          const float4 s01 = sp[0], s23 = sp[1];
          const float sx = ((s01.x + s01.z) + s23.x) + s23.z;
          const float sxx = ((s01.y + s01.w) + s23.y) + s23.w;
          const float inv_k = 1.0f / static_cast<float>(K);
          const float mean = sx * inv_k;
          const float var = fmaxf(sxx * inv_k - mean * mean, 0.f);
          ln.rstd = rsqrtf(var + 1e-5f);
          ln.nmr = -mean * ln.rstd;
        }
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_strip =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
      if (ep.out_f32)
        epilogue_strip_f32<BN>(ep, t_strip, stg, row0, n_blk * BN, M, lane);
      else
        epilogue_strip_act<BN>(ep, t_strip, stg, row0, n_blk * BN, M, lane, res, ln);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// ----------------------------------------------------------------------------- SIMT reference
__global__ void gemm_simt_kernel(const act_t* __restrict__ A, const act_t* __restrict__ W, int M,
                                 int N, int K, GemmEpilogue ep) {
  __shared__ float sa[16][17];
  __shared__ float sw[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row = blockIdx.y * 16 + ty;
  const int col = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    sa[ty][tx] = (row < M) ? from_act(A[static_cast<size_t>(row) * K + k0 + tx]) : 0.f;
    const int wr = blockIdx.x * 16 + ty;
    sw[ty][tx] = (wr < N) ? from_act(W[static_cast<size_t>(wr) * K + k0 + tx]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += sa[ty][k] * sw[tx][k];
    __syncthreads();
  }
  if (row < M && col < N) {
    float v = acc;
    if (ep.bias) v += ep.bias[col];
    if (ep.act == 1) v = quick_gelu(v);
    if (ep.residual) v += from_act(ep.residual[static_cast<size_t>(row) * ep.ld_res + col]);
    if (ep.out_f32)
      static_cast<float*>(ep.out)[static_cast<size_t>(row) * ep.ldo + col] = v;
    else
      static_cast<act_t*>(ep.out)[static_cast<size_t>(row) * ep.ldo + col] = to_act(v);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

}  // namespace

int make_tmap_act_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                     uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) return -1;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * sizeof(act_t)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), box_rows};
  cuuint32_t estr[2] = {1, 1};
#ifdef OAKE_USE_BF16
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

int gemm_block_n(int N) { return (N % 256 == 0) ? 256 : 128; }

template <int BN>
static cudaError_t launch_gemm_bn(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmW,
                                  int M, int N, int K, const GemmEpilogue& ep, int num_sms) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = ((M + BM - 1) / BM) * (N / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_tcgen05_kernel<BN><<<grid, 256, C::kSmemBytes, st>>>(tmA, tmW, M, N, K, ep);
  return cudaGetLastError();
}

cudaError_t launch_gemm(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmW, int M, int N,
                        int K, const GemmEpilogue& ep, int num_sms) {
  if (M <= 0) return cudaSuccess;
  if (N % 128 != 0 || K % BK != 0) return cudaErrorInvalidValue;
  if (gemm_block_n(N) == 256) return launch_gemm_bn<256>(st, tmA, tmW, M, N, K, ep, num_sms);
  return launch_gemm_bn<128>(st, tmA, tmW, M, N, K, ep, num_sms);
}

cudaError_t launch_gemm_simt(cudaStream_t st, const act_t* A, const act_t* W, int M, int N, int K,
                             const GemmEpilogue& ep) {
  if (M <= 0) return cudaSuccess;
  dim3 grid((N + 15) / 16, (M + 15) / 16), block(16, 16);
  gemm_simt_kernel<<<grid, block, 0, st>>>(A, W, M, N, K, ep);
  return cudaGetLastError();
}

}  // namespace oake
