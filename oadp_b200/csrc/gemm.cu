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
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // + align slack
  static constexpr int kTmemCols = 2 * BN;
};

__device__ __forceinline__ float quick_gelu(float u) { return u / (1.0f + __expf(-1.702f * u)); }

// 32 consecutive output columns of one row.
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& ep, float (&v)[32], int row,
                                               int col0) {
  if (ep.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = __ldg(b4 + j);
      v[4 * j + 0] += b.x;
      v[4 * j + 1] += b.y;
      v[4 * j + 2] += b.z;
      v[4 * j + 3] += b.w;
    }
  }
  if (ep.act == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
  }
  if (ep.residual != nullptr) {
    const float4* r4 =
        reinterpret_cast<const float4*>(ep.residual + static_cast<size_t>(row) * ep.ld_res + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 r = r4[j];
      v[4 * j + 0] += r.x;
      v[4 * j + 1] += r.y;
      v[4 * j + 2] += r.z;
      v[4 * j + 3] += r.w;
    }
  }
  if (ep.out_f32) {
    float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(ep.out) +
                                           static_cast<size_t>(row) * ep.ldo + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    uint4* o4 = reinterpret_cast<uint4*>(static_cast<act_t*>(ep.out) +
                                         static_cast<size_t>(row) * ep.ldo + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u;
      u.x = pack2(v[8 * j + 0], v[8 * j + 1]);
      u.y = pack2(v[8 * j + 2], v[8 * j + 3]);
      u.z = pack2(v[8 * j + 4], v[8 * j + 5]);
      u.w = pack2(v[8 * j + 6], v[8 * j + 7]);
      o4[j] = u;
    }
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
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const uint32_t t_base =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_base + c * 32, r);
        tmem_ld_wait();
        if (row < M) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_chunk(ep, v, row, n_blk * BN + c * 32);
        }
      }
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
    if (ep.residual) v += ep.residual[static_cast<size_t>(row) * ep.ld_res + col];
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
