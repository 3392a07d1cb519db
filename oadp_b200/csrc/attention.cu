// Per-(crop, head) attention for the OAKE tower.
//
// Main stream: softmax((q/8) k^T) v over T = P+1 tokens (50 or 197), no mask -- what
// nn.MultiheadAttention computes inside every CLIP ResidualAttentionBlock (reference call sites
// oadp/oake/globals.py:57, blocks.py:129, objects.py:330).
// Side stream (objects only): the single y query attends over the P patch keys of the SAME
// layer's K/V (shared with the main stream instead of re-projected as the reference does at
// oadp/oake/objects.py:238-245) plus itself, with additive bias -100 * mask on the patches.
//
// Whole K/V/Q of one (crop, head) fit in shared memory (<= 90 KB), so there is no online
// softmax: S = QK^T lives in registers (mma.sync m16n8k16, fp32 accumulate), quad-shuffle row
// max/sum, P re-used in registers as the A operand of PV.  Attention is 1-4 % of the tower's
// FLOPs; the dense contractions go through the tcgen05 GEMM.
#include "kernels.cuh"

namespace oake {

namespace {

constexpr int kDh = 64;
constexpr int kLds = 72;  // padded smem row (elements): 144 B keeps ldmatrix conflict-free

template <int P>
struct ACfg {
  static constexpr int T = P + 1;
  static constexpr int TP = ((T + 15) / 16) * 16;
  static constexpr int MT = TP / 16;
  static constexpr int NT = TP / 8;
  static constexpr int kWarps = (MT <= 4) ? MT : (MT + 1) / 2;
  static constexpr int kThreads = kWarps * 32;
  static constexpr int kSmemBytes = 3 * TP * kLds * 2;
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                          uint32_t a3, uint32_t b0, uint32_t b1) {
#ifdef OAKE_USE_BF16
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
#else
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
#endif
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// local token i of crop b -> row of the activation matrix ([B*P patches | B class rows | ...]).
__device__ __forceinline__ int token_row(int i, int b, int B, int P) {
  return i < P ? b * P + i : B * P + b;
}

template <int P>
__global__ void __launch_bounds__(ACfg<P>::kThreads)
attention_main_kernel(const act_t* __restrict__ qkv, act_t* __restrict__ out, int B, int heads) {
  using C = ACfg<P>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  act_t* sQ = reinterpret_cast<act_t*>(smem_raw);
  act_t* sK = sQ + C::TP * kLds;
  act_t* sV = sK + C::TP * kLds;

  const int b = blockIdx.x / heads;
  const int h = blockIdx.x - b * heads;
  const int W = heads * kDh;
  const int ld = 3 * W;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int idx = threadIdx.x; idx < C::TP * 8; idx += C::kThreads) {
    const int i = idx >> 3;
    const int c = idx & 7;
    uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q;
    if (i < C::T) {
      const act_t* src = qkv + static_cast<size_t>(token_row(i, b, B, P)) * ld + h * kDh + c * 8;
      q = *reinterpret_cast<const uint4*>(src);
      k = *reinterpret_cast<const uint4*>(src + W);
      v = *reinterpret_cast<const uint4*>(src + 2 * W);
    }
    *reinterpret_cast<uint4*>(sQ + i * kLds + c * 8) = q;
    *reinterpret_cast<uint4*>(sK + i * kLds + c * 8) = k;
    *reinterpret_cast<uint4*>(sV + i * kLds + c * 8) = v;
  }
  __syncthreads();

  const int l8 = lane & 7;
  const int g1 = (lane >> 3) & 1;
  const int g2 = lane >> 4;
  const float scale = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)

  for (int mt = warp; mt < C::MT; mt += C::kWarps) {
    float s[C::NT][4];
#pragma unroll
    for (int n = 0; n < C::NT; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;

#pragma unroll
    for (int kk = 0; kk < kDh / 16; ++kk) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(smem_u32(sQ + (mt * 16 + l8 + g1 * 8) * kLds + kk * 16 + g2 * 8), a0, a1, a2, a3);
#pragma unroll
      for (int np = 0; np < C::NT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(sK + (np * 16 + l8 + g2 * 8) * kLds + kk * 16 + g1 * 8), b0, b1, b2, b3);
        mma_16816(s[2 * np], a0, a1, a2, a3, b0, b1);
        mma_16816(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }

    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < C::NT; ++n) {
      const int col = n * 8 + 2 * (lane & 3);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool valid = (col + (e & 1)) < C::T;
        s[n][e] = valid ? s[n][e] * scale : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
      mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int n = 0; n < C::NT; ++n) {
      s[n][0] = exp2f(s[n][0] - mx0);
      s[n][1] = exp2f(s[n][1] - mx0);
      s[n][2] = exp2f(s[n][2] - mx1);
      s[n][3] = exp2f(s[n][3] - mx1);
      sum0 += s[n][0] + s[n][1];
      sum1 += s[n][2] + s[n][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    float o[kDh / 8][4];
#pragma unroll
    for (int n = 0; n < kDh / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < C::MT; ++kt) {
      const uint32_t a0 = pack2(s[2 * kt][0], s[2 * kt][1]);
      const uint32_t a1 = pack2(s[2 * kt][2], s[2 * kt][3]);
      const uint32_t a2 = pack2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
      const uint32_t a3 = pack2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
      for (int np = 0; np < kDh / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(sV + (kt * 16 + l8 + g1 * 8) * kLds + np * 16 + g2 * 8), b0, b1, b2, b3);
        mma_16816(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma_16816(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }

    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    const int i0 = mt * 16 + (lane >> 2);
    const int i1 = i0 + 8;
    if (i0 < C::T) {
      act_t* dst = out + static_cast<size_t>(token_row(i0, b, B, P)) * W + h * kDh + 2 * (lane & 3);
#pragma unroll
      for (int n = 0; n < kDh / 8; ++n)
        *reinterpret_cast<uint32_t*>(dst + n * 8) = pack2(o[n][0] * inv0, o[n][1] * inv0);
    }
    if (i1 < C::T) {
      act_t* dst = out + static_cast<size_t>(token_row(i1, b, B, P)) * W + h * kDh + 2 * (lane & 3);
#pragma unroll
      for (int n = 0; n < kDh / 8; ++n)
        *reinterpret_cast<uint32_t*>(dst + n * 8) = pack2(o[n][2] * inv1, o[n][3] * inv1);
    }
  }
}

// One warp per (crop, head); 4 warps per CTA.
template <int P>
__global__ void __launch_bounds__(128)
attention_side_kernel(const act_t* __restrict__ qkv, const float* __restrict__ mask,
                      act_t* __restrict__ out, int B, int heads) {
  constexpr int NKEY = P + 1;
  constexpr int PER_LANE = (NKEY + 31) / 32;
  __shared__ float s_q[4][kDh];
  __shared__ float s_p[4][PER_LANE * 32];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 4 + warp;
  if (gw >= B * heads) return;
  const int b = gw / heads;
  const int h = gw - b * heads;
  const int W = heads * kDh;
  const int ld = 3 * W;
  const int yrow = B * P + B + b;

  {
    const float2 q2 = unpack2(
        *reinterpret_cast<const uint32_t*>(qkv + static_cast<size_t>(yrow) * ld + h * kDh + 2 * lane));
    s_q[warp][2 * lane] = q2.x;
    s_q[warp][2 * lane + 1] = q2.y;
  }
  __syncwarp();

  float sc[PER_LANE];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < PER_LANE; ++t) {
    const int j = lane + 32 * t;
    float v = -INFINITY;
    if (j < NKEY) {
      const int krow = j < P ? b * P + j : yrow;
      const uint4* k4 =
          reinterpret_cast<const uint4*>(qkv + static_cast<size_t>(krow) * ld + W + h * kDh);
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = k4[c];
        const float2 f0 = unpack2(u.x), f1 = unpack2(u.y), f2 = unpack2(u.z), f3 = unpack2(u.w);
        const float* q = &s_q[warp][c * 8];
        dot += f0.x * q[0] + f0.y * q[1] + f1.x * q[2] + f1.y * q[3] + f2.x * q[4] + f2.y * q[5] +
               f3.x * q[6] + f3.y * q[7];
      }
      // objects.py:204-214: bias = -100 * mask on patches (finite), 0 on the y key itself.
      const float bias = j < P ? -100.0f * mask[static_cast<size_t>(b) * P + j] : 0.f;
      v = dot * 0.125f + bias;
    }
    sc[t] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < PER_LANE; ++t) {
    const float p = (lane + 32 * t < NKEY) ? __expf(sc[t] - mx) : 0.f;
    s_p[warp][lane + 32 * t] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  __syncwarp();

  float o0 = 0.f, o1 = 0.f;
  for (int j = 0; j < NKEY; ++j) {
    const int vrow = j < P ? b * P + j : yrow;
    const float2 v2 = unpack2(*reinterpret_cast<const uint32_t*>(
        qkv + static_cast<size_t>(vrow) * ld + 2 * W + h * kDh + 2 * lane));
    const float p = s_p[warp][j];
    o0 += p * v2.x;
    o1 += p * v2.y;
  }
  const float inv = 1.0f / sum;
  *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(yrow) * W + h * kDh + 2 * lane) =
      pack2(o0 * inv, o1 * inv);
}

template <int P>
cudaError_t launch_main_p(cudaStream_t st, const act_t* qkv, act_t* out, int B, int heads) {
  using C = ACfg<P>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_main_kernel<P>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  attention_main_kernel<P><<<B * heads, C::kThreads, C::kSmemBytes, st>>>(qkv, out, B, heads);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_attention_main(cudaStream_t st, const act_t* qkv, act_t* out, int B, int P,
                                  int heads) {
  if (B <= 0) return cudaSuccess;
  if (P == 49) return launch_main_p<49>(st, qkv, out, B, heads);
  if (P == 196) return launch_main_p<196>(st, qkv, out, B, heads);
  return cudaErrorInvalidValue;
}

cudaError_t launch_attention_side(cudaStream_t st, const act_t* qkv, const float* mask, act_t* out,
                                  int B, int P, int heads) {
  if (B <= 0) return cudaSuccess;
  if (P != 196) return cudaErrorInvalidValue;
  attention_side_kernel<196><<<(B * heads + 3) / 4, 128, 0, st>>>(qkv, mask, out, B, heads);
  return cudaGetLastError();
}

}  // namespace oake
