// Per-(crop, head) attention for the OAKE tower, one fused kernel.
//
// Main stream: softmax((q/8) k^T) v over T = P+1 tokens (50 or 197), no mask -- what
// nn.MultiheadAttention computes inside every CLIP ResidualAttentionBlock (reference call sites
// oadp/oake/globals.py:57, blocks.py:129, objects.py:330).
// Side stream (objects only): the y query attends over the P patch keys of the SAME layer's K/V
// (shared with the main stream instead of re-projected as the reference does at
// oadp/oake/objects.py:238-245) plus itself, with additive bias -100 * mask on the patches.  It is
// simply one more query row (and one more key/value row that only it may see) of the same tile, so
// it costs no extra launch and no second pass over K/V.
//
// All of K/V/Q of one (crop, head) fit in shared memory (<= 90 KB).  Each warp owns 32 query rows
// (two m16 tiles sharing every K/V fragment it loads) and walks the keys in chunks of 32 with an
// online softmax, so only a 32-key slice of S lives in registers: ~2 CTAs/SM instead of 1, half
// the ldmatrix traffic of a one-tile-per-warp layout (mma.sync m16n8k16, fp32 accumulate).
// Attention is 1-4 % of the tower's FLOPs; the dense contractions go through the tcgen05 GEMM.
#include <stdlib.h>

#include "kernels.cuh"

namespace oake {

namespace {

constexpr int kDh = 64;
constexpr int kLds = 72;  // padded smem row (elements): 144 B keeps ldmatrix conflict-free
constexpr float kLog2e = 1.4426950408889634f;

template <int P, bool SIDE>
struct ACfg {
  static constexpr int T = P + 1;             // main-stream tokens: P patches + class
  static constexpr int TQ = T + (SIDE ? 1 : 0);  // + the side token y
  static constexpr int TP = ((TQ + 15) / 16) * 16;
  static constexpr int MT = TP / 16;
  static constexpr int WPH = (MT + 1) / 2;    // warps per head (two m-tiles each)
  static constexpr int HPC = 1;               // heads per CTA (the async load below assumes 1)
  static constexpr int kMinCtas = (P == 49) ? 8 : 2;
  static constexpr int kThreads = WPH * HPC * 32;
  static constexpr int kHeadSmem = 3 * TP * kLds * 2;
  static constexpr int kSmemBytes = HPC * kHeadSmem;
  static constexpr int NCH = (TP + 31) / 32;  // key chunks of 32
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

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const int bytes = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
// waits until at most `n` of this thread's most recent groups are still in flight (n is a small
// compile-time-known value after unrolling; the switch keeps the immediate operand form)
__device__ __forceinline__ void cp_async_wait_pending(int n) {
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;\n" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;\n" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;\n" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;\n" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;\n" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;\n" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 6;\n" ::: "memory"); break;
  }
}

// local token i of crop b -> row of the activation matrix [B*P patches | B class rows | B side rows]
__device__ __forceinline__ int token_row(int i, int b, int B, int P) {
  return i < P ? b * P + i : (i == P ? B * P + b : B * P + B + b);
}

// side_only: the last block of the objects tower needs only the y row (its main-stream output is
// dead); every warp but the one owning that row retires after the loads.
template <int P, bool SIDE>
__global__ void __launch_bounds__(ACfg<P, SIDE>::kThreads, ACfg<P, SIDE>::kMinCtas)
attention_kernel(const act_t* __restrict__ qkv, const float* __restrict__ mask, act_t* __restrict__ out,
                 int B, int heads, int side_only) {
  using C = ACfg<P, SIDE>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int groups = heads / C::HPC;
  const int b = blockIdx.x / groups;
  const int hg = blockIdx.x - b * groups;
  const int W = heads * kDh;
  const int ld = 3 * W;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- asynchronous load of Q/K/V (rows >= TQ zero-filled), one cp.async group per 32-key chunk:
  //      group 0 = Q + keys [0,32), group c = keys [32c, 32c+32).  The chunk loop below waits for
  //      exactly the group it is about to consume, so the rest of K/V streams in under the math.
  {
    act_t* sQw = reinterpret_cast<act_t*>(smem_raw);
    act_t* sKw = sQw + C::TP * kLds;
    act_t* sVw = sKw + C::TP * kLds;
    const act_t* base = qkv + (hg * C::HPC) * kDh;
    for (int idx = threadIdx.x; idx < C::TP * 8; idx += C::kThreads) {
      const int i = idx >> 3, c = idx & 7;
      const bool ok = i < C::TQ;
      const act_t* src = base + static_cast<size_t>(token_row(ok ? i : 0, b, B, P)) * ld + c * 8;
      cp_async16_zfill(sQw + i * kLds + c * 8, src, ok);
    }
#pragma unroll 1
    for (int ch = 0; ch < C::NCH; ++ch) {
      const int rows = min(32, C::TP - ch * 32);
      for (int idx = threadIdx.x; idx < rows * 8; idx += C::kThreads) {
        const int i = ch * 32 + (idx >> 3), c = idx & 7;
        const bool ok = i < C::TQ;
        const act_t* src = base + static_cast<size_t>(token_row(ok ? i : 0, b, B, P)) * ld + c * 8;
        cp_async16_zfill(sKw + i * kLds + c * 8, src + W, ok);
        cp_async16_zfill(sVw + i * kLds + c * 8, src + 2 * W, ok);
      }
      cp_async_commit();
    }
  }

  const int hl = warp / C::WPH;
  const int wq = warp - hl * C::WPH;
  const int h = hg * C::HPC + hl;
  const int row_base = wq * 32;  // this warp's 32 query rows: m-tiles 2wq and 2wq+1
  const bool idle = SIDE && side_only && !(row_base <= C::T && C::T < row_base + 32);  // keeps hitting the barriers
  const act_t* sQ = reinterpret_cast<const act_t*>(smem_raw + hl * C::kHeadSmem);
  const act_t* sK = sQ + C::TP * kLds;
  const act_t* sV = sK + C::TP * kLds;
  const bool second = row_base + 16 < C::TP;  // the last warp may own a single m-tile

  const int l8 = lane & 7;
  const int g1 = (lane >> 3) & 1;
  const int g2 = lane >> 4;
  const float scale = 0.125f * kLog2e;  // 1/sqrt(64), folded into the base-2 exponent

  // rows of this thread: mt*16 + lane/4 (+8)
  float o[2][kDh / 8][4];
  float mrow[2][2], lrow[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int n = 0; n < kDh / 8; ++n) o[mt][n][0] = o[mt][n][1] = o[mt][n][2] = o[mt][n][3] = 0.f;
    mrow[mt][0] = mrow[mt][1] = -INFINITY;
    lrow[mt][0] = lrow[mt][1] = 0.f;
  }
  // which of this thread's rows is the side token (at most one, in the tile that contains row T)
  bool is_y[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) is_y[mt][hh] = SIDE && (row_base + mt * 16 + (lane >> 2) + hh * 8 == C::T);

#pragma unroll
  for (int c = 0; c < C::NCH; ++c) {
    const int key0 = c * 32;
    cp_async_wait_pending(C::NCH - 1 - c);  // this thread's copies of chunk c (and Q) have landed ...
    __syncthreads();                        // ... and so have everyone else's
    if (idle) continue;
    float s[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < 4; ++n) s[mt][n][0] = s[mt][n][1] = s[mt][n][2] = s[mt][n][3] = 0.f;

#pragma unroll
    for (int kk = 0; kk < kDh / 16; ++kk) {
      uint32_t a[2][4];
      ldsm_x4(smem_u32(sQ + (row_base + l8 + g1 * 8) * kLds + kk * 16 + g2 * 8), a[0][0], a[0][1], a[0][2], a[0][3]);
      if (second)
        ldsm_x4(smem_u32(sQ + (row_base + 16 + l8 + g1 * 8) * kLds + kk * 16 + g2 * 8), a[1][0], a[1][1], a[1][2],
                a[1][3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        if (key0 + np * 16 < C::TP) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(smem_u32(sK + (key0 + np * 16 + l8 + g2 * 8) * kLds + kk * 16 + g1 * 8), b0, b1, b2, b3);
          mma_16816(s[0][2 * np], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
          mma_16816(s[0][2 * np + 1], a[0][0], a[0][1], a[0][2], a[0][3], b2, b3);
          if (second) {
            mma_16816(s[1][2 * np], a[1][0], a[1][1], a[1][2], a[1][3], b0, b1);
            mma_16816(s[1][2 * np + 1], a[1][0], a[1][1], a[1][2], a[1][3], b2, b3);
          }
        }
      }
    }

    // ---- scale, mask, online softmax (base 2).  Chunks that hold only patch keys need no masking
    //      (decided at compile time after unrolling); the running maximum is only advanced -- and
    //      O / l rescaled -- when some row of the warp outgrew it by more than 2^8, so that in the
    //      steady state a chunk costs one FMUL + one FADD + one EX2 per score.
    const bool interior = (key0 + 32 <= P);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int n = 0; n < 4; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = key0 + n * 8 + 2 * (lane & 3) + (e & 1);
          const int hh = e >> 1;
          float v = s[mt][n][e] * scale;
          if (interior) {
            if (SIDE && is_y[mt][hh]) v += -100.0f * kLog2e * __ldg(mask + static_cast<size_t>(b) * P + col);
          } else {
            bool valid = col < C::T;  // main rows see the P patches and the class token
            if (SIDE && is_y[mt][hh]) {
              // objects.py:204-247: y sees the patches (bias -100 * mask) and itself (bias 0), not CLS
              valid = col < P || col == C::T;
              if (col < P) v += -100.0f * kLog2e * __ldg(mask + static_cast<size_t>(b) * P + col);
            }
            v = valid ? v : -INFINITY;
          }
          s[mt][n][e] = v;
          mx[hh] = fmaxf(mx[hh], v);
        }
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
      }
      const bool grow = (mx[0] > mrow[mt][0] + 8.f) || (mx[1] > mrow[mt][1] + 8.f);
      if (__any_sync(0xffffffffu, grow)) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const float m_new = fmaxf(mrow[mt][hh], mx[hh]);
          const float corr = exp2f(mrow[mt][hh] - m_new);  // first chunk: exp2(-inf) = 0
          mrow[mt][hh] = m_new;
          lrow[mt][hh] *= corr;
#pragma unroll
          for (int n = 0; n < kDh / 8; ++n) {
            o[mt][n][2 * hh] *= corr;
            o[mt][n][2 * hh + 1] *= corr;
          }
        }
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float m_use = mrow[mt][hh];
        float part = 0.f;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const float p0 = exp2f(s[mt][n][2 * hh] - m_use);
          const float p1 = exp2f(s[mt][n][2 * hh + 1] - m_use);
          s[mt][n][2 * hh] = p0;
          s[mt][n][2 * hh + 1] = p1;
          part += p0 + p1;
        }
        lrow[mt][hh] += part;
      }
    }

    // ---- O += P V for this chunk
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      if (key0 + ks * 16 < C::TP) {
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          a[mt][0] = pack2(s[mt][2 * ks][0], s[mt][2 * ks][1]);
          a[mt][1] = pack2(s[mt][2 * ks][2], s[mt][2 * ks][3]);
          a[mt][2] = pack2(s[mt][2 * ks + 1][0], s[mt][2 * ks + 1][1]);
          a[mt][3] = pack2(s[mt][2 * ks + 1][2], s[mt][2 * ks + 1][3]);
        }
#pragma unroll
        for (int np = 0; np < kDh / 16; ++np) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(smem_u32(sV + (key0 + ks * 16 + l8 + g1 * 8) * kLds + np * 16 + g2 * 8), b0, b1, b2, b3);
          mma_16816(o[0][2 * np], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
          mma_16816(o[0][2 * np + 1], a[0][0], a[0][1], a[0][2], a[0][3], b2, b3);
          if (second) {
            mma_16816(o[1][2 * np], a[1][0], a[1][1], a[1][2], a[1][3], b0, b1);
            mma_16816(o[1][2 * np + 1], a[1][0], a[1][1], a[1][2], a[1][3], b2, b3);
          }
        }
      }
    }
  }

  // ---- normalise and store
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float l = lrow[mt][hh];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const int i = row_base + mt * 16 + (lane >> 2) + hh * 8;
      const bool want = !idle && (side_only ? (SIDE && i == C::T) : (i < C::TQ));
      if (want && (mt == 0 || second)) {
        const float inv = 1.0f / l;
        act_t* dst = out + static_cast<size_t>(token_row(i, b, B, P)) * W + h * kDh + 2 * (lane & 3);
#pragma unroll
        for (int n = 0; n < kDh / 8; ++n)
          *reinterpret_cast<uint32_t*>(dst + n * 8) = pack2(o[mt][n][2 * hh] * inv, o[mt][n][2 * hh + 1] * inv);
      }
    }
  }
}

template <int P, bool SIDE>
cudaError_t launch_attn(cudaStream_t st, const act_t* qkv, const float* mask, act_t* out, int B, int heads,
                        int side_only) {
  using C = ACfg<P, SIDE>;
  if (heads % C::HPC != 0) return cudaErrorInvalidValue;
  if (cudaError_t e = ensure_dynamic_smem<attention_kernel<P, SIDE>>(C::kSmemBytes); e != cudaSuccess) return e;
  attention_kernel<P, SIDE><<<B * (heads / C::HPC), C::kThreads, C::kSmemBytes, st>>>(qkv, mask, out, B, heads,
                                                                                      side_only);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Side row only (last block of the objects tower: the main-stream output is dead, objects.py:249-258).
// One warp per (crop, head): a single query row against 196 patch keys + itself is two GEMVs, bound
// by the one pass over K and V (50 KB per warp) -- no tile machinery.
//   scores: lane = key (7 keys per lane), q in registers, each K row read as 8 x 16 B
//   softmax over the 197 scores with warp reductions, bias -100 * mask on the patches
//   output: lane = two of the 64 dims, p_j broadcast by shuffle, each V row read as one 128 B line
// ------------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(256) attention_side_row_kernel(const act_t* __restrict__ qkv,
                                                                 const float* __restrict__ mask,
                                                                 act_t* __restrict__ out, int B, int heads) {
  constexpr int NKEY = P + 1;                 // patches + the side token itself
  constexpr int KPL = (NKEY + 31) / 32;       // keys per lane
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= B * heads) return;
  const int b = warp / heads, h = warp - b * heads;
  const int W = heads * kDh;
  const size_t ld = 3 * static_cast<size_t>(W);
  const size_t y_row = static_cast<size_t>(B) * P + B + b;
  auto key_row = [&](int j) -> size_t { return j < P ? static_cast<size_t>(b) * P + j : y_row; };

  float q[kDh];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(qkv + y_row * ld + h * kDh);
#pragma unroll
    for (int c = 0; c < kDh / 8; ++c) {
      const uint4 u = __ldg(qp + c);
      const float2 a = unpack2(u.x), b2 = unpack2(u.y), c2 = unpack2(u.z), d = unpack2(u.w);
      q[8 * c + 0] = a.x; q[8 * c + 1] = a.y; q[8 * c + 2] = b2.x; q[8 * c + 3] = b2.y;
      q[8 * c + 4] = c2.x; q[8 * c + 5] = c2.y; q[8 * c + 6] = d.x; q[8 * c + 7] = d.y;
    }
  }
  float sc[KPL];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const int j = lane + 32 * i;
    sc[i] = -INFINITY;
    if (j < NKEY) {
      const uint4* kp = reinterpret_cast<const uint4*>(qkv + key_row(j) * ld + W + h * kDh);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < kDh / 8; ++c) {
        const uint4 u = __ldg(kp + c);
        const float2 a = unpack2(u.x), b2 = unpack2(u.y), c2 = unpack2(u.z), d = unpack2(u.w);
        acc = fmaf(q[8 * c + 0], a.x, acc); acc = fmaf(q[8 * c + 1], a.y, acc);
        acc = fmaf(q[8 * c + 2], b2.x, acc); acc = fmaf(q[8 * c + 3], b2.y, acc);
        acc = fmaf(q[8 * c + 4], c2.x, acc); acc = fmaf(q[8 * c + 5], c2.y, acc);
        acc = fmaf(q[8 * c + 6], d.x, acc); acc = fmaf(q[8 * c + 7], d.y, acc);
      }
      const float bias = j < P ? -100.0f * kLog2e * __ldg(mask + static_cast<size_t>(b) * P + j) : 0.f;
      sc[i] = fmaf(acc, 0.125f * kLog2e, bias);
    }
    mx = fmaxf(mx, sc[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    sc[i] = exp2f(sc[i] - mx);  // -inf (no such key) -> 0
    sum += sc[i];
  }
  sum = warp_sum(sum);
  // V: 16 independent row loads in flight per step (keys past the end re-read the last row with
  // weight 0), then the 16 weights are broadcast and accumulated
  float o0 = 0.f, o1 = 0.f;
  const uint32_t* vbase = reinterpret_cast<const uint32_t*>(qkv + 2 * W + h * kDh) + lane;
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
#pragma unroll
    for (int l0 = 0; l0 < 32; l0 += 16) {
      if (32 * i + l0 >= NKEY) break;  // compile-time after unrolling
      uint32_t u[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int j = min(32 * i + l0 + e, NKEY - 1);
        u[e] = __ldg(vbase + key_row(j) * (ld / 2));
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float pj = __shfl_sync(0xffffffffu, sc[i], l0 + e);  // 0 for keys past the end
        const float2 v = unpack2(u[e]);
        o0 = fmaf(pj, v.x, o0);
        o1 = fmaf(pj, v.y, o1);
      }
    }
  }
  const float inv = 1.0f / sum;
  reinterpret_cast<uint32_t*>(out + y_row * W + h * kDh)[lane] = pack2(o0 * inv, o1 * inv);
}

// OAKE_ATTN=mma keeps the tile kernel for the side-row-only launch too (A/B runs, tests).
bool attention_side_row_fast() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OAKE_ATTN");
    v = (e != nullptr && e[0] == 'm') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace

cudaError_t launch_attention(cudaStream_t st, const act_t* qkv, const float* mask, act_t* out, int B, int P,
                             int heads, int with_side, int side_only) {
  if (B <= 0) return cudaSuccess;
  if (attention_use_tc(P, side_only)) return launch_attention_cs(st, qkv, mask, out, B, P, heads, with_side, side_only);
  if (with_side) {
    if (P != 196 || mask == nullptr) return cudaErrorInvalidValue;
    if (side_only && attention_side_row_fast()) {
      const int warps = B * heads;
      attention_side_row_kernel<196><<<(warps + 7) / 8, 256, 0, st>>>(qkv, mask, out, B, heads);
      return cudaGetLastError();
    }
    return launch_attn<196, true>(st, qkv, mask, out, B, heads, side_only);
  }
  if (side_only) return cudaErrorInvalidValue;
  if (P == 49) return launch_attn<49, false>(st, qkv, nullptr, out, B, heads, 0);
  if (P == 196) return launch_attn<196, false>(st, qkv, nullptr, out, B, heads, 0);
  return cudaErrorInvalidValue;
}

}  // namespace oake
