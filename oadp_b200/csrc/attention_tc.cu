// tcgen05 attention for the OAKE tower (SURVEY 2.2 K4 + K7): one CTA per (crop, head, 128-row query
// tile); the math of nn.MultiheadAttention inside every CLIP ResidualAttentionBlock (reference call
// sites oadp/oake/globals.py:57, blocks.py:129, objects.py:330) plus the objects side token of
// oadp/oake/objects.py:224-247 as one more query row / key row of the same tile.
//
//   S = Q K^T      tcgen05.mma, A = Q tile (smem, K-major), B = K tile (smem, K-major), D -> TMEM
//   P = softmax    straight out of TMEM (tcgen05.ld): 16 warps, four per TMEM lane quarter, each thread
//                  owning one query row x one quarter of the keys; two passes (max, then exp / sum)
//                  with the four partial maxima / sums of a row exchanged through shared memory; P
//                  goes back into TMEM as packed fp16, aliased over the (fully consumed) S columns
//   O = P V        tcgen05.mma with A = P from TMEM, B = V tile (smem, MN-major: rows = keys), D -> TMEM
//
// TMEM columns per CTA: S [0, NK) fp32, P [0, NK/2) packed fp16 (written only after the S pieces
// underneath were consumed), O [NK/2, NK/2 + 64) (written after S is dead): 256 columns for T = 197,
// 128 for T = 50, so two to four CTAs share an SM and overlap each other's load / MMA / softmax
// phases.  K and V are loaded once per CTA by TMA (patch rows) + a few 16-byte copies (class / side
// rows, zero padding).
#include <stdlib.h>

#include "kernels.cuh"

namespace oake {

namespace {

constexpr int kDh = 64;
constexpr float kLog2e = 1.4426950408889634f;

template <int P, bool SIDE>
struct TCfg {
  static constexpr int T = P + 1;
  static constexpr int TQ = T + (SIDE ? 1 : 0);
  static constexpr int NK = ((TQ + 15) / 16) * 16;  // keys, padded to the MMA K granule
  static constexpr int MT = (TQ + 127) / 128;       // 128-row query tiles per (crop, head)
  static constexpr int kQBytes = 128 * 128;
  static constexpr int kKVRows = ((NK + 7) / 8) * 8;
  static constexpr int kKVBytes = kKVRows * 128;
  static constexpr int kXchgBytes = 2 * 4 * 128 * 4;  // partial max / sum: [2][4 groups][128 rows]
  static constexpr int kSmemBytes = 1024 + kQBytes + 2 * kKVBytes + kXchgBytes + 64;
  static constexpr int kUnits = NK / 16;  // 16-key units, dealt 4,3,3,3 (NK = 208) / 1,1,1,1 (NK = 64)
  static constexpr int kOCol = NK / 2;                                  // O accumulator columns start
  static constexpr int kTmemCols = (NK / 2 + 64 > NK ? NK / 2 + 64 : NK) <= 128 ? 128 : 256;
  static constexpr int kSoftmaxWarps = 16;
  static constexpr int kThreads = 32 * (kSoftmaxWarps + 1);  // + 1 control warp (TMA, MMA issue)
};

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]: A = 128 x 16 packed 16-bit values (lane = row, 8 columns)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// MN-major operand (rows = K index, each row = 64 contiguous MN elements = one 128-byte swizzle
// span, 8-row groups 1024 B apart): the image TMA writes for a {64, rows} box with SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO: stride between 64-element MN blocks (single block)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: stride between 8-row K groups
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_f16_bmn(int m, int n) {
  return make_idesc_f16(m, n) | (1u << 16);  // B is MN-major
}

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const int bytes = valid ? 16 : 0;  // src-size 0: the destination is zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes)
               : "memory");
}

// 16-byte chunk `c` of row `r` inside a 128B-swizzled tile (tile base 1024-aligned)
__device__ __forceinline__ uint8_t* sw128(uint8_t* tile, int r, int c) { return tile + r * 128 + ((c ^ (r & 7)) << 4); }

__device__ __forceinline__ int token_row(int i, int b, int B, int P) {
  return i < P ? b * P + i : (i == P ? B * P + b : B * P + B + b);
}

template <int P, bool SIDE>
__global__ void __launch_bounds__(TCfg<P, SIDE>::kThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ,   // qkv [R, 3W], box {64, 128}
                    const __grid_constant__ CUtensorMap tmKV,  // qkv [R, 3W], box {64, P}
                    const act_t* __restrict__ qkv, const float* __restrict__ mask, act_t* __restrict__ out,
                    int B, int heads, int mt_first, int only_y) {
  using C = TCfg<P, SIDE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::kQBytes;
  uint8_t* sV = sK + C::kKVBytes;
  float* s_max = reinterpret_cast<float*>(sV + C::kKVBytes);  // [4][128]
  float* s_sum = s_max + 4 * 128;                              // [4][128]
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(s_sum + 4 * 128);
  uint64_t* bar_s = bar_load + 1;
  uint64_t* bar_p = bar_load + 2;
  uint64_t* bar_o = bar_load + 3;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_load + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mts = C::MT - mt_first;  // query tiles per (crop, head) handled by this launch
  const int item = blockIdx.x / mts;
  const int mt = mt_first + (blockIdx.x - item * mts);
  const int b = item / heads;
  const int h = item - b * heads;
  const int W = heads * kDh;
  const int ld = 3 * W;

  constexpr int kCtl = C::kSoftmaxWarps;  // index of the control warp
  if (warp == kCtl && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, C::kSoftmaxWarps);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<C::kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == kCtl) {
    // ------------------------------------------------------------------ loads + MMA issue
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_load, C::kQBytes + 2 * P * 128);
      // Q: 128 token rows starting at token mt*128 (all of them patch rows of this crop, or the
      // rows that follow them in memory -- finite garbage that only feeds rows nobody stores)
      tma_load_2d(sQ, &tmQ, bar_load, h * kDh, b * P + mt * 128);
      tma_load_2d(sK, &tmKV, bar_load, W + h * kDh, b * P);
      tma_load_2d(sV, &tmKV, bar_load, 2 * W + h * kDh, b * P);
    }
    // class / side rows and zero padding of K and V (rows the TMA box does not touch): cp.async, so
    // that all of them are in flight together with the TMA boxes
    for (int idx = lane; idx < (C::kKVRows - P) * 8; idx += 32) {
      const int i = P + (idx >> 3), c = idx & 7;
      const bool ok = i < C::TQ;
      const act_t* src = qkv + static_cast<size_t>(token_row(ok ? i : P, b, B, P)) * ld + h * kDh + c * 8;
      cp_async16_zfill(sw128(sK, i, c), src + W, ok);
      cp_async16_zfill(sw128(sV, i, c), src + 2 * W, ok);
    }
    // class / side query rows: they overwrite rows of the Q box, so they are staged in registers and
    // written once the box has landed
    uint4 qrow[(C::TQ - P) * 8 / 32 + 1];
#pragma unroll
    for (int it = 0; it < (C::TQ - P) * 8 / 32 + 1; ++it) {
      const int idx = lane + 32 * it;
      qrow[it] = make_uint4(0, 0, 0, 0);
      if (idx < (C::TQ - P) * 8) {
        const int i = P + (idx >> 3), c = idx & 7;
        const int r = i - mt * 128;
        if (r >= 0 && r < 128)
          qrow[it] = *reinterpret_cast<const uint4*>(qkv + static_cast<size_t>(token_row(i, b, B, P)) * ld + h * kDh + c * 8);
      }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    mbar_wait(bar_load, 0);
#pragma unroll
    for (int it = 0; it < (C::TQ - P) * 8 / 32 + 1; ++it) {
      const int idx = lane + 32 * it;
      if (idx < (C::TQ - P) * 8) {
        const int i = P + (idx >> 3), c = idx & 7;
        const int r = i - mt * 128;
        if (r >= 0 && r < 128) *reinterpret_cast<uint4*>(sw128(sQ, r, c)) = qrow[it];
      }
    }
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core
    __syncwarp();
    tc_fence_after();
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, C::NK);
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK);
#pragma unroll
      for (int k = 0; k < kDh / 16; ++k)
        umma_f16(tmem_base, make_smem_desc_k_sw128(q_addr + k * 32), make_smem_desc_k_sw128(k_addr + k * 32),
                 idesc_s, k != 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_p, 0);
    tc_fence_after();
    if (elect_one()) {
      constexpr uint32_t idesc_o = make_idesc_f16_bmn(128, kDh);
      const uint32_t v_addr = smem_u32(sV);
#pragma unroll
      for (int k = 0; k < C::NK / 16; ++k)
        umma_f16_ts(tmem_base + C::kOCol, tmem_base + k * 8, make_smem_desc_mn_sw128(v_addr + k * 2048), idesc_o,
                    k != 0 ? 1u : 0u);
      umma_commit(bar_o);
    }
  } else {
    // ------------------------------------------------------------------ softmax
    // thread = (query row r, key group g): warps q, q+4, q+8, q+12 all read TMEM lane quarter q
    const int q = warp & 3;
    const int g = warp >> 2;
    const int r = q * 32 + lane;  // row of this tile
    const int i = mt * 128 + r;   // token index
    const bool is_y = SIDE && i == C::T;
    const bool live = only_y ? is_y : (i < C::TQ);  // only_y: last objects block, just the side row
    // 16-key units of this group: the first (kUnits % 4) groups get one more
    const int units = C::kUnits / 4 + (g < C::kUnits % 4 ? 1 : 0);
    const int unit0 = g * (C::kUnits / 4) + (g < C::kUnits % 4 ? g : C::kUnits % 4);
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float scale = 0.125f * kLog2e;
    const float* mrow = mask + static_cast<size_t>(b) * P;
    constexpr int kMaxUnits = (C::kUnits + 3) / 4;
    mbar_wait(bar_s, 0);
    tc_fence_after();

    // Units whose 16 keys are all patches need no masking for a main-stream row: one FMNMX per
    // score in pass 1 (on the raw score: the scale is positive), FFMA + EX2 + FADD (+ half a pack)
    // in pass 2.  The unit that holds the class / side / padding keys and the single side-stream
    // row of a tile (bias -100 * mask, different key set) take the generic masked path.
    auto generic = [&](float raw, int col) -> float {  // scaled, biased, masked score
      float v = raw * scale;
      bool valid = col < C::T;
      if (is_y) {  // objects.py:204-247: patches with bias -100 * mask, itself, not the class token
        valid = col < P || col == C::T;
        if (col < P) v += -100.0f * kLog2e * __ldg(mrow + col);
      }
      return valid ? v : -INFINITY;
    };

    // pass 1: partial row maximum (base-2 domain)
    float mx_raw = -INFINITY, mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < kMaxUnits; ++u) {
      if (u < units) {
        uint32_t r16[16];
        tmem_ld_32x16(t_row + (unit0 + u) * 16, r16);
        tmem_ld_wait();
        if (live) {
          const int col0 = (unit0 + u) * 16;
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          if (col0 + 16 <= P && !is_y) {
#pragma unroll
            for (int j = 0; j < 16; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r16[j]));
            mx_raw = fmaxf(mx_raw, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) m4[j & 3] = fmaxf(m4[j & 3], generic(__uint_as_float(r16[j]), col0 + j));
            mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
          }
        }
      }
    }
    mx = fmaxf(mx, mx_raw * scale);
    s_max[g * 128 + r] = mx;
    tc_fence_before();
    asm volatile("bar.sync 1, %0;\n" ::"n"(C::kSoftmaxWarps * 32) : "memory");
    tc_fence_after();
    mx = fmaxf(fmaxf(s_max[r], s_max[128 + r]), fmaxf(s_max[256 + r], s_max[384 + r]));
    const float neg_mx = -mx;

    // pass 2: p = exp2(s - max), partial row sum; P is kept in registers until every warp has
    // finished reading S (the packed P columns alias S columns owned by other key groups)
    float sum = 0.f;
    uint32_t pk[kMaxUnits][8];
#pragma unroll
    for (int u = 0; u < kMaxUnits; ++u) {
      if (u < units) {
        uint32_t r16[16];
        tmem_ld_32x16(t_row + (unit0 + u) * 16, r16);
        tmem_ld_wait();
        const int col0 = (unit0 + u) * 16;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        if (!live) {
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[u][j] = 0u;
        } else if (col0 + 16 <= P && !is_y) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float p0 = exp2f(fmaf(__uint_as_float(r16[2 * j]), scale, neg_mx));
            const float p1 = exp2f(fmaf(__uint_as_float(r16[2 * j + 1]), scale, neg_mx));
            s4[j & 3] += p0 + p1;
            pk[u][j] = pack2(p0, p1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float p0 = exp2f(generic(__uint_as_float(r16[2 * j]), col0 + 2 * j) + neg_mx);
            const float p1 = exp2f(generic(__uint_as_float(r16[2 * j + 1]), col0 + 2 * j + 1) + neg_mx);
            s4[j & 3] += p0 + p1;
            pk[u][j] = pack2(p0, p1);
          }
        }
        sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
    }
    s_sum[g * 128 + r] = sum;
    tc_fence_before();
    asm volatile("bar.sync 1, %0;\n" ::"n"(C::kSoftmaxWarps * 32) : "memory");
    tc_fence_after();
#pragma unroll
    for (int u = 0; u < kMaxUnits; ++u) {
      if (u < units)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(
                         t_row + (unit0 + u) * 8),
                     "r"(pk[u][0]), "r"(pk[u][1]), "r"(pk[u][2]), "r"(pk[u][3]), "r"(pk[u][4]), "r"(pk[u][5]),
                     "r"(pk[u][6]), "r"(pk[u][7])
                     : "memory");
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    sum = (s_sum[r] + s_sum[128 + r]) + (s_sum[256 + r] + s_sum[384 + r]);

    // O: each key group stores 16 of the row's 64 output values (32 contiguous bytes)
    mbar_wait(bar_o, 0);
    tc_fence_after();
    {
      const float inv = 1.0f / sum;
      uint32_t o16[16];
      tmem_ld_32x16(t_row + C::kOCol + g * 16, o16);
      tmem_ld_wait();
      if (live) {
        act_t* dst = out + static_cast<size_t>(token_row(i, b, B, P)) * W + h * kDh + g * 16;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 u;
          u.x = pack2(__uint_as_float(o16[8 * j + 0]) * inv, __uint_as_float(o16[8 * j + 1]) * inv);
          u.y = pack2(__uint_as_float(o16[8 * j + 2]) * inv, __uint_as_float(o16[8 * j + 3]) * inv);
          u.z = pack2(__uint_as_float(o16[8 * j + 4]) * inv, __uint_as_float(o16[8 * j + 5]) * inv);
          u.w = pack2(__uint_as_float(o16[8 * j + 6]) * inv, __uint_as_float(o16[8 * j + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + j * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int P, bool SIDE>
cudaError_t launch_tc(cudaStream_t st, const act_t* qkv, const float* mask, act_t* out, int B, int heads, int rows,
                      int mt_first, int only_y) {
  using C = TCfg<P, SIDE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<P, SIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  CUtensorMap tmQ, tmKV;
  if (make_tmap_act_2d(&tmQ, qkv, rows, 3 * heads * kDh, 128) || make_tmap_act_2d(&tmKV, qkv, rows, 3 * heads * kDh, P))
    return cudaErrorInvalidValue;
  const int grid = B * heads * (C::MT - mt_first);
  attention_tc_kernel<P, SIDE><<<grid, C::kThreads, C::kSmemBytes, st>>>(tmQ, tmKV, qkv, mask, out, B, heads, mt_first,
                                                                         only_y);
  return cudaGetLastError();
}

}  // namespace

// Opt-in (OAKE_ATTN=tc).  This non-persistent form is numerically verified but latency-bound: each
// CTA runs load -> S -> softmax -> PV -> store serially and TMEM (256 columns per 128-row tile)
// allows only two tiles in flight per SM; the mma.sync kernel of attention.cu is the default until
// this one is made persistent with ping-pong tiles.
bool attention_use_tc(int P) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OAKE_ATTN");
    v = (e == nullptr) ? 2 : (e[0] == 'm' ? 0 : 1);
  }
  (void)P;
  return v == 1;  // measured r1: mma.sync 4.6 ms vs tcgen05 5.8 ms per 478-crop objects batch
}

cudaError_t launch_attention_tc(cudaStream_t st, const act_t* qkv, const float* mask, act_t* out, int B, int P,
                                int heads, int with_side, int side_only) {
  if (B <= 0) return cudaSuccess;
  const int rows = B * (P + 1) + (with_side ? B : 0);
  if (with_side) {
    if (P != 196 || mask == nullptr) return cudaErrorInvalidValue;
    // side_only: the y row lives in the second query tile; the first one is not needed at all
    return launch_tc<196, true>(st, qkv, mask, out, B, heads, rows, side_only ? 1 : 0, side_only ? 1 : 0);
  }
  if (side_only) return cudaErrorInvalidValue;
  if (P == 49) return launch_tc<49, false>(st, qkv, nullptr, out, B, heads, rows, 0, 0);
  if (P == 196) return launch_tc<196, false>(st, qkv, nullptr, out, B, heads, rows, 0, 0);
  return cudaErrorInvalidValue;
}

}  // namespace oake
