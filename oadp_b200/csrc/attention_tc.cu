// Persistent tcgen05 attention for the 197-token OAKE tower (SURVEY 2.2 K4 + K7): the math of
// nn.MultiheadAttention inside every CLIP ResidualAttentionBlock (reference call site
// oadp/oake/objects.py:330) plus the objects side token of oadp/oake/objects.py:224-247 as one more
// query row / key row of the same tile.
//
// One CTA per SM walks over (crop, head) items.  An item is two 128-row query tiles (tokens 0..127 and
// 128..197) against one 208-key K/V tile.  Warp roles:
//
//   warps 0-3   softmax group A: query tile 0 of every item, TMEM columns [0, 256)
//   warps 4-7   softmax group B: query tile 1 of every item, TMEM columns [256, 512)
//   warp  8     loader: TMA boxes for Q / K / V (+ 16-byte copies for the class / side rows, zero
//               padding, and the side row's additive bias) into a two-stage shared-memory ring
//   warp  9     MMA issue (one thread): S = Q K^T and O = P V, interleaved PV(n,t), S(n+1,t)
//
// A softmax thread owns one query row (TMEM lane) for all 208 keys, so a row needs no exchange with
// other threads: pass 1 reads S for the row maximum, pass 2 reads S again, writes P = exp2(s - max)
// as packed fp16 over the S columns it has already consumed and accumulates the row sum; O lands in
// the dead S columns behind P and is scaled by 1 / sum on its way to global memory.  While group A
// is in its exp pass, the tensor core runs group B's MMAs and vice versa (the MUFU pipe is the
// bound: 197 x 197 exponentials per head against 2 x 128 x 208 x 64 MACs).
//
// Tile 1 has only 70 live rows.  Warp w always runs on SM sub-partition w % 4 and owns TMEM lanes
// 32 (w % 4) .. +31, so a fixed placement would leave sub-partition 3 idle for tile 1; odd items
// therefore place the 70 rows at lanes 56..125 instead of 0..69.
#include <stdlib.h>

#include "attn_tc_common.cuh"

namespace oake {

namespace {

using namespace attn;

template <bool SIDE>
struct PCfg {
  static constexpr int P = 196;
  static constexpr int T = P + 1;
  static constexpr int TQ = T + (SIDE ? 1 : 0);
  static constexpr int NK = 208;                 // keys, padded to the MMA K granule (16)
  static constexpr int kUnits = NK / 16;         // 13
  static constexpr int kRows1 = TQ - 128;        // live rows of tile 1: 68 patches + class (+ side)
  static constexpr int kPatch1 = P - 128;        // patch rows of tile 1 (one TMA box)
  static constexpr int kShift = 56;              // lane offset of tile 1's rows on odd items
  static constexpr int kQTile = 128 * 128;       // bytes: 128 rows x 64 halves
  static constexpr int kKV = NK * 128;           // bytes
  static constexpr int kStage = 2 * kQTile + 2 * kKV;
  static constexpr int kOCol = NK / 2;           // two O accumulators (even / odd key units) behind the packed P
  static constexpr int kMaskFloats = 208;        // staged mask row (196 used), 16-byte granules
  static constexpr int kBufCols = 256;           // TMEM columns per softmax group
  static constexpr int kNumBars = 16;
  static constexpr int kMaskBytes = 2048;        // both stages' mask rows, padded
  static constexpr int kOutStage = 32 * 128;     // bytes per softmax warp: 32 rows x 64 halves, for coalesced stores
  static constexpr int kSmemBytes = 1024 + 2 * kStage + kMaskBytes + 8 * kOutStage + kNumBars * 8 + 16;
  static constexpr int kSoftmaxWarps = 8;
  static constexpr int kLoaderWarp = 8;
  static constexpr int kMmaWarp = 9;
  static constexpr int kThreads = 32 * 10;
};

// How a softmax warp biases its scores.
//   kPlain: main-stream rows only.  Keys 0..196 (patches, class) valid, no bias: pass 1 is a plain
//           maximum of the raw scores.
//   kBits / kLoad: the warp that holds the side row (warp-uniform code, per-lane select).  The side
//           row adds -100 log2e mask[key] on patches, -inf on the class key, 0 on its own key
//           (objects.py:204-247); main-stream rows add 0 / -inf by key validity.  kBits takes the mask
//           from 7 words of bits (the reference's masks are 0 / 1 by construction, objects.py:147-153);
//           kLoad reads the fp32 mask row for any other mask values.
enum BiasMode { kPlain = 0, kBits = 1, kLoad = 2 };

// One query row: S (fp32, TMEM columns [0, NK) of this thread's lane) -> P (packed fp16, columns
// [0, NK/2)); returns the row sum of the unrounded exponentials.  Keys are walked in three 64-key
// steps (all patches: one loop body) and a 16-key tail (patches 192..195, class, side, padding).
template <bool SIDE, int MODE>
__device__ __forceinline__ float softmax_row(uint32_t t_row, bool is_y, uint32_t ymask_addr, uint32_t ybits) {
  using C = PCfg<SIDE>;
  constexpr float scale = 0.125f * kLog2e;
  constexpr float kNegB = -100.0f * kLog2e;
  constexpr int kMain = 192;  // keys handled by the loop
  static_assert(kMain + 16 == C::NK && kMain <= C::P, "tail layout");
  // bias of a patch key (MODE != kPlain); `w` = this lane's mask bits of the key's 32-key group
  auto patch_bias = [&](int col, int j, uint32_t w) -> float {
    if (MODE == kBits) return (w >> j) & 1u ? kNegB : 0.f;
    return is_y ? kNegB * lds_f32(ymask_addr + col * 4) : 0.f;
  };
  // lane g < 7 holds the bits of keys 32g..32g+31 in `ybits`; rows other than the side row see zeros
  auto group_bits = [&](int g) -> uint32_t {
    if (MODE != kBits) return 0u;
    const uint32_t w = __shfl_sync(0xffffffffu, ybits, g);
    return is_y ? w : 0u;
  };
  // bias of a tail key 192 + j
  auto tail_bias = [&](int j, uint32_t w) -> float {
    const int col = kMain + j;
    if (col < C::P) return MODE == kPlain ? 0.f : patch_bias(col, j, w);
    const float main_b = col < C::T ? 0.f : -INFINITY;                 // class key valid, side key / padding not
    if (MODE == kPlain) return main_b;
    const float y_b = col == C::T ? 0.f : -INFINITY;                    // the side token sees itself, not the class key
    return is_y ? y_b : main_b;
  };

  // ---- pass 1: row maximum (base-2 domain)
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
  for (int c = 0; c < kMain / 64; ++c) {
    uint32_t ra[32], rb[32];
    tmem_ld_32x32(t_row + c * 64, ra);
    tmem_ld_32x32(t_row + c * 64 + 32, rb);
    const uint32_t wa = group_bits(2 * c), wb = group_bits(2 * c + 1);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (MODE == kPlain) {
        m4[j & 3] = fmaxf(m4[j & 3], fmaxf(__uint_as_float(ra[j]), __uint_as_float(rb[j])));
      } else {
        m4[j & 3] = fmaxf(m4[j & 3], fmaf(__uint_as_float(ra[j]), scale, patch_bias(c * 64 + j, j, wa)));
        m4[j & 3] = fmaxf(m4[j & 3], fmaf(__uint_as_float(rb[j]), scale, patch_bias(c * 64 + 32 + j, j, wb)));
      }
    }
  }
  float mx;
  {
    uint32_t rt[16];
    tmem_ld_32x16(t_row + kMain, rt);
    const uint32_t wt = group_bits(6);
    tmem_ld_wait();
    if (MODE == kPlain) {  // the scale is positive: max(s) * scale == max(s * scale)
      mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale;
    } else {
      mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (kMain + j <= C::T) mx = fmaxf(mx, fmaf(__uint_as_float(rt[j]), scale, tail_bias(j, wt)));
  }
  const float neg_mx = -mx;

  // ---- pass 2: p = exp2(s - max) -> packed fp16 over the consumed S columns; row sum
  // (bias added after the fma: a main-stream row gets bit-identical results in every mode, so batch
  // composition stays invisible)
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int c = 0; c < kMain / 64; ++c) {
    uint32_t ra[32], rb[32];
    tmem_ld_32x32(t_row + c * 64, ra);
    tmem_ld_32x32(t_row + c * 64 + 32, rb);
    const uint32_t wa = group_bits(2 * c), wb = group_bits(2 * c + 1);
    tmem_ld_wait();
    uint32_t pa[16], pb[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float p[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jj = 2 * j + e;
        if (MODE == kPlain) {
          p[e] = ex2(fmaf(__uint_as_float(ra[jj]), scale, neg_mx));
          p[2 + e] = ex2(fmaf(__uint_as_float(rb[jj]), scale, neg_mx));
        } else {
          p[e] = ex2(fmaf(__uint_as_float(ra[jj]), scale, neg_mx) + patch_bias(c * 64 + jj, jj, wa));
          p[2 + e] = ex2(fmaf(__uint_as_float(rb[jj]), scale, neg_mx) + patch_bias(c * 64 + 32 + jj, jj, wb));
        }
      }
      s4[j & 3] += (p[0] + p[1]) + (p[2] + p[3]);
      pa[j] = pack2(p[0], p[1]);
      pb[j] = pack2(p[2], p[3]);
    }
    tmem_st_32x16(t_row + c * 32, pa);
    tmem_st_32x16(t_row + c * 32 + 16, pb);
  }
  {
    uint32_t rt[16];
    tmem_ld_32x16(t_row + kMain, rt);
    const uint32_t wt = group_bits(6);
    tmem_ld_wait();
    uint32_t pt[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jj = 2 * j + e;
        p[e] = kMain + jj <= C::T ? ex2(fmaf(__uint_as_float(rt[jj]), scale, neg_mx) + tail_bias(jj, wt)) : 0.f;
      }
      s4[j & 3] += p[0] + p[1];
      pt[j] = pack2(p[0], p[1]);
    }
    tmem_st_32x8(t_row + kMain / 2, pt);
  }
  return (s4[0] + s4[1]) + (s4[2] + s4[3]);
}

template <bool SIDE>
__global__ void __launch_bounds__(PCfg<SIDE>::kThreads, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tmQ0,  // qkv [R, 3W], box {64, 128}
                    const __grid_constant__ CUtensorMap tmQ1,  // qkv [R, 3W], box {64, 68}
                    const __grid_constant__ CUtensorMap tmKV,  // qkv [R, 3W], box {64, 196}
                    const act_t* __restrict__ qkv, const float* __restrict__ mask, act_t* __restrict__ out,
                    int B, int heads) {
  using C = PCfg<SIDE>;
  constexpr int P = C::P;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* ymask = reinterpret_cast<float*>(smem + 2 * C::kStage);  // [2][kMaskFloats]: mask row of the item's crop
  uint8_t* out_stage = smem + 2 * C::kStage + C::kMaskBytes;  // [8 warps][kOutStage]
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + 8 * C::kOutStage);
  uint64_t* qk_full = bars + 0;   // [stage]  loader -> MMA
  uint64_t* qk_free = bars + 2;   // [stage]  MMA (S of both tiles retired) -> loader
  uint64_t* v_full = bars + 4;    // [stage]  loader -> MMA, side-row warp (mask row)
  uint64_t* v_free = bars + 6;    // [stage]  MMA (PV of both tiles retired) -> loader
  uint64_t* s_full = bars + 8;    // [tile]   MMA -> softmax group
  uint64_t* p_ready = bars + 10;  // [tile]   softmax group -> MMA
  uint64_t* o_full = bars + 12;   // [tile]   MMA -> softmax group
  uint64_t* o_free = bars + 14;   // [tile]   softmax group (O drained) -> MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + C::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int W = heads * kDh;
  const int ld = 3 * W;
  const int items = B * heads;
  const int N = (items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == C::kLoaderWarp && lane == 0) {
    tma_prefetch_desc(&tmQ0);
    tma_prefetch_desc(&tmQ1);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 33);  // expect_tx arrival + one cp.async arrival per loader lane
      mbar_init(&qk_free[i], 1);
      mbar_init(&v_full[i], 33);
      mbar_init(&v_free[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);  // one arrival per warp of the group
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == C::kLoaderWarp) {
    // ================================================================== loader
    for (int n = 0; n < N; ++n) {
      const int s = n & 1, u = n >> 1;
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / heads, h = item - b * heads;
      uint8_t* stg = smem + s * C::kStage;
      uint8_t* sQ0 = stg;
      uint8_t* sQ1 = stg + C::kQTile;
      uint8_t* sK = stg + 2 * C::kQTile;
      uint8_t* sV = sK + C::kKV;
      const int shift = s ? C::kShift : 0;

      mbar_wait(&qk_free[s], (u & 1) ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&qk_full[s], C::kQTile + C::kPatch1 * 128 + P * 128);
        tma_load_2d(sQ0, &tmQ0, &qk_full[s], h * kDh, b * P);
        tma_load_2d(sQ1 + shift * 128, &tmQ1, &qk_full[s], h * kDh, b * P + 128);
        tma_load_2d(sK, &tmKV, &qk_full[s], W + h * kDh, b * P);
      }
      __syncwarp();
      // K rows of the class / side token and the zero padding up to NK
      for (int idx = lane; idx < (C::NK - P) * 8; idx += 32) {
        const int i = P + (idx >> 3), c = idx & 7;
        const bool ok = i < C::TQ;
        const act_t* src = qkv + static_cast<size_t>(token_row(ok ? i : P, b, B, P)) * ld + W + h * kDh + c * 8;
        cp_async16_zfill(sw128(sK, i, c), src, ok);
      }
      // Q rows of the class / side token: tile-1 rows shift + 68 (, + 69)
      for (int idx = lane; idx < (C::TQ - P) * 8; idx += 32) {
        const int i = P + (idx >> 3), c = idx & 7;
        const act_t* src = qkv + static_cast<size_t>(token_row(i, b, B, P)) * ld + h * kDh + c * 8;
        cp_async16_zfill(sw128(sQ1, shift + i - 128, c), src, true);
      }
      cp_async_arrive_noinc(&qk_full[s]);  // every lane, with or without copies of its own

      mbar_wait(&v_free[s], (u & 1) ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&v_full[s], P * 128);
        tma_load_2d(sV, &tmKV, &v_full[s], 2 * W + h * kDh, b * P);
      }
      __syncwarp();
      for (int idx = lane; idx < (C::NK - P) * 8; idx += 32) {
        const int i = P + (idx >> 3), c = idx & 7;
        const bool ok = i < C::TQ;
        const act_t* src = qkv + static_cast<size_t>(token_row(ok ? i : P, b, B, P)) * ld + 2 * W + h * kDh + c * 8;
        cp_async16_zfill(sw128(sV, i, c), src, ok);
      }
      if (SIDE) {  // the crop's mask row (196 floats = 49 granules) for the side-row warp
        const float* mrow = mask + static_cast<size_t>(b) * P;
        for (int g = lane; g < P / 4; g += 32) cp_async16_zfill(ymask + s * C::kMaskFloats + g * 4, mrow + g * 4, true);
      }
      cp_async_arrive_noinc(&v_full[s]);
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");  // nothing of this warp in flight at exit
  } else if (warp == C::kMmaWarp) {
    // ================================================================== MMA issue
    constexpr uint32_t idesc_s = make_idesc_f16(128, C::NK);
    constexpr uint32_t idesc_o = make_idesc_f16_bmn(128, kDh);
    const uint32_t smem_base = smem_u32(smem);
    auto issue_s = [&](int n, int t) {  // lane 0 only
      const uint32_t stg = smem_base + (n & 1) * C::kStage;
      const uint32_t q_addr = stg + t * C::kQTile, k_addr = stg + 2 * C::kQTile;
#pragma unroll
      for (int k = 0; k < kDh / 16; ++k)
        umma_f16(tmem_base + t * C::kBufCols, make_smem_desc_k_sw128(q_addr + k * 32),
                 make_smem_desc_k_sw128(k_addr + k * 32), idesc_s, k != 0 ? 1u : 0u);
      umma_commit(&s_full[t]);
    };
    auto issue_pv = [&](int n, int t) {  // lane 0 only
      const uint32_t v_addr = smem_base + (n & 1) * C::kStage + 2 * C::kQTile + C::kKV;
      const uint32_t buf = tmem_base + t * C::kBufCols;
#pragma unroll
      for (int k = 0; k < C::kUnits; ++k)
        umma_f16_ts(buf + C::kOCol + (k & 1) * kDh, buf + k * 8, make_smem_desc_mn_sw128(v_addr + k * 2048), idesc_o,
                    k >= 2 ? 1u : 0u);  // even / odd key units accumulate separately: two short chains
      umma_commit(&o_full[t]);
    };
    if (N > 0) {
      mbar_wait(&qk_full[0], 0);
      fence_proxy_async();  // the loader's cp.async rows (generic proxy) -> tensor core reads
      tc_fence_after();
      if (lane == 0) {
        issue_s(0, 0);
        issue_s(0, 1);
        umma_commit(&qk_free[0]);
      }
      __syncwarp();
    }
    for (int n = 0; n < N; ++n) {
      const int s = n & 1, u = n >> 1;
      mbar_wait(&v_full[s], u & 1);
      fence_proxy_async();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_ready[t], n & 1);
        tc_fence_after();
        if (lane == 0) {
          issue_pv(n, t);
          if (t == 1) umma_commit(&v_free[s]);
        }
        __syncwarp();
        if (n + 1 < N) {
          if (t == 0) {
            mbar_wait(&qk_full[s ^ 1], ((n + 1) >> 1) & 1);
            fence_proxy_async();
          }
          mbar_wait(&o_free[t], n & 1);
          tc_fence_after();
          if (lane == 0) {
            issue_s(n + 1, t);
            if (t == 1) umma_commit(&qk_free[s ^ 1]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================================== softmax groups
    const int t = warp >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;  // lane of the tile
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * C::kBufCols;
    for (int n = 0; n < N; ++n) {
      const int s = n & 1, u = n >> 1;
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / heads, h = item - b * heads;
      const int rr = r - ((t == 1 && s) ? C::kShift : 0);
      const bool live = t == 0 || (rr >= 0 && rr < C::kRows1);
      const int i = t * 128 + rr;  // token index
      const bool is_y = SIDE && live && i == C::T;
      const bool warp_live = __any_sync(0xffffffffu, live);
      const bool warp_y = SIDE && __any_sync(0xffffffffu, is_y);

      mbar_wait(&s_full[t], n & 1);
      tc_fence_after();
      float sum = 1.f;
      if (warp_live) {
        if (warp_y) {
          mbar_wait(&v_full[s], u & 1);  // the crop's mask row lands with the V stage
          const uint32_t ymask_addr = smem_u32(ymask + s * C::kMaskFloats);
          uint32_t ybits = 0u, other = 0u;  // lane g keeps the bits of keys 32g .. 32g+31
#pragma unroll
          for (int g = 0; g < 7; ++g) {
            const int col = g * 32 + lane;
            const float m = col < P ? lds_f32(ymask_addr + col * 4) : 0.f;
            const uint32_t w = __ballot_sync(0xffffffffu, m != 0.f);
            other |= __ballot_sync(0xffffffffu, m != 0.f && m != 1.f);
            if (lane == g) ybits = w;
          }
          if (other == 0u) {
            sum = softmax_row<SIDE, kBits>(t_row, is_y, ymask_addr, ybits);
          } else {
            sum = softmax_row<SIDE, kLoad>(t_row, is_y, ymask_addr, 0u);
          }
        } else {
          sum = softmax_row<SIDE, kPlain>(t_row, false, 0u, 0u);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[t]);

      if (warp_live) {
        uint32_t oa0[32], oa1[32], ob0[32], ob1[32];
        mbar_wait(&o_full[t], n & 1);
        tc_fence_after();
        tmem_ld_32x32(t_row + C::kOCol, oa0);
        tmem_ld_32x32(t_row + C::kOCol + 32, oa1);
        tmem_ld_32x32(t_row + C::kOCol + kDh, ob0);
        tmem_ld_32x32(t_row + C::kOCol + kDh + 32, ob1);
        tmem_ld_wait();
        // the accumulators are in registers: the tile's TMEM goes back to the tensor core before
        // the scaling and the global stores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[t]);
        // rows -> shared memory (one 128-byte row per lane, 16-byte chunks XOR-swizzled), then every
        // store instruction writes four whole 128-byte rows of the output
        const uint32_t stg = smem_u32(out_stage + warp * C::kOutStage);
        {
          const float inv = 1.0f / sum;
          auto o_at = [&](int c) -> float {
            return c < 32 ? __uint_as_float(oa0[c & 31]) + __uint_as_float(ob0[c & 31])
                          : __uint_as_float(oa1[c & 31]) + __uint_as_float(ob1[c & 31]);
          };
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 v;
            v.x = pack2(o_at(8 * j + 0) * inv, o_at(8 * j + 1) * inv);
            v.y = pack2(o_at(8 * j + 2) * inv, o_at(8 * j + 3) * inv);
            v.z = pack2(o_at(8 * j + 4) * inv, o_at(8 * j + 5) * inv);
            v.w = pack2(o_at(8 * j + 6) * inv, o_at(8 * j + 7) * inv);
            sts128(stg + lane * 128 + ((j ^ (lane & 7)) << 4), v);
          }
        }
        __syncwarp();
        {
          const int chunk = lane & 7;
          const int shift1 = (t == 1 && s) ? C::kShift : 0;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int row = 4 * k + (lane >> 3);          // row of this warp's 32
            const int rr2 = q * 32 + row - shift1;        // row of the tile's live range
            const bool live2 = t == 0 || (rr2 >= 0 && rr2 < C::kRows1);
            if (live2) {
              const uint4 v = lds128(stg + row * 128 + ((chunk ^ (row & 7)) << 4));
              const int tok = t * 128 + rr2;
              *reinterpret_cast<uint4*>(out + static_cast<size_t>(token_row(tok, b, B, P)) * W + h * kDh + chunk * 8) = v;
            }
          }
        }
        __syncwarp();
      } else {
        if (lane == 0) mbar_arrive(&o_free[t]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <bool SIDE>
cudaError_t launch_pp(cudaStream_t st, const act_t* qkv, const float* mask, act_t* out, int B, int heads, int rows) {
  using C = PCfg<SIDE>;
  static bool attr_set = false;
  static int num_sms = 0;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_pp_kernel<SIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kSmemBytes);
    if (e != cudaSuccess) return e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    attr_set = true;
  }
  CUtensorMap tmQ0, tmQ1, tmKV;
  const uint64_t cols = 3ull * heads * kDh;
  if (make_tmap_act_2d(&tmQ0, qkv, rows, cols, 128) || make_tmap_act_2d(&tmQ1, qkv, rows, cols, C::kPatch1) ||
      make_tmap_act_2d(&tmKV, qkv, rows, cols, C::P))
    return cudaErrorInvalidValue;
  const int items = B * heads;
  const int grid = items < num_sms ? items : num_sms;
  attention_pp_kernel<SIDE><<<grid, C::kThreads, C::kSmemBytes, st>>>(tmQ0, tmQ1, tmKV, qkv, mask, out, B, heads);
  return cudaGetLastError();
}

}  // namespace

// The persistent tcgen05 kernel serves the 197-token tower whenever whole tiles are wanted; the
// 50-token tower and the last objects block (side row only) stay on the mma.sync kernel of
// attention.cu.  OAKE_ATTN=mma forces the mma.sync kernel everywhere (A/B runs, tests).
bool attention_use_tc(int P, int side_only) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OAKE_ATTN");
    v = (e != nullptr && e[0] == 'm') ? 0 : 1;
  }
  return v == 1 && P == 196 && !side_only;
}

cudaError_t launch_attention_tc(cudaStream_t st, const act_t* qkv, const float* mask, act_t* out, int B, int P,
                                int heads, int with_side, int side_only) {
  if (B <= 0) return cudaSuccess;
  if (P != 196 || side_only) return cudaErrorInvalidValue;
  {
    static int use_cs = -1;
    if (use_cs < 0) {
      const char* e = getenv("OAKE_ATTN");
      use_cs = (e != nullptr && e[0] == 'p') ? 0 : 1;
    }
    if (use_cs) return launch_attention_cs(st, qkv, mask, out, B, P, heads, with_side, side_only);
  }
  const int rows = B * (P + 1) + (with_side ? B : 0);
  if (with_side) {
    if (mask == nullptr) return cudaErrorInvalidValue;
    return launch_pp<true>(st, qkv, mask, out, B, heads, rows);
  }
  return launch_pp<false>(st, qkv, nullptr, out, B, heads, rows);
}

}  // namespace oake
