// Persistent tcgen05 attention for the 197-token OAKE tower (SURVEY 2.2 K4 + K7): the math of
// nn.MultiheadAttention inside every CLIP ResidualAttentionBlock (reference call site
// oadp/oake/objects.py:330) plus the objects side token of oadp/oake/objects.py:224-247 as one more
// query row / key row of the same tile.
//
// One CTA per SM walks over (crop, head) items; an item is two 128-row query tiles (tokens 0..127 and
// 128..197) against one 208-key K/V tile: S = Q K^T into TMEM, P = exp2(S - max) as packed fp16
// written over the consumed S columns, O = P V with P as the TMEM operand and V MN-major.  Division
// of labour (chosen from ncu stall samples of a first version in which each tile belonged to ONE
// group of four softmax warps that also drained it: per group the chain softmax -> PV -> O drain ->
// next S was serial and the MUFU pipe idle through its gaps, profiles/r1_11_attention_stalls.txt):
//
//   warps 0-7    softmax: ALL eight work on the current tile -- warp w owns TMEM lanes 32 (w % 4)..
//                and one half of the keys (w < 4: keys 0..95, w >= 4: keys 96..207); the two warps of
//                a lane quarter exchange their partial row maxima through shared memory (one named
//                barrier per pair) and leave partial row sums for the drain warps.  They alternate
//                between the two TMEM tiles and never wait for a PV product.
//   warps 8-11   drain: one per lane quarter; O -> registers (the tile's TMEM is released at once),
//                1 / sum, swizzled smem strip, whole 128-byte rows to global memory.
//   warp  12     loader (TMA boxes + cp.async rows, two-stage ring, non-blocking)
//   warp  13     MMA issue (one thread), order PV(n,0) S(n+1,0) PV(n,1) S(n+1,1)
//   warp  15     (objects) the side token's query row, see below
//
// so that PV(n,t), its drain and S(n+1,t) run underneath the softmax of the other tile.
//
// The objects side token (objects.py:224-247: one more query that sees the patches through the additive mask
// bias -100 * mask, itself, and not the class token) is row 197 of the item = row 69 of tile 1.  Its biased softmax
// used to be a special path of the two softmax warps that held it -- 1.5x their instructions for ONE live row, and a
// tile waits for its slowest warp: 171 us per launch against 135 us without the side row (B = 478,
// tools/bench_attn_side.py).  Now two SIDE warps (14 and 15: lane quarters 2 and 3, where that row lands on even /
// odd items) take it off the critical path: as soon as S of tile 1 is complete -- the softmax warps are still busy
// with tile 0 -- the side warp of the item's parity reads its quarter of S from TMEM, the one lane that holds the row
// spreads the 208 scores over shared memory, all 32 lanes do bias, maximum, exponentials and sum, and the packed
// probabilities wait in shared memory.  The eight softmax warps run ONE plain path for all rows; the thread that owns
// the side row swaps in the prepared probabilities (a few 16-byte loads) before the P store.  (A first attempt
// computed the row with mma.sync from the K / V tiles in a warp of its own: correct, but the legacy HMMA pipe
// issues about one m16n8k16 per 66 cycles and sub-partition here -- 291 us per launch.)  Any fp32 mask values are honoured.
//
// TMEM columns of a tile (256 per tile, 512 per CTA):
//   S   [0, 208)  fp32 scores, keys in order (196 patches, class, side, 10 x padding)
//   P   [0, 48)   packed fp16, keys 0..95    (lower half, written in place behind its reads)
//       [96, 152) packed fp16, keys 96..207  (upper half, likewise)
//   O   [152, 216) fp32 accumulator (dead S columns of the upper half)
#include <stdlib.h>

#include "attn_tc_common.cuh"

namespace oake {

namespace {

using namespace attn;

// -DOAKE_ATTN_TRACE: block 0 records clock64() at the hand-over points of the first 32 tiles (developer builds
// only: tools/gpu_attn_trace.sh); otherwise the macro vanishes.
#ifdef OAKE_ATTN_TRACE
__device__ long long g_attn_trace[8 * 32 * 4];
#define OAKE_TRACE(role, k, ev)                                                                        \
  do {                                                                                                 \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (role) < 7 && (k) < 32)                          \
      g_attn_trace[((role) * 32 + (k)) * 4 + ((ev) & 3)] = clock64();                                        \
  } while (0)
#else
#define OAKE_TRACE(role, k, ev) \
  do {                          \
  } while (0)
#endif

// Two measured-and-rejected variants of the softmax, kept behind build flags (A/B on one box, ms per 11 launches at
// B = 478: default 2.88-2.90; SUMMMA 2.87-3.03; MAXPIPE 3.07-3.14; both 3.19-3.25 -- tools/gpu_attn_ab.sh):
// -DOAKE_ATTN_SUMMMA=1: the row sum as one more product of P -- against a 16 x 16 tile of ones, into 16 spare TMEM
//   columns of the tile, issued behind P V; the drain warp reads it next to O.  Removes one FADD per key from the
//   softmax warps: no gain, i.e. the exponential phase is not bound by its instruction count.
// -DOAKE_ATTN_MAXPIPE=1: two score chunks in flight in the row-maximum pass and three-input maxima: slower (the 64
//   registers of scores in flight push a few values into local memory, which lives in L2 next to 195 KB of shared memory).
#ifndef OAKE_ATTN_SUMMMA
#define OAKE_ATTN_SUMMMA 0
#endif
#ifndef OAKE_ATTN_MAXPIPE
#define OAKE_ATTN_MAXPIPE 0
#endif
// -DOAKE_ATTN_TAILMERGE=0: the upper half loads its 16-column tail (keys 192..207) in a fourth round of its own.
// Default: the tail is requested together with the third 32-column chunk, so both halves of a lane quarter wait for
// three TMEM loads per pass -- a tile waits for its slowest warp, and the upper half was it.
#ifndef OAKE_ATTN_TAILMERGE
#define OAKE_ATTN_TAILMERGE 1
#endif
constexpr bool kTailMerge = OAKE_ATTN_TAILMERGE != 0;
// -DOAKE_ATTN_EXPPIPE=1: the exponential pass in 16-column pieces, the next piece's TMEM load in flight during the
// arithmetic of the current one (two 16-register buffers).  Same arithmetic per key.
#ifndef OAKE_ATTN_EXPPIPE
#define OAKE_ATTN_EXPPIPE 0
#endif
constexpr bool kExpPipe = OAKE_ATTN_EXPPIPE != 0;
constexpr bool kSumMma = OAKE_ATTN_SUMMMA != 0;
constexpr bool kMaxPipe = OAKE_ATTN_MAXPIPE != 0;

template <bool SIDE>
struct CCfg {
  static constexpr int P = 196;
  static constexpr int T = P + 1;                // keys every row may see: patches + class
  static constexpr int TQ = T + (SIDE ? 1 : 0);  // rows / keys of an item: + the side token (row 197, seen only by itself)
  static constexpr int NK = 208;                 // keys, padded to the MMA K granule (16)
  static constexpr int kUnits = NK / 16;         // 13
  static constexpr int kSplit = 96;              // first key of the upper half
  static constexpr int kRows1 = TQ - 128;        // live rows of tile 1: 68 patches + class (+ side)
  static constexpr int kPatch1 = P - 128;        // patch rows of tile 1 (one TMA box)
  static constexpr int kShift = 56;              // lane offset of tile 1's rows on odd items
  static constexpr int kQTile = 128 * 128;       // bytes: 128 rows x 64 halves
  static constexpr int kKV = NK * 128;           // bytes
  static constexpr int kStage = 2 * kQTile + 2 * kKV;
  static constexpr int kPHi = kSplit;            // P columns of the upper half start here
  static constexpr int kOCol = kSplit + (NK - kSplit) / 2;  // 152
  static constexpr int kSumCol = kOCol + 64;                // 216: row sums (P x ones), 16 columns
  static constexpr int kOnesBytes = 2048;                   // 16 rows x 128 B of fp16 ones: any layout reads ones
  // side warps' scratch, one set per item parity: 208 fp32 scores, 104 words of packed probabilities, 2 partial sums
  static constexpr int kSideX = 0, kSideP = 2 * 208 * 4, kSideSum = kSideP + 2 * 104 * 4;
  static constexpr int kSideBytes = SIDE ? kSideSum + 64 : 0;
  static constexpr int kBufCols = 256;           // TMEM columns per tile
  static constexpr int kMaskFloats = 208;        // staged mask row (196 used), 16-byte granules
  static constexpr int kMaskBytes = 2048;        // both stages' mask rows, padded
  static constexpr int kXchgBytes = 2 * 2 * 2 * 128 * 4;  // partial max + sum: [2][tile][half][row]
  static constexpr int kOutStage = 32 * 128;     // bytes per drain warp: 32 rows x 64 halves
  static constexpr int kNumBars = 18;
  static constexpr int kSmemBytes = 1024 + 2 * kStage + kMaskBytes + kXchgBytes + 4 * kOutStage + kOnesBytes + kSideBytes + kNumBars * 8 + 16;
  static constexpr int kSoftmaxWarps = 8;
  static constexpr int kDrainWarp0 = 8;
  static constexpr int kLoaderWarp = 12;
  static constexpr int kMmaWarp = 13;
  static constexpr int kSideWarp0 = 14;          // (objects) warps 14 / 15 = lane quarters 2 / 3 = side row of even / odd items
  static constexpr int kThreads = SIDE ? 32 * 16 : 32 * 14;
  // Register-resident form (RS): 16 warps = four warpgroups, the unit `setmaxnreg` works on.  At launch every
  // thread owns 128 registers (64 K / 512); the two softmax warpgroups then grow to kRegsSoftmax (a half row of
  // S, 96 / 112 scores, lives in registers from one TMEM read to the exponentials), paid for by the drain
  // warpgroup (kRegsDrain) and the loader / MMA / two idle warps (kRegsMisc): 2*176 + 104 + 56 = 512 = 4 * 128.
  static constexpr int kThreadsRS = 32 * 16;
  static constexpr int kRegsSoftmax = 176, kRegsDrain = 104, kRegsMisc = 56;
};

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N)); }
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;\n" : "=f"(d) : "f"(a), "f"(b), "f"(c));  // FMNMX3
  return d;
}

// Named barrier of the two softmax warps of lane quarter q (ids 1..4, 64 threads; id 0 is
// __syncthreads).  Immediate ids: ptxas then reserves five barriers instead of all sixteen.
// __noinline__: both halves' instantiations then arrive from the same instruction, which is what
// compute-sanitizer's synccheck expects of the participants of one barrier.
__device__ __noinline__ void pair_sync(int q) {
  switch (q) {
    case 0: asm volatile("bar.sync 1, 64;\n" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;\n" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;\n" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;\n" ::: "memory"); break;
  }
}

// RS kernels: the same barrier without the call (an ABI call with a half row of scores live in registers would
// move them around the callee-saved set); the id is a register operand.
__device__ __forceinline__ void pair_sync_inline(int q) { asm volatile("bar.sync %0, 64;\n" ::"r"(q + 1) : "memory"); }

// One half (HI = 0: keys 0..95, HI = 1: keys 96..207) of one query row of the tile.  Keys 0..196 (patches, class)
// are valid for every row; key 197 (the objects side token's own key) and the padding are not.
template <bool SIDE, int HI>
struct HalfRow {
  using C = CCfg<SIDE>;
  static constexpr float scale = 0.125f * kLog2e;
  static constexpr int k0 = HI ? C::kSplit : 0;   // first key
  static constexpr int kTail = 192;               // keys 192..207: patches 192..195, class, (side), padding

  uint32_t t_row;

  // partial row maximum in the base-2 domain (scaled)
  __device__ __forceinline__ float pass_max() const {
    if (kMaxPipe) return pass_max_pipelined();
    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    uint32_t rt[16];
    constexpr bool merge = HI && kTailMerge;
#pragma unroll 1
    for (int c = 0; c < (merge ? 2 : 3); ++c) {
      uint32_t ra[32];
      tmem_ld_32x32(t_row + k0 + c * 32, ra);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(ra[j]));
    }
    if (merge) {  // third chunk and tail in one round
      uint32_t ra[32];
      tmem_ld_32x32(t_row + k0 + 64, ra);
      tmem_ld_32x16(t_row + kTail, rt);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(ra[j]));
    }
    float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    if (HI) {
      if (!merge) {
        tmem_ld_32x16(t_row + kTail, rt);
        tmem_ld_wait();
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (kTail + j < C::T) mx = fmaxf(mx, __uint_as_float(rt[j]));
    }
    return mx * scale;  // the scale is positive: max(s) * scale == max(s * scale)
  }

  // chunks 0 and 1 requested together, chunk 2 (and the tail) requested while chunk 1 is reduced (-DOAKE_ATTN_MAXPIPE=1)
  __device__ __forceinline__ float pass_max_pipelined() const {
    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    auto reduce = [&](const uint32_t (&a)[32]) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) m4[(j >> 1) & 3] = fmax3(m4[(j >> 1) & 3], __uint_as_float(a[j]), __uint_as_float(a[j + 1]));
    };
    uint32_t ra[32], rb[32], rt[16];
    tmem_ld_32x32(t_row + k0, ra);
    tmem_ld_32x32(t_row + k0 + 32, rb);
    tmem_ld_wait();
    reduce(ra);
    tmem_ld_32x32(t_row + k0 + 64, ra);
    reduce(rb);
    if (HI) tmem_ld_32x16(t_row + kTail, rt);  // (after chunk 1 is consumed: its registers are free)
    tmem_ld_wait();
    reduce(ra);
    float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    if (HI) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (kTail + j < C::T) mx = fmaxf(mx, __uint_as_float(rt[j]));  // keys 192 .. 196: patches and the class key
    }
    return mx * scale;
  }

  // p = exp2(s - max) -> packed fp16 behind the reads; partial row sum of the unrounded values.
  // `y_words` != 0 (the one thread that owns the objects side row): the row's packed probabilities, prepared by a
  // side warp (shared-window address of word 0 = keys 0, 1), replace what the plain path computed for it.
  // 16-column pieces, double buffered (kExpPipe)
  __device__ __forceinline__ float pass_exp_pipelined(float neg_mx, uint32_t y_words) const {
    constexpr int pbase = HI ? C::kPHi : 0;
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    // `po`: pair offset of the piece inside its 32-key chunk (0 / 8): the FMA-pipe exponentials are chosen by key
    auto piece = [&](const uint32_t (&r)[16], int po, int col8) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool poly = po == 0 ? use_poly(j) : use_poly(8 + j);
        const float p0 = ex2_sel(fmaf(__uint_as_float(r[2 * j]), scale, neg_mx), poly);
        const float p1 = ex2_sel(fmaf(__uint_as_float(r[2 * j + 1]), scale, neg_mx), poly);
        if (!kSumMma) s4[j & 3] += p0 + p1;
        pk[j] = pack2(p0, p1);
      }
      if (SIDE && y_words != 0u) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint4 w = lds128(y_words + (k0 / 2 + col8 + 4 * j) * 4);
          pk[4 * j + 0] = w.x;
          pk[4 * j + 1] = w.y;
          pk[4 * j + 2] = w.z;
          pk[4 * j + 3] = w.w;
        }
      }
      tmem_st_32x8(t_row + pbase + col8, pk);
    };
    uint32_t a[16], b[16];
    tmem_ld_32x16(t_row + k0, a);
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      tmem_ld_wait();
      tmem_ld_32x16(t_row + k0 + c * 32 + 16, b);
      piece(a, 0, c * 16);
      tmem_ld_wait();
      if (HI || c < 2) tmem_ld_32x16(t_row + k0 + (c + 1) * 32, a);  // next chunk, or the upper half's tail (k0 + 96 = 192)
      piece(b, 8, c * 16 + 8);
    }
    if (HI) {
      tmem_ld_wait();
      uint32_t pt[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float p[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int jj = 2 * j + e;
          p[e] = kTail + jj < C::T ? ex2(fmaf(__uint_as_float(a[jj]), scale, neg_mx)) : 0.f;
        }
        if (!kSumMma) s4[j & 3] += p[0] + p[1];
        pt[j] = pack2(p[0], p[1]);
      }
      if (SIDE && y_words != 0u) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint4 w = lds128(y_words + (kTail / 2 + 4 * j) * 4);
          pt[4 * j + 0] = w.x;
          pt[4 * j + 1] = w.y;
          pt[4 * j + 2] = w.z;
          pt[4 * j + 3] = w.w;
        }
      }
      tmem_st_32x8(t_row + pbase + 48, pt);
    }
    return (s4[0] + s4[1]) + (s4[2] + s4[3]);
  }

  __device__ __forceinline__ float pass_exp(float neg_mx, uint32_t y_words) const {
    if (kExpPipe) return pass_exp_pipelined(neg_mx, y_words);
    constexpr int pbase = HI ? C::kPHi : 0;
    constexpr bool merge = HI && kTailMerge;
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    auto do_chunk = [&](int c, const uint32_t (&ra)[32]) {
      uint32_t pa[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p0 = ex2_sel(fmaf(__uint_as_float(ra[2 * j]), scale, neg_mx), use_poly(j));
        const float p1 = ex2_sel(fmaf(__uint_as_float(ra[2 * j + 1]), scale, neg_mx), use_poly(j));
        if (!kSumMma) s4[j & 3] += p0 + p1;
        pa[j] = pack2(p0, p1);
      }
      if (SIDE && y_words != 0u) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 w = lds128(y_words + (k0 / 2 + c * 16 + 4 * j) * 4);
          pa[4 * j + 0] = w.x;
          pa[4 * j + 1] = w.y;
          pa[4 * j + 2] = w.z;
          pa[4 * j + 3] = w.w;
        }
      }
      tmem_st_32x16(t_row + pbase + c * 16, pa);
    };
#pragma unroll 1
    for (int c = 0; c < (merge ? 2 : 3); ++c) {
      uint32_t ra[32];
      tmem_ld_32x32(t_row + k0 + c * 32, ra);
      tmem_ld_wait();
      do_chunk(c, ra);
    }
    uint32_t rt[16];
    if (merge) {  // third chunk and tail in one round
      uint32_t ra[32];
      tmem_ld_32x32(t_row + k0 + 64, ra);
      tmem_ld_32x16(t_row + kTail, rt);
      tmem_ld_wait();
      do_chunk(2, ra);
    }
    if (HI) {
      if (!merge) {
        tmem_ld_32x16(t_row + kTail, rt);
        tmem_ld_wait();
      }
      uint32_t pt[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float p[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int jj = 2 * j + e;
          p[e] = kTail + jj < C::T ? ex2(fmaf(__uint_as_float(rt[jj]), scale, neg_mx)) : 0.f;
        }
        if (!kSumMma) s4[j & 3] += p[0] + p[1];
        pt[j] = pack2(p[0], p[1]);
      }
      if (SIDE && y_words != 0u) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint4 w = lds128(y_words + (kTail / 2 + 4 * j) * 4);
          pt[4 * j + 0] = w.x;
          pt[4 * j + 1] = w.y;
          pt[4 * j + 2] = w.z;
          pt[4 * j + 3] = w.w;
        }
      }
      tmem_st_32x8(t_row + pbase + 48, pt);
    }
    return (s4[0] + s4[1]) + (s4[2] + s4[3]);
  }
};

// pass_max -> exchange with the other half's warp -> pass_exp; leaves the partial sum in `xsum`.
// warp_y: this warp's rows include the objects side row (is_y: this thread's row is it) -- its probabilities and its
// sum come from the side warp (`y_ready`, `y_words`, `y_sum`), everything else is the one plain path.
template <bool SIDE, int HI>
__device__ __forceinline__ void softmax_half(uint32_t t_row, int q, float* xmax_mine, const float* xmax_other,
                                             float* xsum_mine, bool warp_y, bool is_y, uint64_t* y_ready,
                                             uint32_t y_parity, uint32_t y_words, const float* y_sum) {
  HalfRow<SIDE, HI> h{t_row};
  const float mine = h.pass_max();
  *xmax_mine = mine;
  pair_sync(q);
  const float mx = fmaxf(mine, *xmax_other);
  if (SIDE && warp_y) mbar_wait(y_ready, y_parity);  // (long complete: the side warp ran under the softmax of tile 0)
  const float sum = h.pass_exp(-mx, (SIDE && is_y) ? y_words : 0u);
  if (!kSumMma) *xsum_mine = (SIDE && is_y) ? y_sum[HI] : sum;
}

// ------------------------------------------------------------------------------------------------
// Register-resident half row (RS kernels).  What bounds the two-pass form above is not a pipe but the
// tcgen05.ld itself: one 32x32b.x32 load occupies a warp for >= 110 cycles however many are queued
// (tools/probes/tmem_probe.cu: 37 B/clk for one warp, 66 for the two of a sub-partition), and a half row was
// swept twice -- seven or eight loads per tile with the MUFU pipe idle under each.  Here the half row is read
// ONCE: all loads are issued back to back into registers, the maximum, the exchange with the other half's warp
// and the exponentials then run from registers.  As the exponentials consume a 32-score chunk its registers are
// refilled with the SAME chunk of the warp's next tile (whose S has long been complete: the PV -> drain -> S
// chain of a tile is shorter than the softmax of the other one), so in steady state a tile starts with its
// scores already in registers and the MUFU pipe sees one gap per tile (maximum + exchange) instead of eight.
// Same arithmetic per element as HalfRow, so batch composition stays invisible; kLoad masks (non 0 / 1 values)
// keep the two-pass form.
// ------------------------------------------------------------------------------------------------
template <bool SIDE, int HI>
struct RegRow {
  using C = CCfg<SIDE>;
  static constexpr float scale = 0.125f * kLog2e;
  static constexpr float kNegB = -100.0f * kLog2e;
  static constexpr int k0 = HI ? C::kSplit : 0;
  static constexpr int kTail = 192;
  static constexpr int pbase = HI ? C::kPHi : 0;

  __device__ static __forceinline__ void issue_chunk(uint32_t t_row, int c, uint32_t (&a)[32]) {
    tmem_ld_32x32(t_row + k0 + c * 32, a);
  }
  __device__ static __forceinline__ void issue_tail(uint32_t t_row, uint32_t (&tail)[16]) {
    if (HI) tmem_ld_32x16(t_row + kTail, tail);
  }
};

// One tile of one half row from registers.  The scores arrive biased already (`bias_side_row`), so there is one
// form for every row.  `pre`: refill the consumed chunks from `t_next`.  Returns the partial row sum.
template <bool SIDE, int HI>
__device__ __forceinline__ float softmax_regs(uint32_t (&a0)[32], uint32_t (&a1)[32], uint32_t (&a2)[32],
                                              uint32_t (&tail)[16], uint32_t t_row, uint32_t t_next, uint64_t* next_full,
                                              uint32_t next_parity, uint32_t& have, int q, float* xmax_mine,
                                              const float* xmax_other, int trace_role = 7, int trace_k = 0) {
  using R = RegRow<SIDE, HI>;
  using C = CCfg<SIDE>;
  constexpr float scale = R::scale;
  constexpr int kTailLive = C::TQ - R::kTail;  // keys 192 .. 196 (, 197): what is not padding

  // ---- maximum of the raw scores (the scale is positive: max(s) * scale == max(s * scale))
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  auto max_chunk = [&](const uint32_t (&a)[32]) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) m4[(j >> 1) & 3] = fmax3(m4[(j >> 1) & 3], __uint_as_float(a[j]), __uint_as_float(a[j + 1]));
  };
  max_chunk(a0);
  max_chunk(a1);
  max_chunk(a2);
  float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  if (HI) {
#pragma unroll
    for (int j = 0; j < kTailLive; ++j) mx = fmaxf(mx, __uint_as_float(tail[j]));
  }
  mx *= scale;
  *xmax_mine = mx;
  pair_sync_inline(q);
  const float neg_mx = -fmaxf(mx, *xmax_other);
  OAKE_TRACE(trace_role, trace_k, 2);  // softmax: maximum exchanged

  // ---- exponentials -> packed fp16 P behind the reads.  A consumed chunk's registers are refilled with the same
  // chunk of the warp's next tile as soon as that tile's S is complete (`next_full`; nullptr = no next tile):
  // asked again before every refill, so a tile whose S lands in the middle of this phase still gets its later
  // chunks early.  `have` collects which chunks (bit c; bit 3 = tail) are on their way.
  have = 0u;
  bool ready = false;
  auto poll = [&]() {
    if (next_full != nullptr && !ready) {
      ready = __all_sync(0xffffffffu, mbar_test_wait(next_full, next_parity));
      if (ready) tc_fence_after();
    }
  };
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
  auto exp_chunk = [&](uint32_t (&a)[32], int c) {
    uint32_t pa[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // (all on the MUFU pipe here: with the polynomial mixed in, this fully unrolled phase ran 50 % SLOWER)
      const float p0 = ex2(fmaf(__uint_as_float(a[2 * j]), scale, neg_mx));
      const float p1 = ex2(fmaf(__uint_as_float(a[2 * j + 1]), scale, neg_mx));
      s4[j & 3] += p0 + p1;
      pa[j] = pack2(p0, p1);
    }
    tmem_st_32x16(t_row + R::pbase + c * 16, pa);
    poll();
    if (ready) {
      R::issue_chunk(t_next, c, a);
      have |= 1u << c;
    }
  };
  exp_chunk(a0, 0);
  exp_chunk(a1, 1);
  exp_chunk(a2, 2);
  if (HI) {
    uint32_t pt[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) p[e] = 2 * j + e < kTailLive ? ex2(fmaf(__uint_as_float(tail[2 * j + e]), scale, neg_mx)) : 0.f;
      s4[j & 3] += p[0] + p[1];
      pt[j] = pack2(p[0], p[1]);
    }
    tmem_st_32x8(t_row + R::pbase + 48, pt);
    poll();
    if (ready) {
      R::issue_tail(t_next, tail);
      have |= 8u;
    }
  }
  return (s4[0] + s4[1]) + (s4[2] + s4[3]);
}

// SIDE kernels, after the scores of a tile have landed in registers: what distinguishes the side row from the
// main-stream rows is added to the raw scores once (objects.py:204-247) -- for the side row -100 mask[key] on
// the patches (in score units: / 0.125) and -inf on the class key (it sees itself instead); for every other row
// -inf on the side token's key.  Main-stream rows add exactly 0 to their patch scores, so a row's bits do not
// depend on whether the side row shares its warp.  Any fp32 mask values are honoured (0 / 1 in the reference).
template <bool SIDE, int HI>
__device__ __forceinline__ void bias_side_row(uint32_t (&a0)[32], uint32_t (&a1)[32], uint32_t (&a2)[32],
                                              uint32_t (&tail)[16], bool warp_y, bool is_y, uint32_t ymask_addr) {
  using R = RegRow<SIDE, HI>;
  using C = CCfg<SIDE>;
  if (!SIDE) return;
  constexpr float kPerMask = -100.0f / 0.125f;
  if (warp_y) {  // warp-uniform: only the warp pair that holds the side row pays for the mask row
    auto add = [&](uint32_t (&a)[32], int c) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float m = lds_f32(ymask_addr + (R::k0 + c * 32 + j) * 4);  // same address in every lane: a broadcast
        a[j] = __float_as_uint(__uint_as_float(a[j]) + (is_y ? kPerMask * m : 0.f));
      }
    };
    add(a0, 0);
    add(a1, 1);
    add(a2, 2);
    if (HI) {
#pragma unroll
      for (int j = 0; j < C::P - R::kTail; ++j) {
        const float m = lds_f32(ymask_addr + (R::kTail + j) * 4);
        tail[j] = __float_as_uint(__uint_as_float(tail[j]) + (is_y ? kPerMask * m : 0.f));
      }
    }
  }
  if (HI) {
    constexpr int jc = C::P - R::kTail, js = C::T - R::kTail;  // class key, side key
    tail[jc] = is_y ? __float_as_uint(-INFINITY) : tail[jc];
    tail[js] = is_y ? tail[js] : __float_as_uint(-INFINITY);
  }
}

template <bool SIDE, bool RS>
__global__ void __launch_bounds__(RS ? CCfg<SIDE>::kThreadsRS : CCfg<SIDE>::kThreads, 1)
attention_cs_kernel(const __grid_constant__ CUtensorMap tmQ0,  // qkv [R, 3W], box {64, 128}
                    const __grid_constant__ CUtensorMap tmQ1,  // qkv [R, 3W], box {64, 68}
                    const __grid_constant__ CUtensorMap tmKV,  // qkv [R, 3W], box {64, 196}
                    const act_t* __restrict__ qkv, const float* __restrict__ mask, act_t* __restrict__ out,
                    int B, int heads) {
  using C = CCfg<SIDE>;
  constexpr int P = C::P;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* ymask = reinterpret_cast<float*>(smem + 2 * C::kStage);  // [2][kMaskFloats]: mask row of the item's crop
  float* xmax = reinterpret_cast<float*>(smem + 2 * C::kStage + C::kMaskBytes);  // [tile][half][128]
  float* xsum = xmax + 2 * 2 * 128;                                               // [tile][half][128]
  uint8_t* out_stage = smem + 2 * C::kStage + C::kMaskBytes + C::kXchgBytes;      // [4 drain warps][kOutStage]
  uint8_t* ones = out_stage + 4 * C::kOutStage;  // [kOnesBytes] fp16 1.0
  uint8_t* side = ones + C::kOnesBytes;  // (objects) side warps' scratch: [parity] scores, packed probabilities, sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + C::kOnesBytes + C::kSideBytes);
  uint64_t* qk_full = bars + 0;   // [stage]  loader -> MMA
  uint64_t* qk_free = bars + 2;   // [stage]  MMA (S of both tiles retired) -> loader
  uint64_t* v_full = bars + 4;    // [stage]  loader -> MMA, side-row warps (mask row)
  uint64_t* v_free = bars + 6;    // [stage]  MMA (PV of both tiles retired) -> loader
  uint64_t* s_full = bars + 8;    // [tile]   MMA -> softmax + drain warps
  uint64_t* p_ready = bars + 10;  // [tile]   8 softmax warps -> MMA, drain warps (row sums)
  uint64_t* o_full = bars + 12;   // [tile]   MMA -> drain warps
  uint64_t* o_free = bars + 14;   // [tile]   4 drain warps (O in registers) -> MMA
  uint64_t* y_ready = bars + 16;  // [parity] side warp (the side row's probabilities are in shared memory) -> its softmax warps
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
      mbar_init(&y_ready[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], C::kSoftmaxWarps);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  if (kSumMma) {
    for (int i = threadIdx.x; i < C::kOnesBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0x3C003C00u;
    fence_proxy_async();  // generic-proxy writes -> tensor core reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // RS: register hand-over between the warpgroups.  Each `setmaxnreg` sits at the top of the code its
  // warpgroup runs (ptxas budgets registers per region: after a join of regions it assumes the smallest).
  if (warp >= C::kLoaderWarp) {
   if (RS) setmaxnreg_dec<C::kRegsMisc>();
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
      if (SIDE) {  // the crop's mask row (196 floats = 49 granules) for the side-row warps
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
    // every operand below derives from warp-uniform values: all 32 lanes run this loop, the election of the
    // issuing lane happens inside the *_warp wrappers
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    auto issue_s = [&](int n, int t) {
      const uint32_t stg = smem_base + (n & 1) * C::kStage;
      const uint32_t q_addr = stg + t * C::kQTile, k_addr = stg + 2 * C::kQTile;
#pragma unroll
      for (int k = 0; k < kDh / 16; ++k)
        umma_f16_ss_warp(tmem_u + t * C::kBufCols, make_smem_desc_k_sw128(q_addr + k * 32),
                         make_smem_desc_k_sw128(k_addr + k * 32), idesc_s, k != 0 ? 1u : 0u);
      umma_commit_warp(&s_full[t]);
    };
    constexpr uint32_t idesc_sum = make_idesc_f16(128, 16);
    const uint64_t ones_desc = make_smem_desc_k_sw128(smem_u32(ones));
    auto issue_pv = [&](int n, int t) {
      const uint32_t v_addr = smem_base + (n & 1) * C::kStage + 2 * C::kQTile + C::kKV;
      const uint32_t buf = tmem_u + t * C::kBufCols;
#pragma unroll
      for (int k = 0; k < C::kUnits; ++k) {
        const uint32_t p_col = k * 16 < C::kSplit ? k * 8 : C::kPHi + (k - C::kSplit / 16) * 8;
        umma_f16_ts_warp(buf + C::kOCol, buf + p_col, make_smem_desc_mn_sw128(v_addr + k * 2048), idesc_o, k != 0 ? 1u : 0u);
      }
      if (kSumMma) {  // row sums = P x ones (every k-step reads the same tile of ones)
#pragma unroll
        for (int k = 0; k < C::kUnits; ++k) {
          const uint32_t p_col = k * 16 < C::kSplit ? k * 8 : C::kPHi + (k - C::kSplit / 16) * 8;
          umma_f16_ts_warp(buf + C::kSumCol, buf + p_col, ones_desc, idesc_sum, k != 0 ? 1u : 0u);
        }
      }
      umma_commit_warp(&o_full[t]);
    };
    if (N > 0) {
      mbar_wait(&qk_full[0], 0);
      fence_proxy_async();  // the loader's cp.async rows (generic proxy) -> tensor core reads
      tc_fence_after();
      issue_s(0, 0);
      issue_s(0, 1);
      umma_commit_warp(&qk_free[0]);
    }
    for (int n = 0; n < N; ++n) {
      const int s = n & 1, u = n >> 1;
      mbar_wait(&v_full[s], u & 1);
      fence_proxy_async();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_ready[t], n & 1);
        OAKE_TRACE(1, 2 * n + t, 0);  // mma: p_ready seen
        tc_fence_after();
        issue_pv(n, t);
        OAKE_TRACE(1, 2 * n + t, 1);  // mma: PV issued
        if (t == 1) umma_commit_warp(&v_free[s]);
        if (n + 1 < N) {
          if (t == 0) {
            mbar_wait(&qk_full[s ^ 1], ((n + 1) >> 1) & 1);
            fence_proxy_async();
          }
          mbar_wait(&o_free[t], n & 1);
          OAKE_TRACE(1, 2 * n + t, 2);  // mma: o_free seen
          tc_fence_after();
          issue_s(n + 1, t);
          OAKE_TRACE(1, 2 * n + t, 3);  // mma: S(n+1, t) issued
          if (t == 1) umma_commit_warp(&qk_free[s ^ 1]);
        }
      }
    }
   } else if (SIDE && !RS && warp >= C::kSideWarp0) {
    // ================================================================== the objects side row's softmax
    // Row 197 = row 69 of tile 1: TMEM lane 69 on even items (quarter 2: warp 14), lane 56 + 69 = 125 on odd items
    // (quarter 3: warp 15).  Every side warp follows s_full[1] item by item (it can never fall two phases behind: the
    // next S of tile 1 needs this item's P, which needs its side row) and works on the items of its parity.
    constexpr float scale = 0.125f * kLog2e;
    constexpr float kNegB = -100.0f * kLog2e;
    const int par = warp - C::kSideWarp0;            // item parity served = scratch set
    const int q = warp & 3;                           // == 2 + par
    const int ylane = (par ? C::kShift : 0) + C::kRows1 - 1 - q * 32;  // 5 / 29
    float* sx = reinterpret_cast<float*>(side + C::kSideX) + par * C::NK;
    uint32_t* sp = reinterpret_cast<uint32_t*>(side + C::kSideP) + par * (C::NK / 2);
    float* ssum = reinterpret_cast<float*>(side + C::kSideSum) + par * 2;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 1 * C::kBufCols;
    for (int n = 0; n < N; ++n) {
      const int s = n & 1, u = n >> 1;
      mbar_wait(&s_full[1], n & 1);
      if (s != par) continue;
      tc_fence_after();
      // the row's 208 scores: TMEM -> the registers of lane `ylane` -> shared memory
#pragma unroll 1
      for (int c = 0; c < C::NK / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
        if (lane == ylane) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(sx + c * 32 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
      }
      {
        uint32_t r[16];
        tmem_ld_32x16(t_row + 192, r);
        tmem_ld_wait();
        if (lane == ylane) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(sx + 192 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
      }
      tc_fence_before();
      mbar_wait(&v_full[s], u & 1);  // the crop's mask row lands with the V stage
      __syncwarp();
      // bias (-100 mask on the patches, the class key and the padding out, its own key 197 in), maximum, exponentials
      const float* ym = ymask + s * C::kMaskFloats;
      float x[7];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const int j = lane + 32 * i;
        float v = -INFINITY;
        if (j < P) v = fmaf(sx[j], scale, kNegB * ym[j]);
        else if (j == C::T) v = sx[j] * scale;
        x[i] = v;
        mx = fmaxf(mx, v);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        x[i] = ex2(x[i] - mx);
        sum += x[i];
      }
      sum = warp_sum(sum);
      // packed pairs (keys 2 w, 2 w + 1): even lanes pack their value with their right neighbour's
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const float nb = __shfl_down_sync(0xffffffffu, x[i], 1);
        const int j = lane + 32 * i;
        if ((lane & 1) == 0 && j < C::NK) sp[j >> 1] = pack2(x[i], nb);
      }
      if (lane == 0) {
        ssum[0] = sum;
        ssum[1] = 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&y_ready[par]);  // (release: the scratch is visible to the softmax warps that wait)
    }
   }  // (warp 14, and 15 without a side token, only complete the fourth warpgroup)
  } else if (RS && warp < C::kDrainWarp0) {
    // ================================================================== softmax warps, register-resident form
    setmaxnreg_inc<C::kRegsSoftmax>();
    const int q = warp & 3;
    const int hi = warp >> 2;
    const int r = q * 32 + lane;
    uint32_t a0[32], a1[32], a2[32], tail[16];
    uint32_t have = 0u;  // chunks of the current tile that are already in (or on their way to) the registers
    constexpr uint32_t kAll = 15u;
    auto tile_rows = [&](int n_, int t_, bool& live_, bool& is_y_) {
      const int shift1 = (t_ == 1 && (n_ & 1)) ? C::kShift : 0;
      const int rr = r - shift1;
      live_ = t_ == 0 || (rr >= 0 && rr < C::kRows1);
      is_y_ = SIDE && live_ && t_ * 128 + rr == C::T;
    };
    bool live, is_y;
    tile_rows(0, 0, live, is_y);
    for (int n = 0; n < N; ++n) {
      const int s = n & 1, u = n >> 1;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const bool warp_live = __any_sync(0xffffffffu, live);
        const bool warp_y = SIDE && __any_sync(0xffffffffu, is_y);
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * C::kBufCols;
        float* xm = xmax + (t * 2) * 128 + r;
        float* xs = xsum + (t * 2) * 128 + r;
        OAKE_TRACE(warp == 0 ? 5 : (warp == 6 ? 6 : 7), 2 * n + t, have);  // (debug: slot = chunks prefetched)
        if (have == 0u) {  // (a tile with refilled chunks has already seen this phase of s_full complete)
          mbar_wait(&s_full[t], n & 1);
          tc_fence_after();
        }
        OAKE_TRACE(warp == 0 ? 0 : (warp == 6 ? 4 : 7), 2 * n + t, 0);  // softmax: s_full seen (or prefetched)
        // the warp's next tile
        const int n2 = t == 0 ? n : n + 1, t2 = t ^ 1;
        bool live2 = false, is_y2 = false;
        if (n2 < N) tile_rows(n2, t2, live2, is_y2);
        if (warp_live) {
          uint32_t ymask_addr = 0u;
          if (warp_y) {
            mbar_wait(&v_full[s], u & 1);  // the crop's mask row lands with the V stage
            ymask_addr = smem_u32(ymask + s * C::kMaskFloats);
          }
          if (hi == 0) {
            if (!(have & 1u)) RegRow<SIDE, 0>::issue_chunk(t_row, 0, a0);
            if (!(have & 2u)) RegRow<SIDE, 0>::issue_chunk(t_row, 1, a1);
            if (!(have & 4u)) RegRow<SIDE, 0>::issue_chunk(t_row, 2, a2);
          } else {
            if (!(have & 1u)) RegRow<SIDE, 1>::issue_chunk(t_row, 0, a0);
            if (!(have & 2u)) RegRow<SIDE, 1>::issue_chunk(t_row, 1, a1);
            if (!(have & 4u)) RegRow<SIDE, 1>::issue_chunk(t_row, 2, a2);
            if (!(have & 8u)) RegRow<SIDE, 1>::issue_tail(t_row, tail);
          }
          // refill from the next tile only if this warp has rows there
          uint64_t* next_full = (n2 < N && __any_sync(0xffffffffu, live2)) ? &s_full[t2] : nullptr;
          const uint32_t t_next = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t2 * C::kBufCols;
          tmem_ld_wait();
          OAKE_TRACE(warp == 0 ? 0 : (warp == 6 ? 4 : 7), 2 * n + t, 1);  // softmax: scores in registers
          if (hi == 0) {
            bias_side_row<SIDE, 0>(a0, a1, a2, tail, warp_y, is_y, ymask_addr);
            *xs = softmax_regs<SIDE, 0>(a0, a1, a2, tail, t_row, t_next, next_full, n2 & 1, have, q, xm, xm + 128,
                                        warp == 0 ? 0 : 7, 2 * n + t);
          } else {
            bias_side_row<SIDE, 1>(a0, a1, a2, tail, warp_y, is_y, ymask_addr);
            xs[128] = softmax_regs<SIDE, 1>(a0, a1, a2, tail, t_row, t_next, next_full, n2 & 1, have, q, xm + 128, xm,
                                            warp == 6 ? 4 : 7, 2 * n + t);
          }
          tmem_st_wait();
        } else {
          have = 0u;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
        OAKE_TRACE(warp == 0 ? 0 : (warp == 6 ? 4 : 7), 2 * n + t, 3);  // softmax: p_ready arrived
        live = live2;
        is_y = is_y2;
      }
    }
    (void)kAll;
  } else {
    // ================================================================== softmax and drain warps (RS: drain only)
    if (RS) setmaxnreg_dec<C::kRegsDrain>();
    const bool drain = RS ? true : warp >= C::kDrainWarp0;
    const int q = warp & 3;                    // TMEM lane quarter (== warp % 4 for both roles)
    const int hi = drain ? 0 : (warp >> 2);    // softmax: which half of the keys
    const int r = q * 32 + lane;               // lane of the tile
    for (int n = 0; n < N; ++n) {
      const int s = n & 1, u = n >> 1;
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / heads, h = item - b * heads;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int shift1 = (t == 1 && s) ? C::kShift : 0;
        const int rr = r - shift1;
        const bool live = t == 0 || (rr >= 0 && rr < C::kRows1);
        const int i = t * 128 + rr;  // token index
        const bool is_y = SIDE && live && i == C::T;
        const bool warp_live = __any_sync(0xffffffffu, live);
        const bool warp_y = SIDE && __any_sync(0xffffffffu, is_y);
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * C::kBufCols;
        float* xm = xmax + (t * 2) * 128 + r;  // + 128 for the upper half
        float* xs = xsum + (t * 2) * 128 + r;

        // every warp passes through s_full: it keeps warps without live rows in step with the tile
        mbar_wait(&s_full[t], n & 1);
        tc_fence_after();

        if (!drain) {
          if (warp_live) {
            // (objects) the side row's probabilities wait in the scratch set of the item's parity
            uint64_t* yr = &y_ready[s];
            const uint32_t yw = smem_u32(side + C::kSideP) + s * (C::NK / 2) * 4;
            const float* ys = reinterpret_cast<const float*>(side + C::kSideSum) + s * 2;
            if (hi == 0) softmax_half<SIDE, 0>(t_row, q, xm, xm + 128, xs, warp_y, is_y, yr, u & 1, yw, ys);
            else softmax_half<SIDE, 1>(t_row, q, xm + 128, xm, xs + 128, warp_y, is_y, yr, u & 1, yw, ys);
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_ready[t]);
        } else {
          if (warp_live) {
            float sum = 0.f;
            if (!kSumMma) {
              mbar_wait(&p_ready[t], n & 1);  // all eight softmax warps done: the row sums are in place
              sum = xs[0] + xs[128];
            }
            mbar_wait(&o_full[t], n & 1);
            OAKE_TRACE(warp == 8 ? 2 : 7, 2 * n + t, 0);  // drain: o_full seen
            tc_fence_after();
            uint32_t o0[32], o1[32], osum = 0u;
            tmem_ld_32x32(t_row + C::kOCol, o0);
            tmem_ld_32x32(t_row + C::kOCol + 32, o1);
            if (kSumMma) tmem_ld_32x1(t_row + C::kSumCol, osum);
            tmem_ld_wait();
            if (kSumMma) sum = __uint_as_float(osum);
            // the accumulator is in registers: the tile's TMEM goes back to the tensor core before
            // the scaling and the global stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[t]);
            OAKE_TRACE(warp == 8 ? 2 : 7, 2 * n + t, 1);  // drain: O in registers, o_free arrived
            const uint32_t stg = smem_u32(out_stage + (warp - C::kDrainWarp0) * C::kOutStage);
            {
              const float inv = 1.0f / sum;
              auto o_at = [&](int c) -> float { return c < 32 ? __uint_as_float(o0[c & 31]) : __uint_as_float(o1[c & 31]); };
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
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// OAKE_ATTN=rs selects the register-resident kernel (RS).  Measured on B200 it is exactly as fast as the two-pass
// form: both are bound by the exponentials of the slower (upper-half) softmax warp of a lane quarter, not by the
// tcgen05.ld sweeps RS removes (profiles/r2_05_attention_analysis.txt).  The two-pass kernel stays the default.
bool attention_rs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OAKE_ATTN");
    v = (e != nullptr && e[0] == 'r') ? 1 : 0;
  }
  return v == 1;
}

template <bool SIDE, bool RS>
cudaError_t launch_cs(cudaStream_t st, const act_t* qkv, const float* mask, act_t* out, int B, int heads, int rows) {
  using C = CCfg<SIDE>;
  if (cudaError_t e = ensure_dynamic_smem<attention_cs_kernel<SIDE, RS>>(C::kSmemBytes); e != cudaSuccess) return e;
  int dev = 0, num_sms = 0;
  if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return e;
  if (cudaError_t e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); e != cudaSuccess) return e;
  CUtensorMap tmQ0, tmQ1, tmKV;
  const uint64_t cols = 3ull * heads * kDh;
  if (make_tmap_act_2d(&tmQ0, qkv, rows, cols, 128) || make_tmap_act_2d(&tmQ1, qkv, rows, cols, C::kPatch1) ||
      make_tmap_act_2d(&tmKV, qkv, rows, cols, C::P))
    return cudaErrorInvalidValue;
  const int items = B * heads;
  const int grid = items < num_sms ? items : num_sms;
  attention_cs_kernel<SIDE, RS><<<grid, RS ? C::kThreadsRS : C::kThreads, C::kSmemBytes, st>>>(tmQ0, tmQ1, tmKV, qkv, mask, out,
                                                                                              B, heads);
  return cudaGetLastError();
}

}  // namespace

// The persistent tcgen05 kernel serves the 197-token tower whenever whole tiles are wanted; the
// 50-token tower stays on the mma.sync kernel of attention.cu and the last objects block (side row
// only) on its one-warp-per-head kernel.  OAKE_ATTN=mma forces the mma.sync kernels everywhere (A/B).
bool attention_use_tc(int P, int side_only) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OAKE_ATTN");
    v = (e != nullptr && e[0] == 'm') ? 0 : 1;
  }
  return v == 1 && P == 196 && !side_only;
}

#ifdef OAKE_ATTN_TRACE
extern "C" int oake_debug_attn_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_attn_trace, sizeof(g_attn_trace)) == cudaSuccess ? 0 : 1;
}
#endif

cudaError_t launch_attention_cs(cudaStream_t st, const act_t* qkv, const float* mask, act_t* out, int B, int P,
                                int heads, int with_side, int side_only) {
  if (B <= 0) return cudaSuccess;
  if (P != 196 || side_only) return cudaErrorInvalidValue;
  const int rows = B * (P + 1) + (with_side ? B : 0);
  if (with_side) {
    if (mask == nullptr) return cudaErrorInvalidValue;
    return launch_cs<true, false>(st, qkv, mask, out, B, heads, rows);  // (the opt-in RS form has no side warp)
  }
  return attention_rs() ? launch_cs<false, true>(st, qkv, nullptr, out, B, heads, rows)
                        : launch_cs<false, false>(st, qkv, nullptr, out, B, heads, rows);
}

}  // namespace oake
