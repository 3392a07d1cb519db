// Helpers of the tcgen05 attention kernel (attention_cs.cu): inline-PTX
// wrappers for TMEM stores / TS-MMA / MN-major descriptors, shared-window loads and stores, the
// token -> activation-row map.  Included inside each file's anonymous namespace users via `oake::attn`.
#pragma once

#include "kernels.cuh"

namespace oake {
namespace attn {

constexpr int kDh = 64;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= 0 on the FMA pipe: round x to the nearest integer n with the 1.5 * 2^23 trick, a degree-3 minimax
// polynomial for 2^f on f = x - n in [-0.5, 0.5] (max relative error 7.5e-5, a third of the half-ulp of the fp16 the
// result is rounded to), and n added into the exponent field.  The MUFU pipe delivers 16 ex2 per clock and SM and
// is what bounds the softmax (tools/probes/tmem_probe.cu); every kPolyEvery-th PAIR of keys (of a 32-key chunk) takes this path
// instead, chosen by key index only, so a row's arithmetic does not depend on where the row sits.
#ifndef OAKE_ATTN_POLY
#define OAKE_ATTN_POLY 4  // (every third pair measured best before the side warps, every fourth after: 139.7 vs 141.4 us)
#endif
constexpr int kPolyEvery = OAKE_ATTN_POLY;  // 0: every exponential on the MUFU pipe
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -120.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.055171605f, f, 0.24261111f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992806f);
  // n into the exponent field: p_bits + (t_bits << 23) as one integer multiply-add (the low 23 bits of t hold n)
  return __int_as_float(__float_as_int(t) * 0x800000 + __float_as_int(p));
}
// pair index j of a 32-key chunk -> which pipe
__device__ __forceinline__ constexpr bool use_poly(int j) { return kPolyEvery > 0 && (j % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1; }
__device__ __forceinline__ float ex2_sel(float x, bool poly) { return poly ? ex2_poly(x) : ex2(x); }

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  // volatile: the two softmax passes must each re-read the mask row instead of keeping 196 biases
  // alive (and spilled) in between
  asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr));
  return v;
}
// Arrives on `bar` (without raising its pending count) once every cp.async this thread has issued
// so far has landed: the loader never blocks on its own copies.
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr) : "memory");
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

// Warp-uniform issue forms.  A tcgen05.mma / commit is one instruction of the WARP's uniform datapath; written as
// `if (lane == 0) tcgen05.mma` ptxas cannot prove that a single lane is active and wraps every MMA in an
// elect / R2UR.BROADCAST / BRA.U.ANY loop over the active lanes (12 dependent instructions, ~95 cycles per MMA:
// the 13 k-steps of P V took 1100-2000 cycles to ISSUE, profiles/r2_05_attention_trace.txt).  Called by all 32
// lanes with warp-uniform operands and the election inside the asm, the same MMA is 4 uniform instructions.
__device__ __forceinline__ void umma_f16_ss_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts_warp(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_warp(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar))
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

}  // namespace attn
}  // namespace oake
