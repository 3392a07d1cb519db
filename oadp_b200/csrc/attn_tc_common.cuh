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

}  // namespace attn
}  // namespace oake
