// Micro-probe (B200): tcgen05.ld throughput per SM as a function of the number of reading warps, and MUFU ex2
// throughput for f32 vs packed f16x2 operands.  Decides what bounds the 197-token softmax (attention_cs.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

#define LD32(taddr, r)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                            \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"          \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),    \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),           \
                 "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),         \
                 "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),         \
                 "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                              \
               : "r"(taddr)                                                                                         \
               : "memory")

// every warp w reads lane quarter (w % 4); `reps` rounds of 4 back-to-back x32 loads (128 columns) + one wait
__global__ void __launch_bounds__(512, 1) ldtm_kernel(long long* out, int warps, int reps, int wait_each) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t base = tmem_ptr + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128 % 512;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (warp < warps) {
    t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      uint32_t a[32], b[32], c[32], d[32];
      LD32(base, a);
      if (wait_each) asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      LD32(base + 32, b);
      if (wait_each) asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      LD32(base + 64, c);
      if (wait_each) asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      LD32(base + 96, d);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= a[j] ^ b[j] ^ c[j] ^ d[j];
    }
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x % 32 == 0 && warp < warps) {
    out[warp * 2] = t1 - t0;
    out[warp * 2 + 1] = acc;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_ptr) : "memory");
}

// MUFU: `reps` x 32 independent ex2 per thread
template <int MODE>
__global__ void __launch_bounds__(512, 1) mufu_kernel(long long* out, float* sink, int warps, int reps, float seed) {
  const int warp = threadIdx.x >> 5;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = seed * (j + 1) - threadIdx.x * 1e-3f;
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (warp < warps) {
    t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) asm volatile("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(v[j]) : "f"(v[j] - 1.0f));
      } else if (MODE == 1) {  // 16 packed ops = 32 elements
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint32_t p;
          asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;\n" : "=r"(p) : "f"(v[2 * j + 1] - 1.0f), "f"(v[2 * j] - 1.0f));
          asm volatile("ex2.approx.f16x2 %0, %1;\n" : "=r"(p) : "r"(p));
          v[2 * j] = __uint_as_float(p);
          v[2 * j + 1] = __uint_as_float(p ^ 0x1u) * 0.5f;
        }
      } else {  // polynomial exp2 on the FMA pipe (Cody-Waite split + degree-3), 32 elements
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = v[j] - 1.0f;
          const float fl = floorf(x);
          const float f = x - fl;
          float p = fmaf(f, 0.0555054f, 0.2402265f);
          p = fmaf(p, f, 0.6931472f);
          p = fmaf(p, f, 1.0f);
          v[j] = __int_as_float(__float_as_int(p) + (static_cast<int>(fl) << 23));
        }
      }
    }
    t1 = clock64();
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  if (warp < warps) sink[threadIdx.x] = s;
  if (threadIdx.x % 32 == 0 && warp < warps) out[warp * 2] = t1 - t0;
}

int main() {
  long long* out;
  float* sink;
  cudaMallocManaged(&out, 64 * sizeof(long long));
  cudaMalloc(&sink, 512 * sizeof(float));
  const int reps = 256;
  printf("== tcgen05.ld 32x32b.x32: cycles per 4 KB load (per warp), bytes/cycle per SM\n");
  for (int wait_each = 0; wait_each < 2; ++wait_each)
    for (int warps : {1, 2, 4, 8, 16}) {
      for (int k = 0; k < 2; ++k) {
        ldtm_kernel<<<1, 512>>>(out, warps, reps, wait_each);
        cudaDeviceSynchronize();
      }
      long long worst = 0;
      for (int w = 0; w < warps; ++w) worst = out[w * 2] > worst ? out[w * 2] : worst;
      const double loads = 4.0 * reps;
      printf("wait_each=%d warps=%2d: %7.1f cyc/load/warp  -> %6.1f B/cyc/SM   (%s)\n", wait_each, warps, worst / loads,
             warps * loads * 4096.0 / worst, cudaGetErrorString(cudaGetLastError()));
    }
  printf("== MUFU ex2: cycles per 32 elements per warp\n");
  const char* names[3] = {"ex2.f32", "cvt+ex2.f16x2", "poly3 (FMA pipe)"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      for (int k = 0; k < 2; ++k) {
        if (mode == 0) mufu_kernel<0><<<1, 512>>>(out, sink, warps, reps, 0.01f);
        if (mode == 1) mufu_kernel<1><<<1, 512>>>(out, sink, warps, reps, 0.01f);
        if (mode == 2) mufu_kernel<2><<<1, 512>>>(out, sink, warps, reps, 0.01f);
        cudaDeviceSynchronize();
      }
      long long worst = 0;
      for (int w = 0; w < warps; ++w) worst = out[w * 2] > worst ? out[w * 2] : worst;
      printf("%-18s warps=%2d (%d per SMSP): %6.1f cyc per 32 elements per warp -> %5.2f elements/cyc/SM (%s)\n", names[mode],
             warps, (warps + 3) / 4, static_cast<double>(worst) / reps, warps * 32.0 * 32.0 * reps / worst,
             cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
