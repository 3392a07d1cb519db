// GPU JPEG decode in front of the OAKE resize: compressed files in, uint8 HWC RGB in
// the same arena `oake_resize_u8` reads -- the step oadp/oake/base.py:53 does with Pillow on the host.
// Integer / byte work, HBM- and latency-bound; results are bit-identical to Pillow's (the arithmetic
// is in jpeg_core.cuh).
//
//   entropy kernel   one CTA per image, one thread per 1024-bit subsequence of the scan: the threads
//                    find their entry states by iteration (Huffman streams re-synchronise), a prefix
//                    sum of the blocks per subsequence places them, a last pass writes the coefficients,
//                    a second prefix sum turns DC differences into DC values (jpeg_core.cuh, "parallel
//                    entropy decode").  ~10^3 threads per COCO-sized file instead of one.
//   restart kernel   files WITH restart markers: the host has located every restart interval while it
//                    copied the scan, each interval is an independent entry point -> one thread per
//                    interval, one CTA per image.
//   idct kernel      one thread per 8x8 block: dequantise, islow IDCT, range limit -> uint8 planes.
//   colour kernel    one thread per four pixels: fancy chroma upsampling + YCbCr -> RGB, HWC stores.
#include <cuda_runtime.h>
#include <stddef.h>

#include <string>

#include "../../include/oake_b200.h"
#include "jpeg_core.cuh"
#include "jpeg_parse.h"

namespace oake {
int fail_msg(const char* fmt, ...);  // encoder.cu
}

using namespace oake;

namespace {

constexpr int kRstThreads = 128;

// Files WITH restart markers: every restart interval is an entry point the host has already located
// (byte aligned, DC predictions reset), so one thread decodes one interval, serially, with absolute DC
// values -- no synchronisation rounds and no prefix sums.  One CTA per image.
__global__ void __launch_bounds__(kRstThreads) jpeg_entropy_rst_kernel(const uint8_t* __restrict__ bytes,
                                                                      const oake_jpeg_desc* __restrict__ descs,
                                                                      uint8_t* __restrict__ scratch,
                                                                      int32_t* __restrict__ status) {
  __shared__ oake_jpeg_huff tables[4];  // dc0 dc1 ac0 ac1
  // geometry part of the descriptor (everything in front of the quantisation tables): the only part the
  // entropy decode touches, kept next to the tables so that the serial loops never wait on global memory
  // for it
  constexpr int kHeadBytes = offsetof(oake_jpeg_desc, quant);
  static_assert(kHeadBytes % 8 == 0, "descriptor layout");
  __shared__ __align__(16) uint8_t head[kHeadBytes];
  const int tid = threadIdx.x;
  if (descs[blockIdx.x].restart_interval == 0) return;  // the subsequence-parallel kernel's share
  {
    const oake_jpeg_desc& g = descs[blockIdx.x];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&g.dc[0]);  // dc[2] and ac[2] are contiguous
    uint32_t* dst = reinterpret_cast<uint32_t*>(tables);
    for (int i = tid; i < static_cast<int>(sizeof(tables) / 4); i += kRstThreads) dst[i] = src[i];
    const uint32_t* hs = reinterpret_cast<const uint32_t*>(&g);
    for (int i = tid; i < kHeadBytes / 4; i += kRstThreads) reinterpret_cast<uint32_t*>(head)[i] = hs[i];
  }
  __syncthreads();
  const oake_jpeg_desc& d = *reinterpret_cast<const oake_jpeg_desc*>(head);
  for (uint32_t c = 0; c < d.ncomp; ++c) {
    const oake_jpeg_comp& k = d.comp[c];
    uint4* p = reinterpret_cast<uint4*>(scratch + k.coef_off);
    const uint32_t n16 = k.blocks_w * k.blocks_h * 8;  // 128 bytes per block
    for (uint32_t i = tid; i < n16; i += kRstThreads) p[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  int bad = 0;
  for (uint32_t k = tid; k < d.restart_count; k += kRstThreads) bad |= jpeg::decode_interval(d, bytes + d.scan_off, tables, scratch, k);
  bad = __syncthreads_or(bad);
  if (tid == 0) status[blockIdx.x] = bad ? 1 : 0;
}

constexpr int kParThreads = 256;

// exclusive prefix sum of one int per thread over the CTA; *total = sum of all.  `warp_sums`: shared, >= 8 ints.
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // warp_sums may still be read from a previous call
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  int before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kParThreads / 32; ++w) {
    const int t = warp_sums[w];
    if (w < warp) before += t;
    all += t;
  }
  *total = all;
  return before + inc - v;
}

__global__ void __launch_bounds__(kParThreads) jpeg_entropy_par_kernel(const uint8_t* __restrict__ bytes,
                                                                      const oake_jpeg_desc* __restrict__ descs,
                                                                      uint8_t* __restrict__ scratch,
                                                                      int32_t* __restrict__ status) {
  __shared__ oake_jpeg_huff tables[4];  // dc0 dc1 ac0 ac1
  constexpr int kHeadBytes = offsetof(oake_jpeg_desc, quant);
  __shared__ __align__(16) uint8_t head[kHeadBytes];
  __shared__ int warp_sums[kParThreads / 32];
  const int tid = threadIdx.x;
  if (descs[blockIdx.x].restart_interval != 0) return;  // the restart-interval kernel's share
  {
    const oake_jpeg_desc& g = descs[blockIdx.x];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&g.dc[0]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(tables);
    for (int i = tid; i < static_cast<int>(sizeof(tables) / 4); i += kParThreads) dst[i] = src[i];
    const uint32_t* hs = reinterpret_cast<const uint32_t*>(&g);
    for (int i = tid; i < kHeadBytes / 4; i += kParThreads) reinterpret_cast<uint32_t*>(head)[i] = hs[i];
  }
  __syncthreads();
  const oake_jpeg_desc& d = *reinterpret_cast<const oake_jpeg_desc*>(head);
  for (uint32_t c = 0; c < d.ncomp; ++c) {
    const oake_jpeg_comp& k = d.comp[c];
    uint4* p = reinterpret_cast<uint4*>(scratch + k.coef_off);
    const uint32_t n16 = k.blocks_w * k.blocks_h * 8;
    for (uint32_t i = tid; i < n16; i += kParThreads) p[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  const jpeg::McuMap map = jpeg::make_mcu_map(d);
  const uint8_t* stream = bytes + d.scan_off;
  const uint32_t bits = static_cast<uint32_t>(d.scan_len * 8);
  const uint32_t n = (bits + jpeg::kSubBits - 1) / jpeg::kSubBits;
  if (n == 0 || n > d.sync_slots) {  // (uniform) an empty scan; the slot count cannot be exceeded by a staged file
    if (tid == 0) status[blockIdx.x] = 1;
    return;
  }
  // subsequence table: entry state, exit state, blocks completed, first block
  volatile uint64_t* entry = reinterpret_cast<volatile uint64_t*>(scratch + d.sync_off);
  volatile uint64_t* exit_ = entry + d.sync_slots;
  volatile uint32_t* count = reinterpret_cast<volatile uint32_t*>(exit_ + d.sync_slots);
  volatile uint32_t* first = count + d.sync_slots;

  // 1. every subsequence from a guessed entry state (the first one from the true start)
  for (uint32_t i = tid; i < n; i += kParThreads) {
    const uint64_t e = jpeg::pack_state(i * jpeg::kSubBits, 0, 0);
    const jpeg::SubResult r = jpeg::decode_subsequence<false>(d, map, stream, tables, e, (i + 1) * jpeg::kSubBits, 0, nullptr);
    entry[i] = e;
    exit_[i] = r.exit;
    count[i] = r.count;
  }
  __syncthreads();
  // 2. chain the states until nothing changes: entry i is final after at most i rounds.  A thread may
  // see the exit state of its predecessor from this round or from the one before (single 8-byte
  // accesses); both are sound, and a round without any change proves every entry equals the exit next
  // to it.
  while (true) {
    int changed = 0;
    for (uint32_t i = tid; i < n; i += kParThreads) {
      if (i == 0) continue;
      const uint64_t e = exit_[i - 1];
      if (e == entry[i]) continue;
      const jpeg::SubResult r = jpeg::decode_subsequence<false>(d, map, stream, tables, e, (i + 1) * jpeg::kSubBits, 0, nullptr);
      entry[i] = e;
      exit_[i] = r.exit;
      count[i] = r.count;
      changed = 1;
    }
    if (!__syncthreads_or(changed)) break;
  }
  // 3. first block of every subsequence: exclusive prefix sum of the block counts
  int decoded_blocks;
  {
    const uint32_t per = (n + kParThreads - 1) / kParThreads;
    const uint32_t lo = min(tid * per, n), hi = min(lo + per, n);
    int sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += static_cast<int>(count[i]);
    int run = block_exclusive_scan(sum, warp_sums, &decoded_blocks);
    for (uint32_t i = lo; i < hi; ++i) {
      first[i] = static_cast<uint32_t>(run);
      run += static_cast<int>(count[i]);
    }
  }
  __syncthreads();
  // 4. the coefficients
  int bad = 0;
  for (uint32_t i = tid; i < n; i += kParThreads) {
    const jpeg::SubResult r = jpeg::decode_subsequence<true>(d, map, stream, tables, entry[i], (i + 1) * jpeg::kSubBits, first[i], scratch);
    bad |= r.bad ? 1 : 0;
  }
  bad = __syncthreads_or(bad);
  // 5. DC differences -> DC values: prefix sum per component over its blocks in scan order
  for (uint32_t c = 0; c < d.ncomp; ++c) {
    const oake_jpeg_comp& k = d.comp[c];
    const uint32_t hv = k.h * k.v, nb = d.mcus_x * d.mcus_y * hv;
    int16_t* plane = reinterpret_cast<int16_t*>(scratch + k.coef_off);
    auto block_of = [&](uint32_t q) -> int16_t* {
      const uint32_t mcu = q / hv, r = q - mcu * hv;
      const uint32_t v = r / k.h, h = r - v * k.h;
      const uint32_t my = mcu / d.mcus_x, mx = mcu - my * d.mcus_x;
      return plane + (static_cast<uint64_t>(my * k.v + v) * k.blocks_w + (mx * k.h + h)) * 64;
    };
    const uint32_t per = (nb + kParThreads - 1) / kParThreads;
    const uint32_t lo = min(tid * per, nb), hi = min(lo + per, nb);
    int sum = 0;
    for (uint32_t q = lo; q < hi; ++q) sum += block_of(q)[0];
    int total;
    int run = block_exclusive_scan(sum, warp_sums, &total);
    for (uint32_t q = lo; q < hi; ++q) {
      int16_t* blk = block_of(q);
      run += blk[0];
      blk[0] = static_cast<int16_t>(run);
    }
  }
  if (tid == 0) status[blockIdx.x] = (bad || static_cast<uint32_t>(decoded_blocks) < d.total_blocks) ? 1 : 0;
}

__global__ void __launch_bounds__(128) jpeg_idct_kernel(const oake_jpeg_desc* __restrict__ descs,
                                                        uint8_t* __restrict__ scratch) {
  const oake_jpeg_desc& d = descs[blockIdx.y];
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.total_blocks) return;
  uint32_t c = 0;
  while (c + 1 < d.ncomp && b >= d.comp[c].blocks_w * d.comp[c].blocks_h) {
    b -= d.comp[c].blocks_w * d.comp[c].blocks_h;
    ++c;
  }
  const oake_jpeg_comp& k = d.comp[c];
  const uint32_t by = b / k.blocks_w, bx = b - by * k.blocks_w;
  // the block's 64 coefficients as eight 16-byte loads
  __align__(16) int16_t coef[64];
  const uint4* src = reinterpret_cast<const uint4*>(scratch + k.coef_off + static_cast<uint64_t>(b) * 128);
#pragma unroll
  for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(coef)[i] = src[i];
  __align__(8) uint8_t px[64];
  jpeg::idct_block(coef, d.quant[k.quant], px, 8);
  const uint32_t pitch = k.blocks_w * 8;
  uint8_t* dst = scratch + k.plane_off + static_cast<uint64_t>(by * 8) * pitch + bx * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) *reinterpret_cast<uint2*>(dst + static_cast<uint64_t>(r) * pitch) = reinterpret_cast<const uint2*>(px)[r];
}

// one thread per four pixels of a row: 12 output bytes, stored as three words when they are aligned
__global__ void __launch_bounds__(256) jpeg_colour_kernel(const oake_jpeg_desc* __restrict__ descs,
                                                          const uint8_t* __restrict__ scratch,
                                                          uint8_t* __restrict__ out) {
  const oake_jpeg_desc& d = descs[blockIdx.y];
  const uint32_t quads_w = (d.width + 3) / 4;
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= quads_w * d.height) return;
  const uint32_t Y = p / quads_w, X = (p - Y * quads_w) * 4;
  const uint32_t valid = min(4u, d.width - X);
  uint8_t rgb[12];
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q) {
    if (q < valid) jpeg::pixel_rgb(d, scratch, X + q, Y, rgb + 3 * q);
    else rgb[3 * q] = rgb[3 * q + 1] = rgb[3 * q + 2] = 0;
  }
  uint8_t* dst = out + d.out_off + (static_cast<uint64_t>(Y) * d.width + X) * 3;
  if (valid == 4 && (reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
    uint32_t* w = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      w[i] = rgb[4 * i] | (static_cast<uint32_t>(rgb[4 * i + 1]) << 8) | (static_cast<uint32_t>(rgb[4 * i + 2]) << 16) |
             (static_cast<uint32_t>(rgb[4 * i + 3]) << 24);
  } else {
#pragma unroll
    for (uint32_t q = 0; q < 4; ++q) {
      if (q < valid) {
        dst[3 * q] = rgb[3 * q];
        dst[3 * q + 1] = rgb[3 * q + 1];
        dst[3 * q + 2] = rgb[3 * q + 2];
      }
    }
  }
}

}  // namespace

extern "C" {

size_t oake_jpeg_desc_bytes(void) { return sizeof(oake_jpeg_desc); }

int oake_jpeg_parse(const uint8_t* data, size_t len, oake_jpeg_desc* desc) {
  if (!data || !desc) return fail_msg("NULL argument");
  std::string why;
  const int rc = jpeg::parse(data, len, desc, &why);
  if (rc != 0) fail_msg("oake_jpeg_parse: %s", why.c_str());
  return rc;
}

size_t oake_jpeg_stream_bound(const oake_jpeg_desc* parsed) {
  return parsed ? static_cast<size_t>(jpeg::stream_bound(*parsed)) : 0;
}

int oake_jpeg_stage(const oake_jpeg_desc* parsed, const uint8_t* file, size_t len, uint8_t* dst, uint64_t stream_off,
                    uint64_t out_off, uint64_t* scratch_off, oake_jpeg_desc* placed, uint64_t* written) {
  if (!parsed || !file || !dst || !scratch_off || !placed || !written) return fail_msg("NULL argument");
  if ((parsed->ncomp != 1 && parsed->ncomp != 3) || parsed->scan_off + parsed->scan_len != len)
    return fail_msg("descriptor does not belong to this file");
  if (stream_off % 4 != 0) return fail_msg("stream_off must be a multiple of 4");
  *written = jpeg::stage(*parsed, file, dst, stream_off, out_off, scratch_off, placed);
  return 0;
}

int oake_jpeg_decode(const uint8_t* bytes, const oake_jpeg_desc* descs_host, const oake_jpeg_desc* descs_dev, int n,
                     void* scratch, uint8_t* out, int32_t* status, void* stream) {
  if (n < 0) return fail_msg("negative count");
  if (n == 0) return 0;
  if (n > 65535) return fail_msg("at most 65535 images per call");
  if (!bytes || !descs_host || !descs_dev || !scratch || !out || !status) return fail_msg("NULL buffer");
  uint32_t max_blocks = 0, max_quads = 0;
  int n_plain = 0;  // files without restart markers
  for (int i = 0; i < n; ++i) {
    const oake_jpeg_desc& d = descs_host[i];
    if ((d.ncomp != 1 && d.ncomp != 3) || d.width == 0 || d.height == 0 || d.total_blocks == 0)
      return fail_msg("descriptor %d was not produced by oake_jpeg_parse", i);
    n_plain += d.restart_interval == 0 ? 1 : 0;
    max_blocks = d.total_blocks > max_blocks ? d.total_blocks : max_blocks;
    const uint32_t quads = ((d.width + 3) / 4) * d.height;
    max_quads = quads > max_quads ? quads : max_quads;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* scr = static_cast<uint8_t*>(scratch);
  if (n_plain) jpeg_entropy_par_kernel<<<n, kParThreads, 0, st>>>(bytes, descs_dev, scr, status);
  if (n_plain < n) jpeg_entropy_rst_kernel<<<n, kRstThreads, 0, st>>>(bytes, descs_dev, scr, status);
  jpeg_idct_kernel<<<dim3((max_blocks + 127) / 128, n), 128, 0, st>>>(descs_dev, scr);
  jpeg_colour_kernel<<<dim3((max_quads + 255) / 256, n), 256, 0, st>>>(descs_dev, scr, out);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("oake_jpeg_decode launch: %s", cudaGetErrorString(e));
}

}  // extern "C"
