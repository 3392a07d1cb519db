// Handle, workspace plan and launch schedule of the OAKE tower + the extern "C" surface declared
// in include/oake_b200.h.
//
// Activation matrix layout (one row = one token, 768 wide):
//     [ B*P patch rows (crop-major) | B class rows | B side rows (T197 only) ]
// Every per-token op of a ResidualAttentionBlock (LayerNorm, QKV, out-proj, MLP) is row-wise, so
// the objects side stream (oadp/oake/objects.py:224-247) is simply B extra rows riding through the
// same launches; only attention distinguishes them.  Keeping the class / side rows contiguous at
// the end lets the last block (whose main-stream output is dead, objects.py:249-258) run its
// out-proj / MLP on those B rows alone.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/oake_b200.h"
#include "kernels.cuh"

using namespace oake;

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

}  // namespace

namespace oake {
// shared with classifier.cu: records the thread-local error string, returns 1
int fail_msg(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
}  // namespace oake

namespace {

enum KClass {
  K_FRONTEND = 0,
  K_GEMM_PATCH,
  K_ASSEMBLE,
  K_GEMM_QKV,
  K_ATTN_MAIN,
  K_ATTN_SIDE,
  K_GEMM_OUT,
  K_GEMM_FC1,
  K_GEMM_FC2,
  K_LN_POST,
  K_GEMM_PROJ,
  K_L2NORM,
  K_NUM
};
const char* kClassNames[K_NUM] = {"frontend", "gemm_patch", "assemble_ln_pre", "gemm_qkv", "attn_main",
                                  "attn_side_only", "gemm_out", "gemm_fc1", "gemm_fc2", "ln_post",
                                  "gemm_proj", "l2norm_half"};

struct LayerMaps {
  CUtensorMap qkv, out, fc1, fc2;
  CUtensorMap q, kv;  // the Q rows [0, W) and the K, V rows [W, 3W) of in_proj alone (last objects block)
};

struct ProfEvent {
  int cls;
  double flops;
  cudaEvent_t a, b;
};

constexpr size_t kAlign = 1024;
size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct Plan {
  int B, P, T, R, tail_start;
  int PB;  // rows per crop of the patch-embedding GEMM: P, or the 15 x 15 block grid of the objects tower
  size_t off_patches, off_patch_out, off_x, off_stats_a, off_stats_b, off_qkv, off_attn, off_mlp,
      off_head_in, off_emb_raw, total;
};

}  // namespace

struct oake_handle {
  int device;
  int num_sms;
  oake_weights w;
  std::vector<oake_layer_weights> layers;
  CUtensorMap tm_conv1, tm_proj;
  act_t* conv1_blk;  // conv1 weight regrouped to [O, (dy, dx), c, 16, 16] for the objects tower's block-matrix GEMM
  CUtensorMap tm_conv1_blk;
  std::vector<LayerMaps> tm_layer;
  act_t* pixel_lut;  // [3*256] ToTensor + Normalize of every byte value, rounded to act_t
  long long launches;
  bool profiling;
  std::vector<ProfEvent> events;
  std::vector<cudaEvent_t> pool;
  double acc_ms[K_NUM];
  double acc_flops[K_NUM];
  long long acc_launches[K_NUM];
};

namespace {

Plan make_plan(const oake_handle* h, int B, int variant) {
  Plan p;
  const int W = h->w.width;
  p.B = B;
  p.P = variant == OAKE_VARIANT_T197 ? 196 : 49;
  p.T = p.P + 1;
  p.R = B * p.T + (variant == OAKE_VARIANT_T197 ? B : 0);
  p.tail_start = variant == OAKE_VARIANT_T197 ? B * p.T : B * p.P;
  const bool blk = variant == OAKE_VARIANT_T197;
  p.PB = blk ? kBlockGrid * kBlockGrid : p.P;
  const size_t patch_cols = blk ? 3 * 16 * 16 : 3 * h->w.patch * h->w.patch;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return o;
  };
  p.off_patches = take(static_cast<size_t>(B) * p.PB * patch_cols * sizeof(act_t));
  p.off_patch_out = take(static_cast<size_t>(B) * p.PB * W * sizeof(float));
  p.off_x = take(static_cast<size_t>(p.R) * W * sizeof(act_t));
  p.off_stats_a = take(static_cast<size_t>(p.R) * kStatSlots * sizeof(float2));
  p.off_stats_b = take(static_cast<size_t>(p.R) * kStatSlots * sizeof(float2));
  p.off_qkv = take(static_cast<size_t>(p.R) * 3 * W * sizeof(act_t));
  p.off_attn = take(static_cast<size_t>(p.R) * W * sizeof(act_t));
  p.off_mlp = take(static_cast<size_t>(p.R) * 4 * W * sizeof(act_t));
  p.off_head_in = take(static_cast<size_t>(B) * W * sizeof(act_t));
  p.off_emb_raw = take(static_cast<size_t>(B) * h->w.out_dim * sizeof(float));
  p.total = off;
  return p;
}

cudaEvent_t get_event(oake_handle* h) {
  if (!h->pool.empty()) {
    cudaEvent_t e = h->pool.back();
    h->pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct Launcher {
  oake_handle* h;
  cudaStream_t st;
  cudaError_t err = cudaSuccess;
  const char* where = "";

  template <class F>
  void run(int cls, double flops, F&& f) {
    if (err != cudaSuccess) return;
    ProfEvent pe;
    if (h->profiling) {
      pe.cls = cls;
      pe.flops = flops;
      pe.a = get_event(h);
      pe.b = get_event(h);
      cudaEventRecord(pe.a, st);
    }
    // one NVTX range per kernel class: `ncu --nvtx --nvtx-include "gemm_fc1/"` / nsys timelines can tell the
    // QKV and c_fc launches of the one GEMM instantiation apart (header-only NVTX3: a no-op without a tool)
    nvtxRangePushA(kClassNames[cls]);
    err = f();
    nvtxRangePop();
    if (err != cudaSuccess) where = kClassNames[cls];
    h->launches += 1;
    if (h->profiling) {
      cudaEventRecord(pe.b, st);
      h->events.push_back(pe);
    }
  }
};

int check_weights(const oake_weights* w) {
  if (w == nullptr) return fail("weights is NULL");
  if (w->layers <= 0 || w->layers > 64) return fail("layers=%d out of range", w->layers);
  if (w->width != 768 || w->heads != 12 || w->patch != 32 || w->out_dim != 512 || w->image != 224)
    return fail("only ViT-B/32 geometry is built (width 768, heads 12, patch 32, out 512, image 224); got "
                "%d/%d/%d/%d/%d",
                w->width, w->heads, w->patch, w->out_dim, w->image);
  if (!w->conv1_w || !w->class_emb || !w->pos_t50 || !w->ln_pre_w || !w->ln_pre_b || !w->ln_post_w ||
      !w->ln_post_b || !w->proj_w || !w->layer)
    return fail("a required weight pointer is NULL");
  return 0;
}

}  // namespace

extern "C" {

const char* oake_last_error(void) { return g_err.c_str(); }
const char* oake_act_dtype(void) { return OAKE_ACT_NAME; }
int oake_abi_version(void) { return OAKE_ABI_VERSION; }

int oake_create(oake_handle** out, int device, const oake_weights* weights) {
  if (out == nullptr) return fail("out is NULL");
  *out = nullptr;
  if (check_weights(weights)) return 1;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail("liboake_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major,
                prop.minor);
  oake_handle* h = new oake_handle();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->w = *weights;
  h->layers.assign(weights->layer, weights->layer + weights->layers);
  h->w.layer = h->layers.data();
  h->pixel_lut = nullptr;
  h->conv1_blk = nullptr;
  h->launches = 0;
  h->profiling = false;
  memset(h->acc_ms, 0, sizeof(h->acc_ms));
  memset(h->acc_flops, 0, sizeof(h->acc_flops));
  memset(h->acc_launches, 0, sizeof(h->acc_launches));
  const int W = h->w.width;
  int rc = 0;
  rc |= make_tmap_act_2d(&h->tm_conv1, h->w.conv1_w, W, 3 * 32 * 32, gemm_block_n(W));
  rc |= make_tmap_act_2d(&h->tm_proj, h->w.proj_w, h->w.out_dim, W, gemm_block_n(h->w.out_dim));
  h->tm_layer.resize(h->w.layers);
  for (int l = 0; l < h->w.layers && rc == 0; ++l) {
    const oake_layer_weights& lw = h->layers[l];
    if (!lw.qkv_w || !lw.qkv_s || !lw.qkv_c || !lw.out_w || !lw.out_b || !lw.fc1_w || !lw.fc1_s ||
        !lw.fc1_c || !lw.fc2_w || !lw.fc2_b) {
      delete h;
      return fail("layer %d has a NULL weight pointer", l);
    }
    rc |= make_tmap_act_2d(&h->tm_layer[l].qkv, lw.qkv_w, 3 * W, W, gemm_block_n(3 * W));
    rc |= make_tmap_act_2d(&h->tm_layer[l].q, lw.qkv_w, W, W, gemm_block_n(W));
    rc |= make_tmap_act_2d(&h->tm_layer[l].kv, static_cast<const act_t*>(lw.qkv_w) + static_cast<size_t>(W) * W, 2 * W, W,
                           gemm_block_n(2 * W));
    rc |= make_tmap_act_2d(&h->tm_layer[l].out, lw.out_w, W, W, gemm_block_n(W));
    rc |= make_tmap_act_2d(&h->tm_layer[l].fc1, lw.fc1_w, 4 * W, W, gemm_block_n(4 * W));
    rc |= make_tmap_act_2d(&h->tm_layer[l].fc2, lw.fc2_w, W, 4 * W, gemm_block_n(W));
  }
  if (rc == 0 && h->w.pos_t197) {  // objects tower: conv1 regrouped for the shifted-A GEMM over the block matrix
    const size_t n = static_cast<size_t>(W) * 3 * 32 * 32;
    e = cudaMalloc(&h->conv1_blk, n * sizeof(act_t));
    if (e == cudaSuccess) e = launch_conv1_regroup(nullptr, static_cast<const act_t*>(h->w.conv1_w), h->conv1_blk, W);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess) {
      if (h->conv1_blk) cudaFree(h->conv1_blk);
      delete h;
      return fail("conv1 regroup: %s", cudaGetErrorString(e));
    }
    rc |= make_tmap_act_2d(&h->tm_conv1_blk, h->conv1_blk, W, 3 * 32 * 32, gemm_block_n(W));
  }
  if (rc != 0) {
    if (h->conv1_blk) cudaFree(h->conv1_blk);
    delete h;
    return fail("cuTensorMapEncodeTiled failed for a weight tensor (rc=%d)", rc);
  }
  {
    // torchvision ToTensor (v / 255) then Normalize ((t - mean) / std), all in fp32 like the
    // reference's transform (clip `_transform`), then the one rounding to the tensor-core type.
    const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
    const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
    std::vector<act_t> lut(768);
    for (int c = 0; c < 3; ++c)
      for (int v = 0; v < 256; ++v) {
        const float t = static_cast<float>(v) / 255.0f;
        const float n = (t - mean[c]) / stdv[c];
#ifdef OAKE_USE_BF16
        lut[c * 256 + v] = __float2bfloat16_rn(n);
#else
        lut[c * 256 + v] = __float2half_rn(n);
#endif
      }
    e = cudaMalloc(&h->pixel_lut, 768 * sizeof(act_t));
    if (e == cudaSuccess) e = cudaMemcpy(h->pixel_lut, lut.data(), 768 * sizeof(act_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      if (h->conv1_blk) cudaFree(h->conv1_blk);
      delete h;
      return fail("pixel table upload: %s", cudaGetErrorString(e));
    }
  }
  *out = h;
  return 0;
}

void oake_destroy(oake_handle* h) {
  if (h == nullptr) return;
  for (auto& pe : h->events) {
    cudaEventDestroy(pe.a);
    cudaEventDestroy(pe.b);
  }
  for (auto e : h->pool) cudaEventDestroy(e);
  if (h->pixel_lut) cudaFree(h->pixel_lut);
  if (h->conv1_blk) cudaFree(h->conv1_blk);
  delete h;
}

int oake_workspace_bytes(const oake_handle* h, int max_crops, int variant, size_t* out_bytes) {
  if (!h || !out_bytes) return fail("NULL argument");
  if (max_crops < 0) return fail("max_crops < 0");
  if (variant != OAKE_VARIANT_T50 && variant != OAKE_VARIANT_T197) return fail("bad variant %d", variant);
  *out_bytes = make_plan(h, max_crops, variant).total + kAlign;
  return 0;
}

}  // extern "C"

namespace {

// The tower.  Exactly one of `pixels` (fp32 NCHW crops) or (`arena`, `crops`) (uint8 crops) is given -- or neither
// with `prepared`: the workspace's front-end matrix has been filled by oake_resize_to_patches.
int encode_impl(oake_handle* h, const float* pixels, const uint8_t* arena, const oake_crop_src* crops, int B,
                int variant, const float* masks, void* out_f16, float* out_raw_f32, void* ws, size_t ws_bytes,
                void* stream, bool prepared = false) {
  if (!h) return fail("handle is NULL");
  if (variant != OAKE_VARIANT_T50 && variant != OAKE_VARIANT_T197) return fail("bad variant %d", variant);
  if (B < 0) return fail("B < 0");
  if (B == 0) return 0;
  if ((!prepared && !pixels && !(arena && crops)) || !out_f16 || !ws) return fail("NULL buffer");
  const bool side = variant == OAKE_VARIANT_T197;
  if (side && !masks) return fail("variant T197 needs masks");
  if (side && !h->w.pos_t197) return fail("handle was created without pos_t197");
  const Plan p = make_plan(h, B, variant);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + kAlign - 1) / kAlign * kAlign);
  if (static_cast<size_t>(base - static_cast<uint8_t*>(ws)) + p.total > ws_bytes)
    return fail("workspace too small: need %zu bytes, got %zu", p.total + kAlign, ws_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const int W = h->w.width, L = h->w.layers, H = h->w.heads, OD = h->w.out_dim;
  const int PC = 3 * h->w.patch * h->w.patch;
  act_t* patches = reinterpret_cast<act_t*>(base + p.off_patches);
  float* patch_out = reinterpret_cast<float*>(base + p.off_patch_out);
  act_t* x = reinterpret_cast<act_t*>(base + p.off_x);
  float2* stats_a = reinterpret_cast<float2*>(base + p.off_stats_a);  // rows of x after out_proj
  float2* stats_b = reinterpret_cast<float2*>(base + p.off_stats_b);  // rows of x entering a block
  act_t* qkv = reinterpret_cast<act_t*>(base + p.off_qkv);
  act_t* attn = reinterpret_cast<act_t*>(base + p.off_attn);
  act_t* mlp = reinterpret_cast<act_t*>(base + p.off_mlp);
  act_t* head_in = reinterpret_cast<act_t*>(base + p.off_head_in);
  float* emb_raw = out_raw_f32 ? out_raw_f32 : reinterpret_cast<float*>(base + p.off_emb_raw);

  const int R = p.R, ts = p.tail_start;
  CUtensorMap tm_patches, tm_x, tm_attn, tm_mlp, tm_x_tail, tm_attn_tail, tm_mlp_tail, tm_head, tm_x_side;
  int rc = 0;
  rc |= make_tmap_act_2d(&tm_patches, patches, static_cast<uint64_t>(B) * p.PB, side ? 768 : PC, 128);
  rc |= make_tmap_act_2d(&tm_x, x, R, W, 128);
  rc |= make_tmap_act_2d(&tm_attn, attn, R, W, 128);
  rc |= make_tmap_act_2d(&tm_mlp, mlp, R, 4 * W, 128);
  rc |= make_tmap_act_2d(&tm_x_tail, x + static_cast<size_t>(ts) * W, B, W, 128);
  rc |= make_tmap_act_2d(&tm_attn_tail, attn + static_cast<size_t>(ts) * W, B, W, 128);
  rc |= make_tmap_act_2d(&tm_mlp_tail, mlp + static_cast<size_t>(ts) * 4 * W, B, 4 * W, 128);
  rc |= make_tmap_act_2d(&tm_head, head_in, B, W, 128);
  tm_x_side = tm_x_tail;
  if (side) rc |= make_tmap_act_2d(&tm_x_side, x + static_cast<size_t>(B) * p.T * W, B, W, 128);
  if (rc != 0) return fail("cuTensorMapEncodeTiled failed for an activation tensor (rc=%d)", rc);

  Launcher go{h, st};
  const int ns = h->num_sms;
  // a 768-wide producer fills 6 of the 8 statistic slots: clear both tables once per call
  go.err = cudaMemsetAsync(stats_a, 0, 2 * align_up(static_cast<size_t>(R) * kStatSlots * sizeof(float2)), st);
  auto gflops = [](double m, double n, double k) { return 2.0 * m * n * k; };

  // K0/K1: crops -> conv1 patch matrix -> patch embedding -> tokens + ln_pre (+ row statistics)
  // (objects tower: the 15 x 15 block matrix instead of im2col, read at four row shifts by the GEMM -- frontend.cu)
  if (!prepared) go.run(K_FRONTEND, 0, [&] {
    if (side) return pixels ? launch_blockcol_pixels(st, pixels, patches, B) : launch_blockcol_u8(st, arena, crops, h->pixel_lut, patches, B);
    if (pixels) return launch_im2col_pixels(st, pixels, patches, B, 32, 0, 7);
    return launch_im2col_u8(st, arena, crops, h->pixel_lut, patches, B, 32, 0, 7);
  });
  {
    GemmEpilogue ep{nullptr, nullptr, nullptr, nullptr, nullptr, patch_out, W, 0, 1, 0};
    const int M = B * p.PB;
    if (side) {
      ep.a_seg_kb = 768 / 64;
      ep.a_shift[0] = 0;
      ep.a_shift[1] = 1;
      ep.a_shift[2] = kBlockGrid;
      ep.a_shift[3] = kBlockGrid + 1;
    }
    // (credited with the algorithmic work: P patches per crop, not the block grid's junk rows)
    go.run(K_GEMM_PATCH, gflops(static_cast<double>(B) * p.P, W, PC),
           [&] { return launch_gemm(st, tm_patches, side ? h->tm_conv1_blk : h->tm_conv1, M, W, PC, ep, ns); });
  }
  go.run(K_ASSEMBLE, 0, [&] {
    return launch_assemble_ln_pre(st, patch_out, h->w.class_emb, side ? h->w.pos_t197 : h->w.pos_t50,
                                  h->w.ln_pre_w, h->w.ln_pre_b, x, stats_b, B, p.P, W, side ? 1 : 0,
                                  side ? kBlockGrid : 0);
  });

  for (int l = 0; l < L; ++l) {
    const oake_layer_weights& lw = h->layers[l];
    const LayerMaps& tm = h->tm_layer[l];
    const bool last = (l == L - 1);
    // rows that still matter after this block's attention
    const int r0 = last ? ts : 0;
    const int rn = last ? B : R;
    act_t* xr = x + static_cast<size_t>(r0) * W;

    if (side && last) {
      // Last objects block: only the side row is alive behind it (objects.py:249-258), and it needs K / V of
      // every token but Q of itself alone: the patch / class rows skip the Q third of in_proj (K, V columns
      // [W, 3W) for all rows), the B side rows get their Q from a GEMM of their own.
      GemmEpilogue kv{lw.qkv_c + W, lw.qkv_s + W, stats_b, nullptr, nullptr, qkv + W, 3 * W, 0, 0, 0};
      go.run(K_GEMM_QKV, gflops(R, 2 * W, W), [&] { return launch_gemm(st, tm_x, tm.kv, R, 2 * W, W, kv, ns); });
      const int y0 = B * p.T;  // first side row
      GemmEpilogue qy{lw.qkv_c, lw.qkv_s, stats_b + static_cast<size_t>(y0) * kStatSlots, nullptr, nullptr,
                      qkv + static_cast<size_t>(y0) * 3 * W, 3 * W, 0, 0, 0};
      go.run(K_GEMM_QKV, gflops(B, W, W), [&] { return launch_gemm(st, tm_x_side, tm.q, B, W, W, qy, ns); });
    } else {  // q,k,v = ln_1(x) W^T + b   (LayerNorm folded, statistics from stats_b)
      GemmEpilogue ep{lw.qkv_c, lw.qkv_s, stats_b, nullptr, nullptr, qkv, 3 * W, 0, 0, 0};
      go.run(K_GEMM_QKV, gflops(R, 3 * W, W), [&] { return launch_gemm(st, tm_x, tm.qkv, R, 3 * W, W, ep, ns); });
    }
    if (side && last)  // only the y row of the last block is alive (objects.py:249-258)
      go.run(K_ATTN_SIDE, 4.0 * B * p.T * W, [&] { return launch_attention(st, qkv, masks, attn, B, p.P, H, 1, 1); });
    else
      go.run(K_ATTN_MAIN, 4.0 * B * p.T * p.T * W + (side ? 4.0 * B * p.T * W : 0.0),
             [&] { return launch_attention(st, qkv, masks, attn, B, p.P, H, side ? 1 : 0, 0); });
    {  // x += attn W_o^T + b ; statistics of the new x -> stats_a
      GemmEpilogue ep{lw.out_b, nullptr, nullptr, xr, stats_a + static_cast<size_t>(r0) * kStatSlots, xr, W, W, 0, 0};
      go.run(K_GEMM_OUT, gflops(rn, W, W),
             [&] { return launch_gemm(st, last ? tm_attn_tail : tm_attn, tm.out, rn, W, W, ep, ns); });
    }
    {  // u = QuickGELU(ln_2(x) W_fc^T + b)   (LayerNorm folded, statistics from stats_a)
      GemmEpilogue ep{lw.fc1_c, lw.fc1_s, stats_a + static_cast<size_t>(r0) * kStatSlots, nullptr, nullptr, mlp + static_cast<size_t>(r0) * 4 * W,
                      4 * W, 0, 0, 1};
      go.run(K_GEMM_FC1, gflops(rn, 4 * W, W),
             [&] { return launch_gemm(st, last ? tm_x_tail : tm_x, tm.fc1, rn, 4 * W, W, ep, ns); });
    }
    {  // x += u W_proj^T + b ; statistics of the new x -> stats_b (the next block's ln_1)
      GemmEpilogue ep{lw.fc2_b, nullptr, nullptr, xr, last ? nullptr : stats_b, xr, W, W, 0, 0};
      go.run(K_GEMM_FC2, gflops(rn, W, 4 * W),
             [&] { return launch_gemm(st, last ? tm_mlp_tail : tm_mlp, tm.fc2, rn, W, 4 * W, ep, ns); });
    }
  }

  // K8: ln_post(output token) @ proj -> L2 normalise -> fp16
  go.run(K_LN_POST, 0, [&] {
    return launch_layernorm(st, x + static_cast<size_t>(ts) * W, h->w.ln_post_w, h->w.ln_post_b, head_in, B, W);
  });
  {
    GemmEpilogue ep{nullptr, nullptr, nullptr, nullptr, nullptr, emb_raw, OD, 0, 1, 0};
    go.run(K_GEMM_PROJ, gflops(B, OD, W), [&] { return launch_gemm(st, tm_head, h->tm_proj, B, OD, W, ep, ns); });
  }
  go.run(K_L2NORM, 0, [&] { return launch_l2norm_half(st, emb_raw, static_cast<__half*>(out_f16), B, OD); });

  if (go.err != cudaSuccess)
    return fail("launch of %s failed: %s", go.where[0] ? go.where : "memset", cudaGetErrorString(go.err));
  return 0;
}

}  // namespace

extern "C" {

int oake_encode_pixels(oake_handle* h, const float* pixels, int B, int variant, const float* masks,
                       void* out_f16, float* out_raw_f32, void* ws, size_t ws_bytes, void* stream) {
  if (B > 0 && !pixels) return fail("pixels is NULL");
  return encode_impl(h, pixels, nullptr, nullptr, B, variant, masks, out_f16, out_raw_f32, ws, ws_bytes, stream);
}

int oake_encode_crops_u8(oake_handle* h, const uint8_t* arena, const oake_crop_src* crops, int B, int variant,
                         const float* masks, void* out_f16, float* out_raw_f32, void* ws, size_t ws_bytes,
                         void* stream) {
  if (B > 0 && (!arena || !crops)) return fail("arena / crops is NULL");
  return encode_impl(h, nullptr, arena, crops, B, variant, masks, out_f16, out_raw_f32, ws, ws_bytes, stream);
}

int oake_resize_to_patches(oake_handle* h, const uint8_t* src_arena, const oake_resize_job* jobs, int n_jobs, int variant,
                           void* ws, size_t ws_bytes, int* err_flag, void* stream) {
  if (!h) return fail("handle is NULL");
  if (variant != OAKE_VARIANT_T50 && variant != OAKE_VARIANT_T197) return fail("bad variant %d", variant);
  if (n_jobs < 0) return fail("negative count");
  if (n_jobs == 0) return 0;
  if (!src_arena || !jobs || !ws || !err_flag) return fail("NULL buffer");
  const Plan p = make_plan(h, n_jobs, variant);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + kAlign - 1) / kAlign * kAlign);
  if (static_cast<size_t>(base - static_cast<uint8_t*>(ws)) + p.total > ws_bytes)
    return fail("workspace too small: need %zu bytes, got %zu", p.total + kAlign, ws_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  nvtxRangePushA("frontend_resize_to_patches");
  cudaError_t e = launch_resize_to_matrix(st, src_arena, jobs, n_jobs, err_flag, h->pixel_lut,
                                          reinterpret_cast<act_t*>(base + p.off_patches), variant == OAKE_VARIANT_T197 ? 1 : 0);
  nvtxRangePop();
  h->launches += 3;
  if (e != cudaSuccess) return fail("resize_to_patches launch: %s", cudaGetErrorString(e));
  return 0;
}

int oake_encode_patches(oake_handle* h, int B, int variant, const float* masks, void* out_f16, float* out_raw_f32,
                        void* ws, size_t ws_bytes, void* stream) {
  return encode_impl(h, nullptr, nullptr, nullptr, B, variant, masks, out_f16, out_raw_f32, ws, ws_bytes, stream, true);
}

int oake_resize_u8(const uint8_t* src_arena, uint8_t* dst_arena, const oake_resize_job* jobs, int n_jobs,
                   int max_tiles, int* err_flag, void* stream) {
  if (n_jobs < 0 || max_tiles < 0) return fail("negative count");
  if (n_jobs == 0) return 0;
  if (!src_arena || !dst_arena || !jobs || !err_flag) return fail("NULL buffer");
  if (max_tiles > 65535 * 16) return fail("max_tiles too large");
  cudaError_t e = launch_resize_u8(static_cast<cudaStream_t>(stream), src_arena, dst_arena, jobs, n_jobs, max_tiles,
                                   err_flag);
  if (e != cudaSuccess) return fail("resize_u8 launch: %s", cudaGetErrorString(e));
  return 0;
}

int oake_object_masks(const float* fg_xyxy, const float* box_xyxy, int B, float* masks, void* stream) {
  if (B < 0) return fail("B < 0");
  if (B == 0) return 0;
  if (!fg_xyxy || !box_xyxy || !masks) return fail("NULL buffer");
  cudaError_t e = launch_object_masks(static_cast<cudaStream_t>(stream), fg_xyxy, box_xyxy, masks, B, 14);
  if (e != cudaSuccess) return fail("object_masks launch: %s", cudaGetErrorString(e));
  return 0;
}

int oake_launch_count(const oake_handle* h, long long* out) {
  if (!h || !out) return fail("NULL argument");
  *out = h->launches;
  return 0;
}

int oake_profile_enable(oake_handle* h, int enable) {
  if (!h) return fail("handle is NULL");
  h->profiling = enable != 0;
  return 0;
}

int oake_profile_collect(oake_handle* h, int cap, const char** names, double* ms, double* flops,
                         long long* launches, int* n) {
  if (!h || !n) return fail("NULL argument");
  for (auto& pe : h->events) {
    cudaError_t e = cudaEventSynchronize(pe.b);
    if (e != cudaSuccess) return fail("cudaEventSynchronize: %s", cudaGetErrorString(e));
    float t = 0.f;
    cudaEventElapsedTime(&t, pe.a, pe.b);
    h->acc_ms[pe.cls] += t;
    h->acc_flops[pe.cls] += pe.flops;
    h->acc_launches[pe.cls] += 1;
    h->pool.push_back(pe.a);
    h->pool.push_back(pe.b);
  }
  h->events.clear();
  int k = 0;
  for (int c = 0; c < K_NUM && k < cap; ++c) {
    if (names) names[k] = kClassNames[c];
    if (ms) ms[k] = h->acc_ms[c];
    if (flops) flops[k] = h->acc_flops[c];
    if (launches) launches[k] = h->acc_launches[c];
    ++k;
  }
  *n = k;
  memset(h->acc_ms, 0, sizeof(h->acc_ms));
  memset(h->acc_flops, 0, sizeof(h->acc_flops));
  memset(h->acc_launches, 0, sizeof(h->acc_launches));
  return 0;
}

// -------------------------------------------------------------------- single-kernel entry points
int oake_test_gemm(const void* A, const void* Wt, int M, int N, int K, const float* bias,
                   const float* colsum, const float* ln_stats, int act, const void* residual,
                   float* out_stats, void* out, int out_f32, int impl, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GemmEpilogue ep{bias, colsum, reinterpret_cast<const float2*>(ln_stats), static_cast<const act_t*>(residual),
                  reinterpret_cast<float2*>(out_stats), out, N, N, out_f32, act};
  if (colsum && (!ln_stats || !bias)) return fail("colsum needs ln_stats and bias (c_n)");
  if (out_stats && (!residual || N % 256 != 0 || N > 128 * kStatSlots))
    return fail("out_stats needs a residual, N %% 256 == 0 and N <= 1024");
  if (residual && (act != 0 || colsum)) return fail("the residual epilogue has no activation / LayerNorm fold");
  if (out_f32 && (colsum || residual || out_stats)) return fail("fp32 output supports bias / activation only");
  cudaError_t e;
  if (impl == 1) {
    if (colsum || out_stats) return fail("the SIMT reference has no LayerNorm fold / statistics");
    e = launch_gemm_simt(st, static_cast<const act_t*>(A), static_cast<const act_t*>(Wt), M, N, K, ep);
  } else {
    int dev = 0, ns = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&ns, cudaDevAttrMultiProcessorCount, dev);
    CUtensorMap tmA, tmW;
    if (make_tmap_act_2d(&tmA, A, M, K, 128) || make_tmap_act_2d(&tmW, Wt, N, K, gemm_block_n(N)))
      return fail("cuTensorMapEncodeTiled failed");
    e = launch_gemm(st, tmA, tmW, M, N, K, ep, ns);
  }
  if (e != cudaSuccess) return fail("gemm launch: %s", cudaGetErrorString(e));
  return 0;
}

int oake_test_layernorm(const void* x_act, const float* w, const float* b, void* out_act, int rows,
                        void* stream) {
  cudaError_t e = launch_layernorm(static_cast<cudaStream_t>(stream), static_cast<const act_t*>(x_act), w, b,
                                   static_cast<act_t*>(out_act), rows, 768);
  if (e != cudaSuccess) return fail("layernorm launch: %s", cudaGetErrorString(e));
  return 0;
}

int oake_test_attention_main(const void* qkv, void* out_act, int B, int P, void* stream) {
  cudaError_t e = launch_attention(static_cast<cudaStream_t>(stream), static_cast<const act_t*>(qkv), nullptr,
                                   static_cast<act_t*>(out_act), B, P, 12, 0, 0);
  if (e != cudaSuccess) return fail("attention_main launch: %s", cudaGetErrorString(e));
  return 0;
}

int oake_test_attention_side(const void* qkv, const float* mask, void* out_act, int B, int P, int side_only,
                             void* stream) {
  cudaError_t e = launch_attention(static_cast<cudaStream_t>(stream), static_cast<const act_t*>(qkv), mask,
                                   static_cast<act_t*>(out_act), B, P, 12, 1, side_only);
  if (e != cudaSuccess) return fail("attention (side) launch: %s", cudaGetErrorString(e));
  return 0;
}

int oake_test_im2col(const float* pixels, void* patches_act, int B, int variant, void* stream) {
  const bool side = variant == OAKE_VARIANT_T197;
  cudaError_t e = side ? launch_blockcol_pixels(static_cast<cudaStream_t>(stream), pixels, static_cast<act_t*>(patches_act), B)
                       : launch_im2col_pixels(static_cast<cudaStream_t>(stream), pixels, static_cast<act_t*>(patches_act), B, 32, 0, 7);
  if (e != cudaSuccess) return fail("im2col launch: %s", cudaGetErrorString(e));
  return 0;
}

// Patch embedding alone (front-end matrix + conv1 GEMM) from fp32 NCHW crops: out fp32 [B * PB, 768], PB = 49 rows
// per crop (T50) or the 15 x 15 block grid (T197: patch (gy, gx) is row gy * 15 + gx, the rest is scratch).
int oake_test_patch_embed(const float* pixels, const void* conv1_w_act, float* out_f32, int B, int variant,
                          void* stream) {
  if (!pixels || !conv1_w_act || !out_f32 || B <= 0) return fail("bad argument");
  const bool side = variant == OAKE_VARIANT_T197;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int W = 768, PC = 3072;
  const int PB = side ? kBlockGrid * kBlockGrid : 49;
  const size_t a_elems = static_cast<size_t>(B) * PB * (side ? 768 : PC);
  act_t *a = nullptr, *wb = nullptr;
  int dev = 0, ns = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&ns, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaMalloc(&a, a_elems * sizeof(act_t));
  if (e == cudaSuccess && side) e = cudaMalloc(&wb, static_cast<size_t>(W) * PC * sizeof(act_t));
  if (e == cudaSuccess && side) e = launch_conv1_regroup(st, static_cast<const act_t*>(conv1_w_act), wb, W);
  if (e == cudaSuccess)
    e = side ? launch_blockcol_pixels(st, pixels, a, B) : launch_im2col_pixels(st, pixels, a, B, 32, 0, 7);
  int rc = 0;
  if (e == cudaSuccess) {
    CUtensorMap tmA, tmW;
    rc = make_tmap_act_2d(&tmA, a, static_cast<uint64_t>(B) * PB, side ? 768 : PC, 128) ||
         make_tmap_act_2d(&tmW, side ? wb : conv1_w_act, W, PC, gemm_block_n(W));
    if (rc == 0) {
      GemmEpilogue ep{nullptr, nullptr, nullptr, nullptr, nullptr, out_f32, W, 0, 1, 0};
      if (side) {
        ep.a_seg_kb = 768 / 64;
        ep.a_shift[1] = 1;
        ep.a_shift[2] = kBlockGrid;
        ep.a_shift[3] = kBlockGrid + 1;
      }
      e = launch_gemm(st, tmA, tmW, B * PB, W, PC, ep, ns);
    }
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (a) cudaFree(a);
  if (wb) cudaFree(wb);
  if (rc != 0) return fail("cuTensorMapEncodeTiled failed");
  if (e != cudaSuccess) return fail("patch embedding: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
