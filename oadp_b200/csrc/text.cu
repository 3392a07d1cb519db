// CLIP ViT-B/32 TEXT tower (second half): the step that produces the classifier's E
// (oadp/prompts/vild.py:56-72, `model.encode_text(tokens)` of openai/CLIP).  Same residual blocks as the
// image tower at width 512 / 8 heads / MLP 2048, so every dense contraction runs on the tcgen05 GEMM of
// gemm.cu with the same fused epilogues (LayerNorm fold, QuickGELU, residual + row statistics); what is
// new here is small row work:
//   text_assemble_kernel     token-embedding gather + positional embedding -> act rows + row statistics
//   text_attention_kernel    causal attention, <= 77 tokens, one CTA per (sequence, head)
//   text_final_kernel        EOT row (argmax of the token ids) -> ln_final -> act row for the projection
// Parity: tests/test_gpu_text.py against oracle/text.py (itself pinned to HuggingFace CLIP).  Token ids are
// validated by the caller (oadp_b200/text.py raises on ids outside the table); the clamp in
// text_assemble_kernel only keeps a C-ABI caller that skipped that check inside the allocation.
#include <cuda_runtime.h>

#include <vector>

#include "../../include/oake_b200.h"
#include "kernels.cuh"

namespace oake {
int fail_msg(const char* fmt, ...);  // encoder.cu
}

using namespace oake;

namespace {

constexpr int kTextWidth = 512;
constexpr int kTextHeads = 8;
constexpr int kTextDh = 64;
constexpr int kTextContext = 77;
constexpr float kLnEps = 1e-5f;
constexpr size_t kAlign = 1024;
size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

// One warp per row: x[r] = token_emb[tokens[r]] + pos[r % L], rounded to act_t; statistics slot 0 = (sum, sum
// of squares) of the STORED values, the other slots zero (what the LayerNorm-folding GEMM expects).
__global__ void __launch_bounds__(256)
text_assemble_kernel(const int32_t* __restrict__ tokens, const float* __restrict__ token_emb,
                     const float* __restrict__ pos, act_t* __restrict__ x, float2* __restrict__ stats, int rows, int L,
                     int vocab) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  int tok = tokens[row];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  const float4* e4 = reinterpret_cast<const float4*>(token_emb + static_cast<size_t>(tok) * kTextWidth);
  const float4* p4 = reinterpret_cast<const float4*>(pos + static_cast<size_t>(row % L) * kTextWidth);
  uint2* o2 = reinterpret_cast<uint2*>(x + static_cast<size_t>(row) * kTextWidth);
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int j = 0; j < kTextWidth / 128; ++j) {
    const float4 a = __ldg(e4 + lane + 32 * j);
    const float4 p = __ldg(p4 + lane + 32 * j);
    uint2 u;
    u.x = pack2(a.x + p.x, a.y + p.y);
    u.y = pack2(a.z + p.z, a.w + p.w);
    o2[lane + 32 * j] = u;
    const float2 lo = unpack2(u.x), hi = unpack2(u.y);
    s += (lo.x + lo.y) + (hi.x + hi.y);
    ss += lo.x * lo.x + lo.y * lo.y + hi.x * hi.x + hi.y * hi.y;
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (lane < kStatSlots) stats[static_cast<size_t>(row) * kStatSlots + lane] = lane == 0 ? make_float2(s, ss) : make_float2(0.f, 0.f);
}

// Causal attention of one (sequence, head): K and V of the <= 77 tokens in shared memory (rows padded to 66
// halves so that lanes reading different keys hit different banks), one warp per query row, scores /
// softmax in fp32.  qkv rows are [q | k | v], head h = columns 64h .. 64h+63 of each third.
constexpr int kKvPitch = kTextDh + 2;

__global__ void __launch_bounds__(128)
text_attention_kernel(const act_t* __restrict__ qkv, act_t* __restrict__ out, int L) {
  __shared__ act_t k_s[kTextContext * kKvPitch];
  __shared__ act_t v_s[kTextContext * kKvPitch];
  __shared__ float q_s[4][kTextDh];
  __shared__ float p_s[4][kTextContext + 3];
  const int b = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row0 = static_cast<size_t>(b) * L;
  const int ld = 3 * kTextWidth;
  for (int i = threadIdx.x; i < L * (kTextDh / 2); i += blockDim.x) {
    const int t = i / (kTextDh / 2), c = (i - t * (kTextDh / 2)) * 2;
    const act_t* src = qkv + (row0 + t) * ld + h * kTextDh + c;
    *reinterpret_cast<uint32_t*>(&k_s[t * kKvPitch + c]) = *reinterpret_cast<const uint32_t*>(src + kTextWidth);
    *reinterpret_cast<uint32_t*>(&v_s[t * kKvPitch + c]) = *reinterpret_cast<const uint32_t*>(src + 2 * kTextWidth);
  }
  __syncthreads();
  for (int t = warp; t < L; t += 4) {
    {  // the query row, scaled by 1 / sqrt(64) as openai/CLIP (nn.MultiheadAttention) does
      const act_t* q = qkv + (row0 + t) * ld + h * kTextDh;
      const float2 f = unpack2(*reinterpret_cast<const uint32_t*>(q + 2 * lane));
      q_s[warp][2 * lane] = f.x * 0.125f;
      q_s[warp][2 * lane + 1] = f.y * 0.125f;
    }
    __syncwarp();
    // scores of keys j = lane, lane + 32, lane + 64 (j <= t: causal)
    float sc[3];
    float mx = -INFINITY;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int j = lane + 32 * g;
      float acc = -INFINITY;
      if (j <= t) {
        acc = 0.f;
#pragma unroll 8
        for (int c = 0; c < kTextDh; c += 2) {
          const float2 kv = unpack2(*reinterpret_cast<const uint32_t*>(&k_s[j * kKvPitch + c]));
          acc = fmaf(q_s[warp][c], kv.x, acc);
          acc = fmaf(q_s[warp][c + 1], kv.y, acc);
        }
      }
      sc[g] = acc;
      mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int j = lane + 32 * g;
      const float e = j <= t ? __expf(sc[g] - mx) : 0.f;
      if (j < L) p_s[warp][j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    // out[d] for d = 2 lane, 2 lane + 1
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j <= t; ++j) {
      const float p = p_s[warp][j];
      const float2 vv = unpack2(*reinterpret_cast<const uint32_t*>(&v_s[j * kKvPitch + 2 * lane]));
      o0 = fmaf(p, vv.x, o0);
      o1 = fmaf(p, vv.y, o1);
    }
    const float inv = 1.0f / sum;
    *reinterpret_cast<uint32_t*>(out + (row0 + t) * kTextWidth + h * kTextDh + 2 * lane) = pack2(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

// One warp per sequence: the row of the EOT token (first position of the largest token id, torch.argmax)
// through ln_final (fp32 statistics) to an act row for the projection GEMM.
__global__ void __launch_bounds__(256)
text_final_kernel(const int32_t* __restrict__ tokens, const act_t* __restrict__ x, const float* __restrict__ w,
                  const float* __restrict__ bias, act_t* __restrict__ head_in, int B, int L) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  int best = -1, at = 0;
  for (int t = lane; t < L; t += 32) {
    const int v = tokens[static_cast<size_t>(b) * L + t];
    if (v > best) {
      best = v;
      at = t;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int ob = __shfl_xor_sync(0xffffffffu, best, o), oa = __shfl_xor_sync(0xffffffffu, at, o);
    if (ob > best || (ob == best && oa < at)) {
      best = ob;
      at = oa;
    }
  }
  const uint2* x2 = reinterpret_cast<const uint2*>(x + (static_cast<size_t>(b) * L + at) * kTextWidth);
  float4 v[kTextWidth / 128];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kTextWidth / 128; ++j) {
    const uint2 u = x2[lane + 32 * j];
    const float2 lo = unpack2(u.x), hi = unpack2(u.y);
    v[j] = make_float4(lo.x, lo.y, hi.x, hi.y);
    s += (lo.x + lo.y) + (hi.x + hi.y);
  }
  const float mean = warp_sum(s) * (1.0f / kTextWidth);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kTextWidth / 128; ++j) {
    v[j].x -= mean;
    v[j].y -= mean;
    v[j].z -= mean;
    v[j].w -= mean;
    ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / kTextWidth) + kLnEps);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(bias);
  uint2* o2 = reinterpret_cast<uint2*>(head_in + static_cast<size_t>(b) * kTextWidth);
#pragma unroll
  for (int j = 0; j < kTextWidth / 128; ++j) {
    const float4 g = __ldg(w4 + lane + 32 * j);
    const float4 be = __ldg(b4 + lane + 32 * j);
    uint2 u;
    u.x = pack2(v[j].x * rstd * g.x + be.x, v[j].y * rstd * g.y + be.y);
    u.y = pack2(v[j].z * rstd * g.z + be.z, v[j].w * rstd * g.w + be.w);
    o2[lane + 32 * j] = u;
  }
}

struct TextLayerMaps {
  CUtensorMap qkv, out, fc1, fc2;
};

struct TextPlan {
  int R;
  size_t off_x, off_stats_a, off_stats_b, off_qkv, off_attn, off_mlp, off_head_in, total;
};

TextPlan make_text_plan(int B, int L) {
  TextPlan p;
  p.R = B * L;
  const int W = kTextWidth;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += align_up(bytes);
    return o;
  };
  p.off_x = take(static_cast<size_t>(p.R) * W * sizeof(act_t));
  p.off_stats_a = take(static_cast<size_t>(p.R) * kStatSlots * sizeof(float2));
  p.off_stats_b = take(static_cast<size_t>(p.R) * kStatSlots * sizeof(float2));
  p.off_qkv = take(static_cast<size_t>(p.R) * 3 * W * sizeof(act_t));
  p.off_attn = take(static_cast<size_t>(p.R) * W * sizeof(act_t));
  p.off_mlp = take(static_cast<size_t>(p.R) * 4 * W * sizeof(act_t));
  p.off_head_in = take(static_cast<size_t>(B) * W * sizeof(act_t));
  p.total = off;
  return p;
}

}  // namespace

struct oake_text_handle {
  int device;
  int num_sms;
  oake_text_weights w;
  std::vector<oake_layer_weights> layers;
  std::vector<TextLayerMaps> tm_layer;
  CUtensorMap tm_proj;
  long long launches;
};

extern "C" {

int oake_text_create(oake_text_handle** out, int device, const oake_text_weights* w) {
  if (!out) return fail_msg("out is NULL");
  *out = nullptr;
  if (!w) return fail_msg("weights is NULL");
  if (w->layers <= 0 || w->layers > 64) return fail_msg("layers=%d out of range", w->layers);
  if (w->width != kTextWidth || w->heads != kTextHeads || w->out_dim != 512 || w->context <= 0 ||
      w->context > kTextContext || w->vocab <= 0)
    return fail_msg("only the ViT-B/32 text geometry is built (width 512, heads 8, context <= 77, out 512); got "
                    "width %d heads %d context %d out %d", w->width, w->heads, w->context, w->out_dim);
  if (!w->token_emb || !w->pos || !w->ln_final_w || !w->ln_final_b || !w->proj_w || !w->layer)
    return fail_msg("a required weight pointer is NULL");
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail_msg("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail_msg("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail_msg("liboake_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  oake_text_handle* h = new oake_text_handle();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->w = *w;
  h->launches = 0;
  const int W = kTextWidth;
  int rc = 0;
  for (int l = 0; l < w->layers; ++l) {
    const oake_layer_weights& lw = w->layer[l];
    if (!lw.qkv_w || !lw.qkv_s || !lw.qkv_c || !lw.out_w || !lw.out_b || !lw.fc1_w || !lw.fc1_s || !lw.fc1_c ||
        !lw.fc2_w || !lw.fc2_b) {
      delete h;
      return fail_msg("layer %d has a NULL weight pointer", l);
    }
    h->layers.push_back(lw);
    TextLayerMaps tm;
    rc |= make_tmap_act_2d(&tm.qkv, lw.qkv_w, 3 * W, W, gemm_block_n(3 * W));
    rc |= make_tmap_act_2d(&tm.out, lw.out_w, W, W, gemm_block_n(W));
    rc |= make_tmap_act_2d(&tm.fc1, lw.fc1_w, 4 * W, W, gemm_block_n(4 * W));
    rc |= make_tmap_act_2d(&tm.fc2, lw.fc2_w, W, 4 * W, gemm_block_n(W));
    h->tm_layer.push_back(tm);
  }
  h->w.layer = h->layers.data();
  rc |= make_tmap_act_2d(&h->tm_proj, w->proj_w, w->out_dim, W, gemm_block_n(w->out_dim));
  if (rc != 0) {
    delete h;
    return fail_msg("cuTensorMapEncodeTiled failed for a weight tensor (rc=%d)", rc);
  }
  *out = h;
  return 0;
}

void oake_text_destroy(oake_text_handle* h) { delete h; }

int oake_text_workspace_bytes(const oake_text_handle* h, int max_sequences, int length, size_t* out_bytes) {
  if (!h || !out_bytes) return fail_msg("NULL argument");
  if (max_sequences < 0) return fail_msg("max_sequences < 0");
  if (length <= 0 || length > h->w.context) return fail_msg("length %d outside [1, %d]", length, h->w.context);
  *out_bytes = make_text_plan(max_sequences, length).total + kAlign;
  return 0;
}

int oake_encode_text(oake_text_handle* h, const int32_t* tokens, int B, int L, float* out_f32, void* ws,
                     size_t ws_bytes, void* stream) {
  if (!h) return fail_msg("handle is NULL");
  if (B < 0) return fail_msg("B < 0");
  if (B == 0) return 0;
  if (L <= 0 || L > h->w.context) return fail_msg("length %d outside [1, %d]", L, h->w.context);
  if (!tokens || !out_f32 || !ws) return fail_msg("NULL buffer");
  const TextPlan p = make_text_plan(B, L);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + kAlign - 1) / kAlign * kAlign);
  if (static_cast<size_t>(base - static_cast<uint8_t*>(ws)) + p.total > ws_bytes)
    return fail_msg("workspace too small: need %zu bytes, got %zu", p.total + kAlign, ws_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int W = kTextWidth, R = p.R, OD = h->w.out_dim, ns = h->num_sms;
  act_t* x = reinterpret_cast<act_t*>(base + p.off_x);
  float2* stats_a = reinterpret_cast<float2*>(base + p.off_stats_a);  // rows of x after out_proj
  float2* stats_b = reinterpret_cast<float2*>(base + p.off_stats_b);  // rows of x entering a block
  act_t* qkv = reinterpret_cast<act_t*>(base + p.off_qkv);
  act_t* attn = reinterpret_cast<act_t*>(base + p.off_attn);
  act_t* mlp = reinterpret_cast<act_t*>(base + p.off_mlp);
  act_t* head_in = reinterpret_cast<act_t*>(base + p.off_head_in);

  CUtensorMap tm_x, tm_attn, tm_mlp, tm_head;
  int rc = 0;
  rc |= make_tmap_act_2d(&tm_x, x, R, W, 128);
  rc |= make_tmap_act_2d(&tm_attn, attn, R, W, 128);
  rc |= make_tmap_act_2d(&tm_mlp, mlp, R, 4 * W, 128);
  rc |= make_tmap_act_2d(&tm_head, head_in, B, W, 128);
  if (rc != 0) return fail_msg("cuTensorMapEncodeTiled failed for an activation tensor (rc=%d)", rc);

  cudaError_t err = cudaSuccess;
  const char* where = "";
  auto run = [&](const char* name, cudaError_t e) {
    if (err == cudaSuccess && e != cudaSuccess) {
      err = e;
      where = name;
    }
    h->launches += 1;
  };
  // a 512-wide producer fills 4 of the 8 statistic slots: clear both tables once per call
  run("memset", cudaMemsetAsync(stats_a, 0, 2 * align_up(static_cast<size_t>(R) * kStatSlots * sizeof(float2)), st));
  text_assemble_kernel<<<(R + 7) / 8, 256, 0, st>>>(tokens, h->w.token_emb, h->w.pos, x, stats_b, R, L, h->w.vocab);
  run("text_assemble", cudaGetLastError());
  for (int l = 0; l < h->w.layers && err == cudaSuccess; ++l) {
    const oake_layer_weights& lw = h->layers[l];
    const TextLayerMaps& tm = h->tm_layer[l];
    {  // q,k,v = ln_1(x) W^T + b   (LayerNorm folded, statistics from stats_b)
      GemmEpilogue ep{lw.qkv_c, lw.qkv_s, stats_b, nullptr, nullptr, qkv, 3 * W, 0, 0, 0};
      run("gemm_qkv", launch_gemm(st, tm_x, tm.qkv, R, 3 * W, W, ep, ns));
    }
    text_attention_kernel<<<dim3(B, kTextHeads), 128, 0, st>>>(qkv, attn, L);
    run("text_attention", cudaGetLastError());
    {  // x += attn W_o^T + b ; statistics of the new x -> stats_a
      GemmEpilogue ep{lw.out_b, nullptr, nullptr, x, stats_a, x, W, W, 0, 0};
      run("gemm_out", launch_gemm(st, tm_attn, tm.out, R, W, W, ep, ns));
    }
    {  // u = QuickGELU(ln_2(x) W_fc^T + b)   (LayerNorm folded, statistics from stats_a)
      GemmEpilogue ep{lw.fc1_c, lw.fc1_s, stats_a, nullptr, nullptr, mlp, 4 * W, 0, 0, 1};
      run("gemm_fc1", launch_gemm(st, tm_x, tm.fc1, R, 4 * W, W, ep, ns));
    }
    {  // x += u W_proj^T + b ; statistics of the new x -> stats_b (the next block's ln_1)
      GemmEpilogue ep{lw.fc2_b, nullptr, nullptr, x, stats_b, x, W, W, 0, 0};
      run("gemm_fc2", launch_gemm(st, tm_mlp, tm.fc2, R, W, 4 * W, ep, ns));
    }
  }
  text_final_kernel<<<(B + 7) / 8, 256, 0, st>>>(tokens, x, h->w.ln_final_w, h->w.ln_final_b, head_in, B, L);
  run("text_final", cudaGetLastError());
  {  // e = ln_final(x[eot]) @ text_projection, fp32 out (the caller normalises and averages the templates)
    GemmEpilogue ep{nullptr, nullptr, nullptr, nullptr, nullptr, out_f32, OD, 0, 1, 0};
    run("gemm_text_projection", launch_gemm(st, tm_head, h->tm_proj, B, OD, W, ep, ns));
  }
  if (err != cudaSuccess) return fail_msg("launch of %s failed: %s", where, cudaGetErrorString(err));
  return 0;
}

}  // extern "C"
