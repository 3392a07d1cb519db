/* oake_b200 -- C ABI of the B200-native OAKE hot path (liboake_b200.so).
 *
 * The reference (LutingWang/OADP) is pure Python and has NO FFI of its own for this path; the
 * boundary it exposes is the Python call `model.encode_image(x)` / `model.visual(objects, masks)`
 * on a `clip.model.CLIP` object and `fc_cls(x)` on the mmdet linear layer.  Each entry point below
 * names the reference interface it stands behind.  A maintainer binds these with ctypes (see
 * INTEGRATION.md); `oadp_b200/binding.py` is that binding.
 *
 * Conventions: plain pointers and sizes, no torch types.  Every buffer (weights, inputs, outputs,
 * workspace) is caller-owned DEVICE memory unless a parameter says "host".  Calls are
 * stream-ordered on the `stream` argument (a cudaStream_t passed as void*), never synchronise and
 * never allocate.  Return value 0 = ok, non-zero = error with a thread-local message available
 * from oake_last_error(); nothing throws across the ABI.  One handle per (device, stream); a
 * handle is not thread-safe.
 */
#ifndef OAKE_B200_H_
#define OAKE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OAKE_ABI_VERSION 1

/* Tower variants.
 * T50 : un-modified CLIP ViT-B/32, conv1 stride 32, 7x7+1 tokens
 *       (oadp/oake/globals.py:57, oadp/oake/blocks.py:129).
 * T197: objects surgery -- conv1 stride 16 / padding 15, 14x14+1 tokens, mask-attended CLS side
 *       stream whose token replaces the transformer output (oadp/oake/objects.py:198-314). */
#define OAKE_VARIANT_T50 0
#define OAKE_VARIANT_T197 1

typedef struct oake_handle oake_handle;

/* Device pointers for one ResidualAttentionBlock (openai/CLIP model.py state-dict names).
 * "act" = the tensor-core element type reported by oake_act_dtype() ("f16" unless built with
 * -DOAKE_USE_BF16); linear weights are stored [out_features, in_features] row-major as in the
 * checkpoint.
 *
 * ln_1 and ln_2 are FOLDED into the GEMM that consumes them (the library never runs them as
 * separate kernels).  For y = LN(x; gamma, beta) W^T + b the caller supplies
 *     w' = W * diag(gamma)            rounded to act          (qkv_w / fc1_w)
 *     s  = row sums of w' AS STORED   fp32 [out_features]     (qkv_s / fc1_s)
 *     c  = W beta + b                 fp32 [out_features]     (qkv_c / fc1_c)
 * and the kernel evaluates  y[m,n] = rstd_m * (x_m . w'_n - mean_m * s_n) + c_n  with mean/rstd of
 * the stored residual row.  `oadp_b200.model.fold_layernorm` builds these from a checkpoint. */
typedef struct {
  const void* qkv_w;  /* attn.in_proj_weight * diag(ln_1.weight)        [2304,768] act */
  const float* qkv_s; /* [2304] */
  const float* qkv_c; /* attn.in_proj_weight @ ln_1.bias + in_proj_bias [2304] */
  const void* out_w;  /* attn.out_proj.weight     [768,768] act  */
  const float* out_b; /* attn.out_proj.bias       [768]          */
  const void* fc1_w;  /* mlp.c_fc.weight * diag(ln_2.weight)            [3072,768] act */
  const float* fc1_s; /* [3072] */
  const float* fc1_c; /* mlp.c_fc.weight @ ln_2.bias + c_fc.bias        [3072] */
  const void* fc2_w;  /* mlp.c_proj.weight        [768,3072] act */
  const float* fc2_b; /* mlp.c_proj.bias          [768]          */
} oake_layer_weights;

typedef struct {
  int32_t layers;  /* 12 */
  int32_t width;   /* 768 */
  int32_t heads;   /* 12 */
  int32_t patch;   /* 32 */
  int32_t out_dim; /* 512 */
  int32_t image;   /* 224 */
  const void* conv1_w;     /* conv1.weight reshaped [768, 3*32*32] act, columns (c,ky,kx) */
  const float* class_emb;  /* class_embedding [768] */
  const float* pos_t50;    /* positional_embedding [50,768] */
  const float* pos_t197;   /* resampled table [197,768] (objects.py:293-296) or NULL */
  const float* ln_pre_w;
  const float* ln_pre_b;
  const float* ln_post_w;
  const float* ln_post_b;
  const void* proj_w;               /* proj^T [512,768] act */
  const oake_layer_weights* layer;  /* HOST array of `layers` entries (device pointers inside) */
} oake_weights;

/* Replaces `clip.load_default(...)` + `.cuda()` (globals.py:47, blocks.py:123, objects.py:290):
 * binds caller-owned device weights, builds the TMA descriptors for them.  The weights must
 * outlive the handle. */
int oake_create(oake_handle** out, int device, const oake_weights* weights);
void oake_destroy(oake_handle* h);

/* Bytes of caller-provided scratch for a batch of up to `max_crops` crops of `variant`. */
int oake_workspace_bytes(const oake_handle* h, int max_crops, int variant, size_t* out_bytes);

/* Replaces `model.encode_image(x)` (variant T50; globals.py:57, blocks.py:129) and
 * `model.visual(o, m)` (variant T197; objects.py:330) PLUS the caller's `F.normalize(...).half()`
 * (globals.py:58-59, blocks.py:130-133, objects.py:331-334).
 *   pixels      : device fp32 [B,3,224,224], CLIP-normalised (what the reference's transform yields)
 *   masks       : device fp32 [B,1,14,14], 1 = background (objects.py:129-155); NULL for T50
 *   out_f16     : device fp16 [B,512]  L2-normalised embedding (the value the reference stores)
 *   out_raw_f32 : device fp32 [B,512]  un-normalised tower output (what encode_image returns), or NULL
 *   ws/ws_bytes : scratch of at least oake_workspace_bytes(B, variant) */
int oake_encode_pixels(oake_handle* h, const float* pixels, int B, int variant, const float* masks,
                       void* out_f16, float* out_raw_f32, void* ws, size_t ws_bytes, void* stream);

/* ---- GPU front end: the work the reference does in DataLoader workers with PIL -------------- */

/* One crop-and-resize: Pillow `image.crop(box).resize((out_w, out_h), BICUBIC)` restricted to the
 * window [win_x, win_x+win_w) x [win_y, win_y+win_h) of the resized image (torchvision CenterCrop),
 * bit-exact on uint8.  Stands where the reference calls PIL through torchvision
 * Resize(224, BICUBIC) + CenterCrop(224) (oadp/oake/globals.py:32, blocks.py:80-81,95,
 * objects.py:116-127) and `image.resize((int(w/1.5), int(h/1.5)))` (blocks.py:73-77).
 * Images are uint8 HWC, 3 channels.  The rectangle may leave the image: outside reads 0, like
 * PIL's crop.  Limits: box_w/out_w and box_h/out_h <= 11 (oake_resize_u8 reports an error above). */
typedef struct {
  int64_t src_off;      /* byte offset of the source image inside the source arena */
  int64_t dst_off;      /* byte offset of the output window inside the destination arena */
  int32_t src_w, src_h; /* source image size, pixels */
  int32_t src_pitch_px; /* source row pitch, pixels */
  int32_t box_x0, box_y0, box_w, box_h; /* crop rectangle in source pixels (after PIL's rounding) */
  int32_t out_w, out_h; /* size the crop is resized to */
  int32_t win_x, win_y, win_w, win_h; /* window of the resized crop that is written */
  int32_t dst_pitch_px; /* destination row pitch, pixels */
} oake_resize_job;

/* jobs: DEVICE array.  max_tiles = max over jobs of ceil(win_w/32)*ceil(win_h/32).  err_flag: device
 * int, set to 1 if a job exceeded the kernel's limits (caller clears and checks it). */
int oake_resize_u8(const uint8_t* src_arena, uint8_t* dst_arena, const oake_resize_job* jobs, int n_jobs,
                   int max_tiles, int* err_flag, void* stream);

/* The same crop-and-resize with the pixels written straight into the tower's front-end matrix inside `ws`
 * (ToTensor + Normalize through the exact fp32 table, rounded once to the tensor-core type): job i becomes crop i of
 * a following oake_encode_patches(h, n_jobs, variant, ...) on the same workspace and stream.  Every job's window must
 * be the whole 224 x 224 crop (win_x = win_y = 0 of a 224 x 224 resize, or the CenterCrop window); dst_off and
 * dst_pitch_px are ignored.  This is the reference's transform + `encode_image` / `visual` call without the uint8
 * crop in between (oadp/oake/globals.py:32,57; objects.py:116-127,330). */
int oake_resize_to_patches(oake_handle* h, const uint8_t* src_arena, const oake_resize_job* jobs, int n_jobs,
                           int variant, void* ws, size_t ws_bytes, int* err_flag, void* stream);
/* The tower on the front-end matrix that oake_resize_to_patches left in `ws` (same B = n_jobs, same variant). */
int oake_encode_patches(oake_handle* h, int B, int variant, const float* masks, void* out_f16, float* out_raw_f32,
                        void* ws, size_t ws_bytes, void* stream);

/* Where a 224x224 uint8 HWC crop lives: arena + off, rows pitch_px pixels apart.  Block crops of
 * oadp/oake/blocks.py:79-81 need no copy at all: they are windows into the pyramid level. */
typedef struct {
  int64_t off;
  int32_t pitch_px;
  int32_t reserved;
} oake_crop_src;

/* oake_encode_pixels for crops that are still uint8: ToTensor + Normalize (CLIP mean/std, exact
 * fp32 table) are fused into the patch gather.  crops: DEVICE array of B entries. */
int oake_encode_crops_u8(oake_handle* h, const uint8_t* arena, const oake_crop_src* crops, int B, int variant,
                         const float* masks, void* out_f16, float* out_raw_f32, void* ws, size_t ws_bytes,
                         void* stream);

/* The 14x14 foreground masks of oadp/oake/objects.py:129-155 for B crops.  box_xyxy: the expanded
 * squares (float, before PIL's rounding), fg_xyxy: proposal - box.lt; masks: fp32 [B,1,14,14], 1 =
 * background.  All device pointers. */
int oake_object_masks(const float* fg_xyxy, const float* box_xyxy, int B, float* masks, void* stream);

/* ---- cosine classifier of oadp.dp (oadp/dp/classifiers.py:19-112, oadp/dp/utils.py:47-51) -----
 * Two ops, split where the reference splits its modules, because the todd distiller hooks capture
 * the OUTPUT of `fc_cls._linear` (configs/dp/models/ *.py): h must exist as a tensor and receives
 * gradient from both the logits and the distillation loss.  All pointers device, fp32, row-major.
 *   h      = F.normalize(x W^T + b)                                     NormalizedLinear.forward
 *   logits = alpha * (h E^T) - shift, columns [ninf_lo, ninf_hi) = -inf BaseClassifier/Classifier/
 *            E = [text (num_all,512) as stored ; F.normalize(bg) if bg != NULL]   ViLDClassifier.forward
 * logits has row pitch k_pad (multiple of 128, >= K = num_all + (bg != NULL)); columns >= K are
 * padding.  in_features must be a multiple of 64 (256 / 1024 in the reference configs). */
int oake_classifier_workspace_bytes(int N, int in_features, int k_pad, size_t* out_bytes);
int oake_normalized_linear_fwd(const float* x, const float* w, const float* b, int N, int in_features, float* h,
                               float* inv_norm, void* ws, size_t ws_bytes, void* stream);
/* dx [N,in], dw [512,in], db [512] may each be NULL. */
int oake_normalized_linear_bwd(const float* x, const float* w, const float* h, const float* inv_norm,
                               const float* dh, int N, int in_features, float* dx, float* dw, float* db, void* ws,
                               size_t ws_bytes, void* stream);
int oake_cosine_logits_fwd(const float* h, const float* text, const float* bg, int N, int num_all, int k_pad,
                           float alpha, float shift, int ninf_lo, int ninf_hi, float* logits, void* ws,
                           size_t ws_bytes, void* stream);
/* dlogits [N,k_pad] (entries of -inf / padding columns are ignored); dh [N,512]; dbg [512] or NULL. */
int oake_cosine_logits_bwd(const float* h, const float* text, const float* bg, const float* dlogits, int N,
                           int num_all, int k_pad, float alpha, int ninf_lo, int ninf_hi, float* dh, float* dbg,
                           void* ws, size_t ws_bytes, void* stream);

/* Inference fast path (no gradient): both halves in ONE call on prepared operands -- what
 * `bbox_head.fc_cls(x)` costs at test time (oadp/dp/roi_heads.py:64-112 calls it twice per image).
 * oake_classifier_prepare casts W [512,in] (-> w_act) and / or packs E = [text ; normalised bg ; zero padding]
 * (-> e_act [k_pad,512]) into the tensor-core type once per weight version (either half may be NULL).
 * oake_classifier_fwd: x [N,in] in x_dtype (OAKE_DTYPE_*; the tensor-core type passes through without a copy),
 * writes h [N,512] fp32 (the tensor the distiller hooks read) and logits [N,k_pad] fp32: 3 launches
 * (GEMM, row normalise, GEMM), 4 with a cast of x.  Workspace: oake_classifier_workspace_bytes. */
#define OAKE_DTYPE_F32 0
#define OAKE_DTYPE_F16 1
#define OAKE_DTYPE_BF16 2
int oake_classifier_prepare(const float* w, int in_features, const float* text, const float* bg, int num_all,
                            int k_pad, void* w_act, void* e_act, void* stream);
int oake_classifier_fwd(const void* x, int x_dtype, const void* w_act, const float* bias, const void* e_act, int N,
                        int in_features, int k_pad, float alpha, float shift, int ninf_lo, int ninf_hi, float* h,
                        float* logits, void* ws, size_t ws_bytes, void* stream);

/* ---- ViLD ensemble scoring (oadp/dp/roi_heads.py:93-112, ViLDEnsembleRoIHead._bbox_forward) ------
 * out = log( softmax(bbox_logits)^lambda * softmax(object_logits)^(1-lambda) ), last column replaced
 * by log(1 - sum of the others).  All fp32 device pointers; logits (N, K1 = num_all + 1) with row
 * pitches ld_* >= K1 (the k_pad pitch of oake_cosine_logits_fwd is accepted as is); lambda [K1] is
 * the `_lambda` buffer of the reference (2/3 base, 1/3 novel + background, roi_heads.py:55-59).
 * -inf logits (ObjectMixin forces the background column to -inf, bbox_heads.py:57-60) score 0. */
int oake_vild_ensemble(const float* bbox_logits, const float* object_logits, const float* lambda, int N, int K1,
                       int ld_bbox, int ld_object, float* out, int ld_out, void* stream);

/* ---- re-softmax + multiclass NMS (second half): mmdet `BBoxHead.get_bboxes` ->
 * `multiclass_nms(bboxes, softmax(cls_score), score_thr, nms, max_per_img)` behind the ensemble scores, and the
 * same call in oadp/dp/test_nni.py:55-92.  Class-agnostic boxes (`reg_class_agnostic=True` in the reference's
 * configs): boxes (N,4) fp32 xyxy shared by all K foreground classes, 16-byte aligned, N <= 4096.
 *   oake_softmax_rows    out[n][:K1] = softmax(in[n][:K1])                       (row pitches ld_*)
 *   oake_multiclass_nms  keep[k][n] (uint8, K x N) = 1 iff candidate (box n, class k) has score > score_thr and
 *                        survives greedy NMS (IoU > iou_thr suppresses, descending scores, ties by box index)
 *                        among the candidates of class k.  scores (N, ld_scores): column k = class k (the
 *                        background column is simply not among the first K).  ws: oake_nms_workspace_bytes(N). */
int oake_softmax_rows(const float* in, int N, int K1, int ld_in, float* out, int ld_out, void* stream);
int oake_nms_workspace_bytes(int N, size_t* out_bytes);
int oake_multiclass_nms(const float* boxes_xyxy, const float* scores, int N, int K, int ld_scores, float score_thr,
                        float iou_thr, uint8_t* keep, void* ws, size_t ws_bytes, void* stream);

/* ---- distillation-side losses: value and gradient w.r.t. the first argument in one call -
 * All fp32 device pointers.  `scale` = weight / numel for reduction='mean', weight for 'sum' (todd
 * BaseLoss.reduce; the configs' WarmupScheduler is a scalar at a given step).  loss: 1 float.  grad:
 * same shape as the first argument or NULL.  ws: oake_loss_workspace_bytes(rkd_rows) bytes (rkd_rows = 0
 * unless oake_rkd_loss is called).  Sums are fixed-order: results are deterministic.
 *   oake_pair_loss       kind 0: todd L1Loss, kind 1: todd MSELoss on (pred, target)
 *                        (configs/dp/models/{vild_ensemble_faster_rcnn_r50_fpn,block,global_}.py)
 *   oake_rkd_loss        RKDLoss, oadp/base/losses.py:68-108: MSE of s s^T - t t^T, rows (N, dim);
 *                        scale = weight / N^2 for 'mean'
 *   oake_asymmetric_loss AsymmetricLoss, oadp/base/losses.py:10-65: x probabilities, y uint8 0/1 */
int oake_loss_workspace_bytes(int rkd_rows, size_t* out_bytes);
int oake_pair_loss(const float* pred, const float* target, long long n, int kind, float scale, float* loss,
                   float* grad, void* ws, size_t ws_bytes, void* stream);
int oake_rkd_loss(const float* s, const float* t, int N, int dim, float scale, float* loss, float* grad_s, void* ws,
                  size_t ws_bytes, void* stream);
int oake_asymmetric_loss(const float* x, const uint8_t* y, long long n, float gamma_neg, float gamma_pos,
                         float clip, float eps, float scale, float* loss, float* grad, void* ws, size_t ws_bytes,
                         void* stream);

/* ---- GPU JPEG decode (the step in front of the resize: oadp/oake/base.py:53,
 * `PIL.Image.open(...).convert('RGB')`) -----------------------------------------------------------
 * Baseline / extended-sequential 8-bit Huffman JPEG, one interleaved scan, grayscale or YCbCr with
 * 4:4:4, 4:2:2 (2x1) or 4:2:0 (2x2) sampling -> uint8 HWC RGB, bit-identical to what Pillow
 * (libjpeg-turbo: islow IDCT, "fancy" triangle chroma upsampling, 16-bit fixed-point YCbCr->RGB)
 * returns for the same file.  The host parses the headers into a descriptor (oake_jpeg_parse, no
 * GPU needed) and copies the entropy-coded segment, minus its stuffing bytes, into the staging
 * buffer (oake_jpeg_stage); the entropy decode, the IDCT and the colour conversion run on the GPU
 * (oake_jpeg_decode).  Files outside that envelope (progressive,
 * arithmetic, CMYK, 12-bit, other sampling, multi-scan) make oake_jpeg_parse return
 * OAKE_JPEG_UNSUPPORTED; the caller then decodes that file with Pillow, as the reference does. */
#define OAKE_JPEG_UNSUPPORTED 2

typedef struct {
  uint16_t look[512];  /* 9-bit prefix -> (code length << 8) | symbol; 0 = code longer than 9 bits */
  int32_t maxcode[18]; /* largest code of each length 1..16 (-1 = none); [17] = sentinel */
  int32_t valoff[17];  /* symbol index = code + valoff[length] */
  uint8_t huffval[256];
} oake_jpeg_huff;

typedef struct {
  uint32_t h, v;               /* sampling factors as used (1,1 for a single-component scan) */
  uint32_t blocks_w, blocks_h; /* 8x8 blocks stored per row / column (padded to whole MCUs) */
  uint32_t width, height;      /* real samples of the component (libjpeg downsampled_width / _height) */
  uint32_t dc_tbl, ac_tbl, quant, _pad;
  uint64_t coef_off;  /* int16 coefficients, [block][64] in zig-zag (file) order: BYTE offset into scratch */
  uint64_t plane_off; /* uint8 samples after the IDCT, row pitch blocks_w * 8: byte offset into scratch */
} oake_jpeg_comp;

typedef struct {
  uint32_t width, height;
  uint32_t ncomp; /* 1 (grayscale) or 3 (YCbCr) */
  uint32_t hmax, vmax;
  uint32_t mcus_x, mcus_y;
  uint32_t restart_interval; /* MCUs, 0 = none */
  uint32_t total_blocks;     /* over all components */
  uint32_t sync_slots;       /* capacity of the subsequence table at sync_off (parallel entropy decode) */
  uint64_t scan_off;      /* entropy-coded bytes: byte offset into the `bytes` arena ... */
  uint64_t scan_len;      /* ... and how many there are up to the end of the file */
  uint64_t out_off;       /* RGB HWC output (width * height * 3 bytes): byte offset into `out` */
  uint64_t scratch_bytes; /* coefficients + planes + subsequence table of this image */
  uint64_t sync_off;      /* subsequence table, 24 bytes per slot: byte offset into scratch */
  uint32_t restart_count; /* restart intervals of the scan: ceil(MCUs / restart_interval), 0 = none */
  uint32_t _pad;
  oake_jpeg_comp comp[3];
  uint16_t quant[4][64]; /* natural (row-major) order */
  oake_jpeg_huff dc[2], ac[2];
} oake_jpeg_desc;

size_t oake_jpeg_desc_bytes(void); /* sizeof(oake_jpeg_desc), for bindings that keep it opaque */
/* HOST call, no GPU: parses `data` (one whole JPEG file, host memory).  Offsets in `desc` are
 * relative (scan_off to the start of the file, coef/plane offsets to this image's scratch, out_off
 * 0) until oake_jpeg_stage.  Returns 0, OAKE_JPEG_UNSUPPORTED (desc->width / height still valid
 * when the frame header was reached), or 1 = malformed (message in oake_last_error). */
int oake_jpeg_parse(const uint8_t* data, size_t len, oake_jpeg_desc* desc);
/* HOST calls: oake_jpeg_stage copies the entropy-coded segment of `file` (the `len` bytes that were
 * parsed into `parsed`) to `dst` -- normally pinned memory -- without the 0x00 bytes the format stuffs
 * after every 0xFF, cut at the end-of-image marker and zero-padded to a multiple of 4 plus 16 bytes,
 * followed -- for a file with restart markers -- by a uint32 table of the byte offset at which each
 * restart interval starts (0xFFFFFFFF for an interval whose marker is missing); at most
 * oake_jpeg_stream_bound(parsed) bytes, the count is returned in *written.  `placed` receives
 * the descriptor rebased onto the caller's arenas: scan_off = stream_off (the offset `dst` will have
 * in the device `bytes` arena, a multiple of 4), out_off, and coefficient / plane offsets moved behind
 * *scratch_off, which is advanced by the image's (256-byte aligned) scratch size. */
size_t oake_jpeg_stream_bound(const oake_jpeg_desc* parsed);
int oake_jpeg_stage(const oake_jpeg_desc* parsed, const uint8_t* file, size_t len, uint8_t* dst, uint64_t stream_off,
                    uint64_t out_off, uint64_t* scratch_off, oake_jpeg_desc* placed, uint64_t* written);
/* Decodes n images.  descs_host / descs_dev: the same n staged descriptors in host memory (read
 * during the call, for grid sizes) and in device memory; bytes: device copy of the staged streams;
 * scratch: device, >= the final *scratch_off of oake_jpeg_stage; out: device arena the RGB pixels
 * go to (the `src_arena` of oake_resize_u8); status: device int32[n], 0 = ok, non-zero = the
 * entropy-coded data of that image was damaged or truncated (its pixels are then undefined). */
int oake_jpeg_decode(const uint8_t* bytes, const oake_jpeg_desc* descs_host, const oake_jpeg_desc* descs_dev, int n,
                     void* scratch, uint8_t* out, int32_t* status, void* stream);

/* ---- CLIP text tower (second half; oadp/prompts/vild.py:56-72 `model.encode_text(tokens)`) --
 * openai/CLIP ViT-B/32 text encoder: width 512, 8 heads, MLP 2048, causal attention over <= 77 tokens,
 * ln_final, the row of the EOT token (argmax of the ids) times text_projection.  Layer weights use
 * oake_layer_weights with the text shapes (qkv [1536,512], out [512,512], fc1 [2048,512], fc2 [512,2048];
 * ln_1 / ln_2 folded as for the image tower).  Parity: tests/test_gpu_text.py. */
typedef struct oake_text_handle oake_text_handle;

typedef struct {
  int32_t layers;  /* 12 */
  int32_t width;   /* 512 */
  int32_t heads;   /* 8 */
  int32_t vocab;   /* 49408 */
  int32_t context; /* rows of `pos`, <= 77 */
  int32_t out_dim; /* 512 */
  const float* token_emb;  /* token_embedding.weight   [vocab, 512] fp32 */
  const float* pos;        /* positional_embedding     [context, 512] fp32 */
  const float* ln_final_w; /* [512] */
  const float* ln_final_b; /* [512] */
  const void* proj_w;      /* text_projection^T        [512, 512] act, row = output feature */
  const oake_layer_weights* layer; /* host array [layers] of device pointers; copied at create */
} oake_text_weights;

int oake_text_create(oake_text_handle** out, int device, const oake_text_weights* weights);
void oake_text_destroy(oake_text_handle* h);
int oake_text_workspace_bytes(const oake_text_handle* h, int max_sequences, int length, size_t* out_bytes);
/* tokens: device int32 [B, L] (L <= context; SOT ... EOT, zero padded, as clip.tokenize produces them);
 * out_f32: device fp32 [B, out_dim], NOT normalised (the prompt builder normalises per template and
 * averages, prompts/vild.py:66-71). */
int oake_encode_text(oake_text_handle* h, const int32_t* tokens, int B, int L, float* out_f32, void* ws,
                     size_t ws_bytes, void* stream);

/* Error string of the last failing call on this thread ("" if none). */
const char* oake_last_error(void);
/* "f16" or "bf16": element type of `act` tensors. */
const char* oake_act_dtype(void);
int oake_abi_version(void);

/* ---- instrumentation (bench.py / tests) ------------------------------------------------- */
/* Number of kernels this handle has launched since creation. */
int oake_launch_count(const oake_handle* h, long long* out);
/* When enabled every launch is bracketed by CUDA events on its stream (bench roofline pass). */
int oake_profile_enable(oake_handle* h, int enable);
/* Synchronises the recorded events, accumulates per kernel class, clears the event list.
 * Writes up to `cap` entries; returns the number of classes in *n.  flops = algorithmic FLOPs
 * issued by that class (0 for non-GEMM classes). */
int oake_profile_collect(oake_handle* h, int cap, const char** names, double* ms, double* flops,
                         long long* launches, int* n);

/* ---- single-kernel entry points (unit tests; all pointers device, row-major) --------------- */
/* out[M,N] = epi(A[M,K] * W[N,K]^T), the tcgen05 GEMM with every epilogue feature:
 *   bias fp32 [N] | NULL; colsum fp32 [N] + ln_stats fp32 [M,8,2] (8 partial (sum x, sum x^2)
 *   pairs per row, added in order) enable the LayerNorm fold | NULL; act 0|1 (QuickGELU);
 *   residual act [M,N] | NULL (may alias out; excludes act / fold); out_stats fp32 [M,8,2]: slot j
 *   receives (sum y, sum y^2) over columns [128j, 128j+128) of the output row (needs residual) | NULL;
 *   out act or fp32 (fp32: bias/act only).
 *   impl 0 = tcgen05, 1 = CUDA-core reference (bias/act/residual only). */
int oake_test_gemm(const void* A, const void* W, int M, int N, int K, const float* bias,
                   const float* colsum, const float* ln_stats, int act, const void* residual,
                   float* out_stats, void* out, int out_f32, int impl, void* stream);
/* Stand-alone LayerNorm of act rows [rows,768] (ln_post). */
int oake_test_layernorm(const void* x_act, const float* w, const float* b, void* out_act, int rows,
                        void* stream);
/* qkv act [R,2304], rows [B*P | B | (B)]; out act [R,768]. */
int oake_test_attention_main(const void* qkv, void* out_act, int B, int P, void* stream);
/* Objects attention: main stream + the side token y in one kernel (side_only = 0), or only the B
 * side rows (side_only = 1, last block).  qkv act [B*(P+2),2304], mask fp32 [B,P]. */
int oake_test_attention_side(const void* qkv, const float* mask, void* out_act, int B, int P, int side_only,
                             void* stream);
/* Front-end matrix from fp32 NCHW crops: variant T50 -> im2col act [B*49, 3072]; variant T197 -> the block matrix
 * act [B*225, 768] (15 x 15 blocks of 16 x 16 pixels of the crop zero-padded by 15, column order (c, ky, kx)). */
int oake_test_im2col(const float* pixels, void* patches_act, int B, int variant, void* stream);
/* Patch embedding alone: out fp32 [B*49, 768] (T50) or [B*225, 768] (T197; patch (gy, gx) = row gy*15+gx of a crop).
 * conv1_w_act: act [768, 3072], the checkpoint layout (oadp/oake/objects.py:299-301 changes stride / padding only). */
int oake_test_patch_embed(const float* pixels, const void* conv1_w_act, float* out_f32, int B, int variant,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OAKE_B200_H_ */
