/*
 * cre.h — C-ABI of libcre_b200.so: the clip-embedding + re-ID hot path on B200 (sm_100a).
 *
 * The reference (UBC-AWP/vision-sam3-yolo-lameless) has no FFI: its hot path is Python that calls
 * HuggingFace transformers / torch / a remote Qdrant server.  Each entry point below replaces the
 * arithmetic behind one reference call site (file:line relative to the reference tree; "HF:" =
 * transformers 5.5.0, models/dinov3_vit/):
 *
 *   cre_preprocess_patchify  <- services/dinov3-pipeline/app/main.py:98-107 (BGR->RGB, PIL, HF processor)
 *                               HF:image_processing_dinov3_vit.py:45-86 (rescale, antialiased bilinear
 *                               resize, normalise) + the im2col half of HF:modeling_dinov3_vit.py:71-81
 *   cre_vit_forward          <- services/dinov3-pipeline/app/main.py:110-113 (model(**inputs),
 *                               last_hidden_state.mean(dim=1)); HF:modeling_dinov3_vit.py:60-92,153-200,
 *                               238-343,381-386,424-450,530-555
 *   cre_pool_clips           <- services/dinov3-pipeline/app/main.py:204-208 (np.mean over frames) and
 *                               services/tracking-service/app/reid/matcher.py:124 (e / (||e|| + 1e-8))
 *   cre_gallery_topk         <- services/tracking-service/app/reid/matcher.py:127-132 and
 *     + cre_merge_topk          services/dinov3-pipeline/app/main.py:168-172 (Qdrant COSINE search, limit=k)
 *
 * Conventions: every function returns 0 on success, a negative code on failure (-1 bad argument,
 * -2 CUDA error, -3 unsupported shape); cre_last_error() returns the thread-local message.  Hot calls
 * are asynchronous on the given stream (a cudaStream_t passed as void*), never allocate and never
 * synchronise; the caller owns every buffer.  All pointers named *_dev are device pointers.
 * A cre_ctx is bound to one device and must not be shared between host threads.
 */
#ifndef CRE_H_
#define CRE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRE_ABI_VERSION 2
#define CRE_TOPK_MAX 8      /* candidates one scan pass keeps per query */
#define CRE_TOPK_LIMIT 256  /* largest k of cre_gallery_topk / cre_merge_topk (k > CRE_TOPK_MAX: ceil(k / 8) scan passes) */

typedef struct cre_ctx cre_ctx;

/* HF DINOv3ViTConfig fields the forward pass needs (HF:configuration_dinov3_vit.py:74-101). */
typedef struct cre_model_cfg {
    int32_t hidden;       /* 768 (ViT-B/16) | 1024 (ViT-L/16); multiple of 64 */
    int32_t layers;       /* 12 | 24 */
    int32_t heads;        /* 12 | 16; hidden / heads must be 64 */
    int32_t mlp;          /* 3072 | 4096 */
    int32_t patch;        /* 16 */
    int32_t registers;    /* 4 register tokens; prefix = 1 + registers */
    float rope_theta;     /* 100.0 */
    float ln_eps;         /* 1e-5 */
} cre_model_cfg;

/* Tensor kinds inside the packed weight blob.  Per-layer kinds take layer in [0, layers); global kinds
 * take layer = -1.  Matrices are bf16 row-major [out, in] (= nn.Linear.weight); vectors are fp32. */
enum cre_weight_kind {
    CRE_W_PATCH = 0,      /* bf16 [hidden, 3*patch*patch]  conv weight .view(hidden, -1)            */
    CRE_B_PATCH = 1,      /* f32  [hidden]                                                          */
    CRE_PREFIX = 2,       /* f32  [1 + registers, hidden]  cls token then register tokens            */
    CRE_LN_F_G = 3,       /* f32  [hidden] final norm weight                                         */
    CRE_LN_F_B = 4,       /* f32  [hidden] final norm bias                                           */
    CRE_LN1_G = 5, CRE_LN1_B = 6,
    CRE_W_QKV = 7,        /* bf16 [3*hidden, hidden]  rows: q_proj, k_proj, v_proj                   */
    CRE_B_QKV = 8,        /* f32  [3*hidden]  (k part is zero: key_bias = False)                     */
    CRE_W_O = 9, CRE_B_O = 10,
    CRE_LS1 = 11,         /* f32  [hidden] layer_scale1.lambda1                                      */
    CRE_LN2_G = 12, CRE_LN2_B = 13,
    CRE_W_UP = 14, CRE_B_UP = 15,     /* bf16 [mlp, hidden], f32 [mlp]                               */
    CRE_W_DOWN = 16, CRE_B_DOWN = 17, /* bf16 [hidden, mlp], f32 [hidden]                            */
    CRE_LS2 = 18,
    CRE_WEIGHT_KINDS = 19
};

const char* cre_last_error(void);
int32_t cre_abi_version(void);

/* Packed weight blob layout: byte size, and byte offset / element count of one tensor. */
int64_t cre_packed_weights_bytes(const cre_model_cfg* cfg);
int64_t cre_weight_offset(const cre_model_cfg* cfg, int32_t layer, int32_t kind);
int64_t cre_weight_elems(const cre_model_cfg* cfg, int32_t layer, int32_t kind);

/* Context: keeps the device weight pointer, cached TMA descriptors, RoPE and resize-weight tables. */
int32_t cre_create(const cre_model_cfg* cfg, const void* packed_weights_dev, int32_t device, cre_ctx** out);
int32_t cre_destroy(cre_ctx* ctx);

/* Bytes of scratch cre_vit_forward needs for `frames` frames of grid_h x grid_w patches. */
int64_t cre_workspace_bytes(const cre_model_cfg* cfg, int32_t frames, int32_t grid_h, int32_t grid_w);

/* K1. uint8 NHWC frames -> antialiased bilinear resize to (resize_h, resize_w) -> (x/255 - mean)/std
 * -> 16x16 patchify of the top-left (resize_h/16)*(resize_w/16) patches -> bf16
 * out[n * P, 3*256], K index = c*256 + ky*16 + kx with c in RGB order.
 * frames_dev: n frames, each h rows of row_pitch bytes (>= 3*w), frame_pitch bytes apart.
 * bgr != 0: channel order in memory is B,G,R (cv2), swapped on the fly. */
int32_t cre_preprocess_patchify(cre_ctx* ctx, const uint8_t* frames_dev, int32_t n, int32_t h, int32_t w,
                                int64_t row_pitch, int64_t frame_pitch, int32_t bgr, int32_t resize_h,
                                int32_t resize_w, const float mean[3], const float std_[3],
                                void* out_patches_dev, void* stream);

/* K1 on regions of interest (per-track crops, SURVEY.md section 8(f) #3: services/tracking-service/app/main.py:332-334 "in
 * production, you'd extract per-track embeddings").  rois_dev: int32 [n_rois, 5] = {frame, x0, y0, x1, y1} (pixel box, x1 / y1
 * exclusive, inside the frame, at least 1 x 1).  ROI r is processed exactly like the image frames[frame][y0:y1, x0:x1] would be
 * by cre_preprocess_patchify (the HF processor resizes every crop to resize_h x resize_w) and lands in patch rows
 * [r * P, (r + 1) * P).  The per-ROI antialias tables are built on the device into scratch_dev
 * (cre_roi_scratch_bytes(n_rois, h, w, resize_h, resize_w) bytes, 256-byte aligned). */
int64_t cre_roi_scratch_bytes(int32_t n_rois, int32_t h, int32_t w, int32_t resize_h, int32_t resize_w);
int32_t cre_preprocess_patchify_roi(cre_ctx* ctx, const uint8_t* frames_dev, int32_t n_frames, int32_t h, int32_t w,
                                    int64_t row_pitch, int64_t frame_pitch, int32_t bgr, const int32_t* rois_dev, int32_t n_rois,
                                    int32_t resize_h, int32_t resize_w, const float mean[3], const float std_[3],
                                    void* scratch_dev, int64_t scratch_bytes, void* out_patches_dev, void* stream);

/* K2 + K3a. bf16 patch rows [n * grid_h * grid_w, 3*256] -> ViT forward -> final LayerNorm ->
 * mean over ALL tokens (cls + registers + patches) -> out_frame_emb_dev f32 [n, hidden].
 * If out_tokens_dev != NULL the final-normed hidden state f32 [n, T, hidden] is also written. */
int32_t cre_vit_forward(cre_ctx* ctx, const void* patches_dev, int32_t n, int32_t grid_h, int32_t grid_w,
                        void* workspace_dev, int64_t workspace_bytes, float* out_frame_emb_dev,
                        float* out_tokens_dev, void* stream);

/* K3b. Per-clip mean of frame embeddings + L2 normalisation.  clip_offsets_dev: int32 [clips + 1],
 * frames of clip c are rows [off[c], off[c+1]).  out_mean_dev f32 [clips, dim] (raw mean, what the
 * reference upserts), out_unit_dev f32 [clips, dim] = mean / (||mean||_2 + 1e-8). Either may be NULL. */
int32_t cre_pool_clips(const float* frame_emb_dev, const int32_t* clip_offsets_dev, int32_t clips,
                       int32_t dim, float* out_mean_dev, float* out_unit_dev, void* stream);

/* K4. Cosine re-ID against one gallery shard.  queries_dev f32 [q, dim] (L2-normalised by the caller via
 * cre_pool_clips), gallery_dev bf16 [rows, dim] row-major, L2-normalised rows.  Scores are
 * fp32-query x bf16-gallery dot products accumulated in fp32: q <= 2 (one message = one query, the reference's call pattern)
 * streams the shard once on the CUDA cores with the fp32 query in registers, one launch; larger q runs a tcgen05 tile GEMM with the
 * queries split hi+lo bf16 internally.  All device pointers 16-byte aligned; frame_emb / out_* of cre_pool_clips likewise.
 * Writes the k best (score desc, index asc) per query: out_scores_dev f32 [q, k], out_idx_dev i32 [q, k]
 * with idx = row_base + local row; missing entries (rows < k) are (-inf, INT32_MAX).  1 <= k <= CRE_TOPK_LIMIT: the reference's
 * callers pass any limit (main.py:165, matcher.py:104-108, gnn-pipeline main.py:52-58); k <= CRE_TOPK_MAX is one pass over the shard,
 * larger k repeats the scan ceil(k / 8) times, each pass admitting only candidates that rank after the previous pass's last entry.
 * scratch_dev: cre_gallery_scratch_bytes(q, dim, k) bytes, 256-byte aligned.  dump_scores_dev (optional, may be NULL):
 * f32 [q, rows] full score matrix for parity tests. */
int64_t cre_gallery_scratch_bytes(int32_t q, int32_t dim, int32_t k);
int32_t cre_gallery_topk(cre_ctx* ctx, const float* queries_dev, int32_t q, int32_t dim,
                         const void* gallery_dev, int32_t rows, int32_t row_base, int32_t k,
                         void* scratch_dev, int64_t scratch_bytes, float* out_scores_dev,
                         int32_t* out_idx_dev, float* dump_scores_dev, void* stream);

/* Merge `lists` candidate lists per query (e.g. one per gallery shard after the all-gather):
 * scores_dev f32 [lists, q, k], idx_dev i32 [lists, q, k] -> best k by (score desc, index asc).  k <= CRE_TOPK_LIMIT;
 * lists * k <= 4096 when k > CRE_TOPK_MAX. */
int32_t cre_merge_topk(const float* scores_dev, const int32_t* idx_dev, int32_t lists, int32_t q, int32_t k,
                       float* out_scores_dev, int32_t* out_idx_dev, void* stream);

/* Gallery maintenance (matcher.py:203-255 create_identity, :257-301 momentum update):
 * v = normalise( momentum * old + (1 - momentum) * unit_query ); momentum = 0 writes the query.  master_dev f32 [rows, dim]
 * (may be NULL) is the full-precision copy of the gallery -- the vector the reference keeps in Qdrant, blends into and upserts
 * (matcher.py:267-301): old = its row, and it receives v.  gallery_dev bf16 [rows, dim], the copy cre_gallery_topk scans, receives
 * bf16(v); with master_dev == NULL old is read from the bf16 row. */
int32_t cre_gallery_update_row(void* gallery_dev, float* master_dev, int32_t dim, int32_t row, const float* unit_query_dev,
                               float momentum, void* stream);

/* ---- building blocks exported for the parity tests (same kernels the calls above launch) ---------- */
enum cre_gemm_epilogue {
    CRE_EPI_BF16 = 0,  /* out_bf16 = acc + bias                         */
    CRE_EPI_F32 = 1,   /* out_f32  = acc + bias                         */
    CRE_EPI_GELU = 3,  /* out_bf16 = gelu_erf(acc + bias)               */
    CRE_EPI_RESID = 4, /* out_f32 += scale * (acc + bias)   (in place)  */
    CRE_EPI_NONE = 7,  /* accumulators dropped: main-loop timing only   */
    CRE_EPI_RESID_LN = 8, /* cre_gemm_ln only: x += scale * (acc + bias), bf16(x - pivot), LayerNorm statistics */
    CRE_EPI_RESID_LN3 = 9, /* same results; one x tile in flight per warp, one pipeline stage more (long K) */
    CRE_EPI_RESID_SP = 10, /* cre_gemm_ln only: the residual stream as two bf16 halves around the row pivot (what cre_vit_forward runs) */
    CRE_EPI_RESID_SP3 = 11 /* same results; one chunk slot per warp, more pipeline stages (long K) */
};
/* D[m, n] = A[m, k] (bf16 row-major) * B[n, k]^T (bf16 row-major); k % 64 == 0; n % 64 == 0 for the bf16-out
 * epilogues, n % 32 == 0 for the fp32-out ones; out_dev 16-byte aligned (written by TMA).
 * cta_group = 1 or 2 (CTA pair, cta_group::2). bias/scale may be NULL where unused. */
int32_t cre_gemm_bf16(cre_ctx* ctx, const void* a_dev, const void* b_dev, int32_t m, int32_t n, int32_t k,
                      int32_t epilogue, const float* bias_dev, const float* scale_dev, void* out_dev,
                      int32_t cta_group, void* stream);
/* ---- LayerNorm folded into the GEMMs (what cre_vit_forward runs inside the blocks; HF:modeling_dinov3_vit.py:411-448
 * norm1 -> attention, norm2 -> mlp).  LN(x) W^T + b = rstd * ((x - pivot) W'^T - (mean - pivot) c1) + c2 with
 * W' = W * gamma, c1 = row sums of W', c2 = b + W beta.  A statistics row is 2 * (dim / 128) + 4 floats:
 * [pivot, -, -, -, (mean_i, M2_i) of every 128-column slot]; dim = 768 | 1024.
 *   cre_row_stats:       x f32 [rows, dim] -> out_xb bf16 [rows, dim] = x - mean, out_stats (pivot = mean).
 *   cre_fold_ln_weights: w bf16 [n, k], gamma/beta f32 [k], bias f32 [n] or NULL -> out_w bf16 [n, k], out_c1/out_c2 f32 [n];
 *                        rows [0, scaled_rows) of all three outputs are multiplied by row_scale (head_dim^-0.5 on the q rows).
 *   cre_gemm_ln:         epilogue CRE_EPI_BF16 | CRE_EPI_GELU: out bf16 [m, n] = (gelu)(LN-folded A B^T), A = the centred bf16
 *                        rows, bias = c2, stats_in = their statistics rows (ln_dim = k);
 *                        epilogue CRE_EPI_RESID_LN (n == ln_dim, n % 256 == 0): out f32 [m, n] (in place) += scale * (A B^T + bias),
 *                        out_xb bf16 [m, n] = out - pivot (pivot = the row mean recorded in stats_in), stats_out = new rows;
 *                        epilogue CRE_EPI_RESID_SP (same shape rules): the residual stream x = pivot + hi + lo is held as
 *                        hi = out_xb bf16 [m, n] and lo = out bf16 [m, n] (pivot = slot 0 of its statistics row); both are updated
 *                        in place to the halves of x + scale * (A B^T + bias) around the NEW pivot (the row mean stats_in
 *                        records), stats_out = new rows.  8 bytes of HBM traffic per element instead of RESID_LN's 10.
 *   cre_row_stats_split: cre_row_stats + out_lo bf16 [rows, dim] = bf16((x - mean) - out_hi): seeds the split stream. */
int32_t cre_row_stats(const float* x_dev, int32_t rows, int32_t dim, void* out_xb_dev, float* out_stats_dev, void* stream);
int32_t cre_row_stats_split(const float* x_dev, int32_t rows, int32_t dim, void* out_hi_dev, void* out_lo_dev, float* out_stats_dev,
                            void* stream);
int32_t cre_fold_ln_weights(const void* w_dev, const float* gamma_dev, const float* beta_dev, const float* bias_dev, int32_t n,
                            int32_t k, int32_t scaled_rows, float row_scale, void* out_w_dev, float* out_c1_dev, float* out_c2_dev,
                            void* stream);
int32_t cre_gemm_ln(cre_ctx* ctx, const void* a_dev, const void* b_dev, int32_t m, int32_t n, int32_t k, int32_t epilogue,
                    const float* bias_dev, const float* c1_dev, const float* scale_dev, const float* stats_in_dev, int32_t ln_dim,
                    float ln_eps, void* out_dev, void* out_xb_dev, float* stats_out_dev, int32_t cta_group, void* stream);
/* out bf16 [rows, dim] = LayerNorm(x f32 [rows, dim]) * gamma + beta */
int32_t cre_layernorm_bf16(const float* x_dev, const float* gamma_dev, const float* beta_dev, int32_t rows,
                           int32_t dim, float eps, void* out_dev, void* stream);
/* Non-causal attention over frames on the fused projection matrix qkv bf16 [n*t, ld]: q at column head*64
 * (pre-scaled by 1/8, rotary already applied), k at k_col0 + head*64 (rotary applied), v at v_col0 + head*64;
 * out bf16 [n*t, heads*64].  ld and the column offsets are multiples of 8 elements. */
int32_t cre_attention(cre_ctx* ctx, const void* qkv_dev, int32_t ld, int32_t k_col0, int32_t v_col0, int32_t n,
                      int32_t t, int32_t heads, void* out_dev, void* scratch_dev, int64_t scratch_bytes, void* stream);
/* scratch of cre_attention (the overflow flags of its single-pass softmax: a unit whose scores outrun the fixed stabiliser is
 * recomputed exactly by a second kernel in the same call, csrc/attention.cu "Exactness"); 256-byte aligned. */
int64_t cre_attention_scratch_bytes(int32_t n, int32_t heads);

/* ---- launch accounting -------------------------------------------------------------------------------
 * cre_kernel_launches: kernels launched by this library in this process so far (bench.py's gpu_launches).
 * cre_profile_start(max): from now on bracket every launch with a CUDA-event pair on its stream (at most `max`
 * launches are recorded).  cre_profile_stop: stop recording, wait for the recorded launches and return their
 * count; ids_out[i] is a cre_kernel_id, ms_out[i] the device time of that launch, work_out[i] its algorithmic
 * work (FLOPs for GEMM / attention kernels, bytes for memory-bound kernels).  Synchronises; not a hot call. */
enum cre_kernel_id {
    CRE_K_PREPROCESS = 0, CRE_K_FILL_PREFIX = 1, CRE_K_GEMM_PATCH = 2, CRE_K_LAYERNORM = 3, CRE_K_GEMM_QKV = 4,
    CRE_K_ATTENTION = 5, CRE_K_GEMM_RESID = 6, CRE_K_GEMM_GELU = 7, CRE_K_FINAL_NORM_MEAN = 8, CRE_K_POOL_CLIPS = 9,
    CRE_K_SPLIT_HI_LO = 10, CRE_K_FILL_TOPK = 11, CRE_K_GEMM_TOPK = 12, CRE_K_MERGE_TOPK = 13, CRE_K_GEMM_PLAIN = 14,
    CRE_K_GALLERY_UPDATE = 15, CRE_K_ROW_STATS = 16, CRE_K_FOLD_LN = 17, CRE_K_ROI_TABLES = 18, CRE_K_ATTENTION_EXACT = 19,
    CRE_K_GEMM_RESID_MLP = 20,   /* residual-update GEMMs with K > N (MLP down projection); CRE_K_GEMM_RESID = the attention-out projection */
    CRE_KERNEL_IDS = 21
};
int64_t cre_kernel_launches(void);
int32_t cre_profile_start(int32_t max_launches);
int32_t cre_profile_stop(int32_t* ids_out, float* ms_out, double* work_out, int32_t cap);

/* Process-wide tuning knob: 1 = one CTA per 128x256 tile (tcgen05 cta_group::1), 2 = CTA pairs on
 * 256x256 tiles (cta_group::2) for the ViT GEMMs.  Results are identical either way. */
int32_t cre_set_cta_group(int32_t cta_group);
/* Generic tuning knobs for the benchmark harness: "cta_group" (1 | 2), "gemm_stages" (0 = default, 3..6:
 * TMA pipeline depth of the cre_gemm_bf16 building block), "attention_fast" (1 = persistent TMEM-resident kernel for
 * T <= 256, default; 0 = general kernel), "ln_fold" (1 = LayerNorm folded into the GEMMs, default; 0 = separate LayerNorm
 * launches), "resid_ln_deep" (bit 0 / bit 1: attention-out / MLP-down projection use the one-slot form CRE_EPI_RESID_LN3 / _SP3),
 * "resid_split" (1 = residual stream as two bf16 halves, CRE_EPI_RESID_SP, default; 0 = fp32 stream + bf16 copy, CRE_EPI_RESID_LN), "attention_split" (1 = split-S
 * kernel for 160 < T <= 208, default), "attention_poly" (0 | 1 | 2: share of that kernel's exponentials on the FMA pipe),
 * "attention_split_mode" (bit 0: direct global stores of O, bit 1: PV of the second key half in one piece), "attention_split_delay"
 * (SM cycles by which the second query-tile group of that kernel trails the first), "attention_long" (1 = persistent key-block
 * kernel for every T outside the split-S range, default; 0 = the round-1 kernels), "preprocess_tma" (2 = TMA-staged K1, row pairs, default; 1 = TMA-staged,
 * one row per warp; 0 = direct-load kernel), "preprocess_identity", "scan_small".  Every setting gives results inside the parity
 * tolerances (the three K1 settings: identical bits).  Unknown keys return -1. */
int32_t cre_set_tuning(const char* key, int32_t value);

#ifdef __cplusplus
}
#endif
#endif /* CRE_H_ */
