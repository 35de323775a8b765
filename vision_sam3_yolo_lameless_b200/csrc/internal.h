// Host-side launcher declarations shared by the translation units of libcre_b200.so.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cre.h"

namespace cre {

void set_error(const char* fmt, ...);
const char* last_error();
int gemm_workers(int m, int n, int cg, int num_sms);
void set_gemm_stages(int stages);
#ifdef CRE_TUNING
void set_gemm_debug(int mode);
#endif
void set_attention_fast(int on);
void set_attention_split(int on);
void set_attention_long(int on);
void set_attention_poly(int v);
void set_attention_split_mode(int v);
void set_attention_split_delay(int cycles);
void set_attention_trace(unsigned long long* buf);

// Kernel ids reported by cre_profile_stop (include/cre.h enum cre_kernel_id)
// RAII bracket around one kernel launch: bumps the launch counter and, when the profiler is on, records an
// event pair on `stream`.  `work` = algorithmic FLOPs (GEMM / attention) or bytes (memory-bound kernels).
class LaunchScope {
public:
    LaunchScope(int id, double work, cudaStream_t stream);
    ~LaunchScope();
    LaunchScope(const LaunchScope&) = delete;
    LaunchScope& operator=(const LaunchScope&) = delete;
private:
    cudaStream_t stream_;
    int slot_;
};
int64_t launch_count();
int profile_start(int max_launches);
int profile_stop(int32_t* ids, float* ms, double* work, int cap);

// 2-D bf16 row-major tensor [rows, cols] (row stride ld elements) -> TMA descriptor with a
// [box_rows x 64] box and 128-byte swizzle.  Returns 0 / negative error code.
int make_tmap_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

int make_tmap_bf16_3d(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t frames, int64_t ld, int64_t frame_stride,
                      int box_rows);

// uint8 tensor [frames, rows, cols] (strides in bytes, multiples of 16) -> TMA descriptor with a [1, box_rows, box_cols] box,
// no swizzle, zero fill out of bounds
int make_tmap_u8_3d(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t frames, int64_t row_pitch,
                    int64_t frame_pitch, int box_cols, int box_rows);
void set_preprocess_tma(int on);
void set_preprocess_identity(int on);

struct GemmParams;
// epi is a cre::GemmEpi value; cg = 1 | 2
int launch_gemm(int epi, int cg, const void* a, int64_t lda, const void* b, int64_t ldb, const GemmParams& p,
                int num_sms, cudaStream_t stream);

struct AttnArgs {
    const void* qkv;  // bf16 [n*t, ld]: q at column head*64, k at k_col0 + head*64, v at v_col0 + head*64
    int ld;
    int k_col0, v_col0;
    int n, t, heads;
    void* out;        // bf16 [n*t, heads*64]
    int* any_flag = nullptr;     // zeroed by the caller: set to 1 when some (frame, head) unit overflowed the fixed stabiliser ...
    int* unit_flags = nullptr;   // ... and [n * heads] flags naming the units (attention.cu "Exactness"); all zero again on return
};
inline int64_t attention_flag_ints(int n, int heads) { return 64 + static_cast<int64_t>(n) * heads; }   // [any (64-int slot) | units]
int launch_attention(const AttnArgs& a, cudaStream_t stream);

int launch_layernorm_bf16(const float* x, const float* g, const float* b, int rows, int dim, float eps,
                          __nv_bfloat16* out, cudaStream_t stream);
// LayerNorm folding: seed statistics + centred bf16 copy of x; weight folding at context creation (elementwise.cu)
int launch_row_stats(const float* x, int rows, int dim, int stride, __nv_bfloat16* xb, __nv_bfloat16* xl, float* stats,
                     cudaStream_t stream);
int launch_fold_ln_weights(const __nv_bfloat16* w, const float* gamma, const float* beta, const float* bias, int n_rows, int k,
                           int scaled_rows, float row_scale, __nv_bfloat16* wf, float* c1, float* c2, cudaStream_t stream);
// final LayerNorm + mean over the t tokens of each frame; tokens_out optional
// x == NULL: the split residual stream (xh + xl + the pivot in slot 0 of every statistics row)
int launch_final_norm_mean(const float* x, const __nv_bfloat16* xh, const __nv_bfloat16* xl, const float* stats, int stats_stride,
                           const float* g, const float* b, int frames, int t, int dim, float eps, float* frame_emb, float* tokens_out,
                           cudaStream_t stream);
int launch_fill_prefix(float* x, const float* prefix, int frames, int t, int prefix_tokens, int dim,
                       cudaStream_t stream);
int launch_pool_clips(const float* frame_emb, const int32_t* offs, int clips, int dim, float* out_mean,
                      float* out_unit, cudaStream_t stream);
int launch_split_hi_lo(const float* q, int rows, int dim, __nv_bfloat16* out, cudaStream_t stream);
int launch_fill_topk(float* scores, int32_t* idx, int64_t count, cudaStream_t stream);
int launch_gallery_scan_small(const float* queries, int q, int dim, const void* gallery, int rows, int row_base, int k, float* part_s,
                              int32_t* part_i, int slots, float* dump, int* done_counter, float* out_scores, int32_t* out_idx,
                              int out_stride, const float* cut_scores, const int32_t* cut_idx, cudaStream_t stream);
int launch_topk_prepare(const float* q, int rows, int dim, __nv_bfloat16* out, float* scores, int32_t* idx, int64_t count,
                        cudaStream_t stream);
int launch_merge_topk(const float* scores, const int32_t* idx, int64_t list_stride, int64_t query_stride,
                      int lists, int per_list, int q, int k, float* out_scores, int32_t* out_idx, int out_stride,
                      cudaStream_t stream);
int launch_gallery_update_row(__nv_bfloat16* gallery, float* master, int dim, int row, const float* unit_q, float momentum,
                              cudaStream_t stream);

struct ResizeTable {   // device-resident separable antialias weights for one axis
    const int32_t* lo;   // [out] first contributing input index
    const int32_t* cnt;  // [out] number of taps
    const float* w;      // [out, kmax] normalised weights (zero padded)
    int kmax;
    int in, out;
};
struct PreprocArgs {
    const uint8_t* frames;
    int n, h, w;
    int64_t row_pitch, frame_pitch;
    int bgr;
    int gh, gw;          // patch grid written (top-left gh*16 x gw*16 pixels of the resized image)
    float mean[3], inv_std[3];
    __nv_bfloat16* out;
    ResizeTable ty, tx;
    const int32_t* rois = nullptr;   // [n_rois, 5] = {frame, x0, y0, x1, y1}; ty / tx then hold one table block per ROI
    int n_rois = 0;
};
int launch_build_roi_tables(const int32_t* rois, int n_rois, int oh, int ow, int ykmax, int xkmax, int32_t* ylo, int32_t* ycnt,
                            float* yw, int32_t* xlo, int32_t* xcnt, float* xw, cudaStream_t stream);
int launch_preprocess(const PreprocArgs& a, cudaStream_t stream);

}  // namespace cre
