// Launcher + TMA descriptor construction for the tcgen05 GEMM family (gemm_tcgen05.cuh).
#include <stdarg.h>
#include <stdio.h>

#include <mutex>

#include "gemm_tcgen05.cuh"
#include "internal.h"

namespace cre {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

// cuTensorMapEncodeTiled is fetched through the runtime so the library has no link-time dependency
// on libcuda.so (it must dlopen on a GPU-less build box for the symbol-export test).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    CRE_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    CRE_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    CRE_REQUIRE((ld * 2) % 16 == 0, "TMA row stride must be a multiple of 16 bytes (ld=%lld)", (long long)ld);
    CRE_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range: %d", box_rows);
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CRE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)",
                (int)r, (long long)rows, (long long)cols, (long long)ld);
    return 0;
}

// bf16 tensor [frames][rows][cols] (row stride ld elements, frame stride in elements) -> [1, box_rows, 64] boxes, 128B swizzle:
// boxes are clipped at the end of a frame's rows (stores) / zero filled (loads)
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t frames, int64_t ld, int64_t frame_stride,
                      int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    CRE_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    CRE_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0 && (frame_stride * 2) % 16 == 0,
                "TMA 3-D tensor needs a 16-byte aligned base and strides");
    CRE_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range: %d", box_rows);
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(frames)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(frame_stride) * 2};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CRE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D bf16) failed with CUresult %d (cols=%lld rows=%lld frames=%lld)", (int)r,
                (long long)cols, (long long)rows, (long long)frames);
    return 0;
}

int make_tmap_u8_3d(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t frames, int64_t row_pitch,
                    int64_t frame_pitch, int box_cols, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    CRE_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    CRE_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && row_pitch % 16 == 0 && frame_pitch % 16 == 0,
                "TMA frame tensor needs 16-byte aligned base and pitches");
    CRE_REQUIRE(box_cols >= 16 && box_cols <= 256 && box_cols % 16 == 0 && box_rows >= 1 && box_rows <= 256, "TMA box %dx%d out of range",
                box_rows, box_cols);
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(frames)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(row_pitch), static_cast<cuuint64_t>(frame_pitch)};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CRE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (frames) failed with CUresult %d (cols=%lld rows=%lld frames=%lld pitch=%lld/%lld)",
                (int)r, (long long)cols, (long long)rows, (long long)frames, (long long)row_pitch, (long long)frame_pitch);
    return 0;
}

// output tile map for the TMA-store epilogues: box = 32 rows x 128 bytes (64 bf16 / 32 fp32 columns), 128B swizzle
static int make_tmap_out(CUtensorMap* out, bool bf16, void* base, int ldo, const GemmParams& p) {
    EncodeTiledFn fn = get_encode_fn();
    CRE_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    const int esize = bf16 ? 2 : 4;
    CRE_REQUIRE(base != nullptr && (reinterpret_cast<uintptr_t>(base) & 15) == 0, "gemm: output pointer must be non-NULL and 16-byte aligned");
    CRE_REQUIRE((static_cast<int64_t>(ldo) * esize) % 16 == 0, "gemm: output row stride must be a multiple of 16 bytes");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(p.N), static_cast<cuuint64_t>(p.M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldo) * esize};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esize), 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CRE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (output) failed with CUresult %d (M=%d N=%d ldo=%d)", (int)r, p.M, p.N,
                p.ldo);
    return 0;
}

// EPI_PATCH: the token matrix x as [frame][token][col] fp32 -- 32-row x 32-column boxes, clipped at a frame's last token
static int make_tmap_tokens(CUtensorMap* out, float* base, int ldo, const GemmParams& p) {
    EncodeTiledFn fn = get_encode_fn();
    CRE_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    CRE_REQUIRE(base != nullptr && (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (static_cast<int64_t>(ldo) * 4) % 16 == 0,
                "gemm: token matrix must be 16-byte aligned");
    CRE_REQUIRE(p.patches_per_frame >= 32 && p.tokens_per_frame == p.patches_per_frame + p.prefix_tokens && p.M % p.patches_per_frame == 0,
                "gemm: PATCH epilogue needs >= 32 patches per frame and M = frames * patches (M=%d, P=%d)", p.M, p.patches_per_frame);
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(p.N), static_cast<cuuint64_t>(p.tokens_per_frame),
                          static_cast<cuuint64_t>(p.M / p.patches_per_frame)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(ldo) * 4, static_cast<cuuint64_t>(p.tokens_per_frame) * ldo * 4};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CRE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (token matrix) failed with CUresult %d", (int)r);
    return 0;
}

template <int EPI, int CG, int STAGES = default_stages_epi(EPI, CG)>
static int launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                      cudaStream_t stream) {
    using Cfg = GemmCfg<EPI, CG, STAGES>;
    CUtensorMap tout = ta, tout2 = ta;   // placeholders for the epilogues that store directly
    if constexpr (EPI == EPI_PATCH) {
        const int rc = make_tmap_tokens(&tout, p.out_f32, p.ldo, p);
        if (rc) return rc;
    } else if constexpr (epi_resid_sp(EPI)) {   // both halves of the split residual stream: loaded and stored through the same maps
        int rc = make_tmap_out(&tout, true, p.out_bf16, p.ldo2, p);
        if (rc) return rc;
        rc = make_tmap_out(&tout2, true, p.x_lo, p.ldo2, p);
        if (rc) return rc;
    } else if constexpr (epi_tma_store(EPI)) {
        const bool bf16 = epi_out_bf16(EPI);
        const int rc = make_tmap_out(&tout, bf16, bf16 ? static_cast<void*>(p.out_bf16) : static_cast<void*>(p.out_f32), p.ldo, p);
        if (rc) return rc;
    }
    if constexpr (epi_resid_ln(EPI)) {
        const int rc = make_tmap_out(&tout2, true, p.out_bf16, p.ldo2, p);
        if (rc) return rc;
    }
    static_assert(Cfg::kSmemBytes <= 227 * 1024, "pipeline does not fit in shared memory");
    auto* kern = gemm_tn_kernel<EPI, CG, STAGES>;
    CRE_SMEM_ATTR_ONCE(kern, Cfg::kSmemBytes);   // per instantiation and device
    const int rows_per_tile = kBlockM * CG;
    const int64_t tiles = static_cast<int64_t>((p.M + rows_per_tile - 1) / rows_per_tile) *
                          ((p.N + kBlockN - 1) / kBlockN);
    int workers = num_sms / CG;
    if (tiles < workers) workers = static_cast<int>(tiles);
    if (workers < 1) workers = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(workers * CG);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int kid = (EPI == EPI_RESID || epi_resid_x(EPI)) && p.K > p.N ? CRE_K_GEMM_RESID_MLP : EPI == EPI_PATCH ? CRE_K_GEMM_PATCH : EPI == EPI_QKV ? CRE_K_GEMM_QKV : (EPI == EPI_RESID || epi_resid_x(EPI)) ? CRE_K_GEMM_RESID
                      : EPI == EPI_GELU ? CRE_K_GEMM_GELU : EPI == EPI_TOPK ? CRE_K_GEMM_TOPK : CRE_K_GEMM_PLAIN;
    // work: FLOPs, except the gallery scan which is bound by reading the bf16 gallery once (bytes)
    const double work = EPI == EPI_TOPK ? 2.0 * p.N * p.b_k_extent : 2.0 * p.M * static_cast<double>(p.N) * p.K;
    LaunchScope scope(kid, work, stream);
    CRE_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, tout, tout2, p));
    return 0;
}

int gemm_workers(int m, int n, int cg, int num_sms) {
    const int rows_per_tile = kBlockM * cg;
    const int64_t tiles = static_cast<int64_t>((m + rows_per_tile - 1) / rows_per_tile) * ((n + kBlockN - 1) / kBlockN);
    int workers = num_sms / cg;
    if (tiles < workers) workers = static_cast<int>(tiles);
    return workers < 1 ? 1 : workers;
}

#ifdef CRE_TUNING
static int g_debug_mode = 0;
void set_gemm_debug(int mode) { g_debug_mode = mode; }
#endif
static int g_tune_stages = 0;  // 0 = default_stages(cg); otherwise a tuning override for the plain epilogues
void set_gemm_stages(int stages) { g_tune_stages = stages; }

// non-default pipeline depths exist only for the epilogues cre_gemm_bf16 exposes (tuning harness)
template <int EPI>
static int launch_tuned(int cg, int stages, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                        cudaStream_t stream) {
    if (cg == 1) {
        if (stages == 3) return launch_one<EPI, 1, 3>(ta, tb, p, num_sms, stream);
        if (stages == 4) return launch_one<EPI, 1, 4>(ta, tb, p, num_sms, stream);
    } else {
        if (stages == 3) return launch_one<EPI, 2, 3>(ta, tb, p, num_sms, stream);
        if (stages == 4) return launch_one<EPI, 2, 4>(ta, tb, p, num_sms, stream);
        if (stages == 5) return launch_one<EPI, 2, 5>(ta, tb, p, num_sms, stream);
        if (stages == 6) return launch_one<EPI, 2, 6>(ta, tb, p, num_sms, stream);
    }
    set_error("gemm: no instantiation for cta_group=%d stages=%d", cg, stages);
    return -3;
}

int launch_gemm(int epi, int cg, const void* a, int64_t lda, const void* b, int64_t ldb, const GemmParams& p,
                int num_sms, cudaStream_t stream) {
    CRE_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem %dx%dx%d", p.M, p.N, p.K);
    CRE_REQUIRE(p.K % kBlockK == 0, "gemm: K=%d must be a multiple of %d", p.K, kBlockK);
    CRE_REQUIRE(p.b_k_extent % kBlockK == 0 && p.b_k_extent > 0, "gemm: bad b_k_extent %d", p.b_k_extent);
    CRE_REQUIRE(epi == EPI_TOPK || p.N % (epi_out_bf16(epi) ? 64 : 32) == 0, "gemm: N=%d must be a multiple of %d for this epilogue",
                p.N, epi_out_bf16(epi) ? 64 : 32);
    CRE_REQUIRE(cg == 1 || cg == 2, "gemm: cta_group must be 1 or 2");
    CRE_REQUIRE(epi != EPI_TOPK || p.K == 2 * p.b_k_extent, "gemm: TOPK expects A = [hi | lo] with K = 2 * b_k_extent (K=%d, extent=%d)", p.K,
                p.b_k_extent);
    if (epi_resid_sp(epi)) CRE_REQUIRE(p.out_bf16 != nullptr && p.x_lo != nullptr, "gemm: RESID_SP needs both halves of the residual stream");
    if (epi_resid_x(epi)) {
        CRE_REQUIRE(p.N % kBlockN == 0 && p.N / 128 == p.ln_slots && p.ln_slots <= 8 && p.ln_stride >= 2 * p.ln_slots + kLnStatsPad,
                    "gemm: RESID_LN needs N %% 256 == 0 and N / 128 = ln_slots <= 8 (N=%d slots=%d stride=%d)", p.N, p.ln_slots, p.ln_stride);
        CRE_REQUIRE(p.ln_stats_in != nullptr && p.ln_stats_out != nullptr && p.bias != nullptr && p.scale != nullptr,
                    "gemm: RESID_LN needs statistics in/out, bias and scale");
    }
    if (p.ln_stats_in != nullptr && !epi_resid_x(epi))
        CRE_REQUIRE(p.c1 != nullptr && p.bias != nullptr && p.ln_slots >= 1 && p.ln_slots <= 8, "gemm: folded LayerNorm needs c1, bias and 1..8 slots");
#ifdef CRE_TUNING
    if (g_debug_mode != 0 && (epi == EPI_NONE || g_debug_mode >= 4)) {
        GemmParams q = p;
        q.debug_mode = g_debug_mode;
        const int saved = g_debug_mode;
        g_debug_mode = 0;
        const int r = launch_gemm(epi, cg, a, lda, b, ldb, q, num_sms, stream);
        g_debug_mode = saved;
        return r;
    }
#endif
    if (epi == EPI_QKV && p.c1 == nullptr) {   // unfolded QKV: c1 is multiplied by 0
        GemmParams q = p;
        q.c1 = p.bias;
        return launch_gemm(epi, cg, a, lda, b, ldb, q, num_sms, stream);
    }
    CUtensorMap ta, tb;
    CRE_REQUIRE(!p.topk_stacked || (epi == EPI_TOPK && p.M <= 64), "gemm: the stacked hi / lo form needs EPI_TOPK and M <= 64 (M=%d)", p.M);
    int rc = make_tmap_bf16(&ta, a, p.M, p.K, lda, p.topk_stacked ? 64 : kBlockM);   // stacked: two 64-row boxes fill one A tile
    if (rc) return rc;
    rc = make_tmap_bf16(&tb, b, p.N, p.b_k_extent, ldb, kBlockN / cg);
    if (rc) return rc;
    if (g_tune_stages != 0 && g_tune_stages != default_stages(cg)) {
        switch (epi) {
            case EPI_BF16: return launch_tuned<EPI_BF16>(cg, g_tune_stages, ta, tb, p, num_sms, stream);
            case EPI_GELU: return launch_tuned<EPI_GELU>(cg, g_tune_stages, ta, tb, p, num_sms, stream);
            case EPI_RESID: return launch_tuned<EPI_RESID>(cg, g_tune_stages, ta, tb, p, num_sms, stream);
            case EPI_NONE: return launch_tuned<EPI_NONE>(cg, g_tune_stages, ta, tb, p, num_sms, stream);
            default: break;
        }
    }
#define CRE_CASE(E)                                                        \
    case E:                                                                \
        return cg == 1 ? launch_one<E, 1>(ta, tb, p, num_sms, stream)      \
                       : launch_one<E, 2>(ta, tb, p, num_sms, stream);
    switch (epi) {
        CRE_CASE(EPI_BF16)
        CRE_CASE(EPI_F32)
        CRE_CASE(EPI_QKV)
        CRE_CASE(EPI_GELU)
        CRE_CASE(EPI_RESID)
        CRE_CASE(EPI_RESID_LN)
        CRE_CASE(EPI_RESID_LN3)
        CRE_CASE(EPI_RESID_SP)
        CRE_CASE(EPI_RESID_SP3)
        CRE_CASE(EPI_PATCH)
        CRE_CASE(EPI_NONE)
        case EPI_TOPK:
            return launch_one<EPI_TOPK, 1>(ta, tb, p, num_sms, stream);
        default:
            set_error("gemm: unknown epilogue %d", epi);
            return -1;
    }
#undef CRE_CASE
}

}  // namespace cre
