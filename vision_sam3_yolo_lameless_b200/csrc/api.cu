// extern "C" surface of libcre_b200.so (include/cre.h): context, packed-weight layout, workspace
// carving and the launch sequence of the ViT forward pass.
#include <math.h>
#include <string.h>

#include <map>
#include <new>
#include <utility>
#include <vector>

#include "gemm_tcgen05.cuh"
#include "internal.h"

using namespace cre;

namespace {

constexpr int64_t kAlign = 256;
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct DevTable {
    int32_t* lo = nullptr;
    int32_t* cnt = nullptr;
    float* w = nullptr;
    int kmax = 0, in = 0, out = 0;
};
// kRopeRowFloats (gemm_tcgen05.cuh): cos[16] | sin[16] | -sin[16] | pad per axis position
struct RopeTable {
    float* axis = nullptr;   // [(gh + gw), kRopeRowFloats]: rows 0..gh-1 = y positions, rows gh.. = x positions
};

bool cfg_ok(const cre_model_cfg* c) {
    if (c == nullptr) {
        set_error("cfg is NULL");
        return false;
    }
    if (c->hidden <= 0 || c->hidden % 64 != 0 || c->heads <= 0 || c->hidden != c->heads * 64) {
        set_error("cfg: hidden=%d heads=%d (need hidden == 64*heads)", c->hidden, c->heads);
        return false;
    }
    if (c->hidden != 768 && c->hidden != 1024) {
        set_error("cfg: hidden=%d unsupported (768 | 1024)", c->hidden);
        return false;
    }
    if (c->layers <= 0 || c->mlp <= 0 || c->mlp % 64 != 0 || c->patch != 16 || c->registers < 0) {
        set_error("cfg: layers=%d mlp=%d patch=%d registers=%d unsupported", c->layers, c->mlp, c->patch, c->registers);
        return false;
    }
    return true;
}

// elements and element size of one tensor kind
void kind_shape(const cre_model_cfg* c, int kind, int64_t* elems, int* esize) {
    const int64_t D = c->hidden, F = c->mlp, PK = 3LL * c->patch * c->patch;
    switch (kind) {
        case CRE_W_PATCH: *elems = D * PK; *esize = 2; break;
        case CRE_B_PATCH: *elems = D; *esize = 4; break;
        case CRE_PREFIX: *elems = (1 + c->registers) * D; *esize = 4; break;
        case CRE_LN_F_G: case CRE_LN_F_B: case CRE_LN1_G: case CRE_LN1_B: case CRE_LN2_G: case CRE_LN2_B:
        case CRE_B_O: case CRE_LS1: case CRE_LS2: case CRE_B_DOWN: *elems = D; *esize = 4; break;
        case CRE_W_QKV: *elems = 3 * D * D; *esize = 2; break;
        case CRE_B_QKV: *elems = 3 * D; *esize = 4; break;
        case CRE_W_O: *elems = D * D; *esize = 2; break;
        case CRE_W_UP: *elems = F * D; *esize = 2; break;
        case CRE_B_UP: *elems = F; *esize = 4; break;
        case CRE_W_DOWN: *elems = D * F; *esize = 2; break;
        default: *elems = 0; *esize = 0; break;
    }
}
inline bool kind_is_global(int kind) { return kind <= CRE_LN_F_B; }

int64_t globals_bytes(const cre_model_cfg* c) {
    int64_t off = 0;
    for (int k = 0; k <= CRE_LN_F_B; ++k) {
        int64_t e; int s;
        kind_shape(c, k, &e, &s);
        off += align_up(e * s, kAlign);
    }
    return off;
}
int64_t layer_bytes(const cre_model_cfg* c) {
    int64_t off = 0;
    for (int k = CRE_LN1_G; k < CRE_WEIGHT_KINDS; ++k) {
        int64_t e; int s;
        kind_shape(c, k, &e, &s);
        off += align_up(e * s, kAlign);
    }
    return off;
}
int64_t weight_offset(const cre_model_cfg* c, int layer, int kind) {
    if (kind < 0 || kind >= CRE_WEIGHT_KINDS) return -1;
    int64_t off = 0;
    int first = 0;
    if (kind_is_global(kind)) {
        if (layer != -1) return -1;
    } else {
        if (layer < 0 || layer >= c->layers) return -1;
        off = globals_bytes(c) + layer * layer_bytes(c);
        first = CRE_LN1_G;
    }
    for (int k = first; k < kind; ++k) {
        int64_t e; int s;
        kind_shape(c, k, &e, &s);
        off += align_up(e * s, kAlign);
    }
    return off;
}

struct Workspace {
    float* x;            // [M, D] residual stream, fp32
    __nv_bfloat16* h;    // [M, D] LayerNorm output / attention output
    __nv_bfloat16* qkv;  // [M, 3D] q (pre-scaled, rotated) | k (rotated) | v
    __nv_bfloat16* mlp;  // [M, F]
    __nv_bfloat16* xb;   // [M, D] bf16(x - row pivot): A operand of the LayerNorm-folded GEMMs = HIGH half of the split residual stream
    __nv_bfloat16* xl;   // [M, D] bf16(x - row pivot - xb): its LOW half (resid_split = 1: x itself is only the patch embedding's output)
    float* stats[2];     // [M, ln_stride] LayerNorm statistics rows (ping-pong between the two norms of a block)
    int* attn_flags;     // [layers] x [64-int "any" slot] then [frames * heads] unit flags: attention overflow tracking
    int64_t attn_flag_bytes;
    int64_t total;
};
inline int ln_slots(const cre_model_cfg* c) { return c->hidden / 128; }
inline int ln_stride(const cre_model_cfg* c) { return 2 * ln_slots(c) + 4; }
Workspace carve(const cre_model_cfg* c, int frames, int gh, int gw, void* base) {
    const int64_t T = static_cast<int64_t>(gh) * gw + 1 + c->registers;
    const int64_t M = frames * T, D = c->hidden, F = c->mlp;
    uint8_t* p = static_cast<uint8_t*>(base);
    int64_t off = 0;
    Workspace w;
    auto take = [&](int64_t bytes) { uint8_t* r = p + off; off += align_up(bytes, 1024); return r; };
    w.x = reinterpret_cast<float*>(take(M * D * 4));
    w.h = reinterpret_cast<__nv_bfloat16*>(take(M * D * 2));
    w.qkv = reinterpret_cast<__nv_bfloat16*>(take(M * 3 * D * 2));
    w.mlp = reinterpret_cast<__nv_bfloat16*>(take(M * F * 2));
    w.xb = reinterpret_cast<__nv_bfloat16*>(take(M * D * 2));
    w.xl = reinterpret_cast<__nv_bfloat16*>(take(M * D * 2));
    w.stats[0] = reinterpret_cast<float*>(take(M * ln_stride(c) * 4));
    w.stats[1] = reinterpret_cast<float*>(take(M * ln_stride(c) * 4));
    w.attn_flag_bytes = (64LL * c->layers + static_cast<int64_t>(frames) * c->heads) * 4;
    w.attn_flags = reinterpret_cast<int*>(take(w.attn_flag_bytes));
    w.total = off;
    return w;
}

// aten _upsample_bilinear2d_aa weights for one axis (ATen/native/cpu/UpSampleKernel.cpp,
// HelperInterpBase::_compute_indices_min_size_weights_aa with the bilinear (triangle) filter), float math.
void build_resize_table(int in, int out, std::vector<int32_t>& lo, std::vector<int32_t>& cnt, std::vector<float>& w,
                        int* kmax_out) {
    const float scale = static_cast<float>(in) / static_cast<float>(out);
    const float support = scale >= 1.0f ? scale : 1.0f;
    const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
    const int kmax = static_cast<int>(ceilf(support)) * 2 + 1;
    lo.assign(out, 0);
    cnt.assign(out, 0);
    w.assign(static_cast<size_t>(out) * kmax, 0.0f);
    for (int i = 0; i < out; ++i) {
        const float center = scale * (static_cast<float>(i) + 0.5f);
        int xmin = static_cast<int>(center - support + 0.5f);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5f);
        if (xmax > in) xmax = in;
        const int xsize = xmax - xmin;
        float total = 0.0f;
        float* wi = w.data() + static_cast<size_t>(i) * kmax;
        for (int j = 0; j < xsize && j < kmax; ++j) {
            float x = (static_cast<float>(j + xmin) - center + 0.5f) * invscale;
            x = fabsf(x);
            const float v = x < 1.0f ? 1.0f - x : 0.0f;
            wi[j] = v;
            total += v;
        }
        if (total != 0.0f)
            for (int j = 0; j < xsize && j < kmax; ++j) wi[j] /= total;
        lo[i] = xmin;
        cnt[i] = xsize < kmax ? xsize : kmax;
    }
    *kmax_out = kmax;
}

}  // namespace

// Linear weights with the preceding LayerNorm folded in (elementwise.cu fold_ln_weights_kernel), built at cre_create
struct FoldedLinear {
    __nv_bfloat16* w = nullptr;   // [out, in] = bf16(W * gamma)
    float* c1 = nullptr;          // [out] row sums of w
    float* c2 = nullptr;          // [out] bias + W beta
};
struct cre_ctx {
    cre_model_cfg cfg;
    const uint8_t* weights;
    int device;
    int num_sms;
    uint8_t* fold_buf = nullptr;
    std::vector<FoldedLinear> fold_qkv, fold_up;
    std::map<std::pair<int, int>, DevTable> resize_tables;
    std::map<std::pair<int, int>, RopeTable> rope_tables;

    template <typename T>
    const T* w(int layer, int kind) const {
        return reinterpret_cast<const T*>(weights + weight_offset(&cfg, layer, kind));
    }
};

namespace {

// First use of an (input size, output size) pair builds its antialias table on the host and uploads it ON THE CALLER'S STREAM
// (pageable source: the driver stages it before cudaMemcpyAsync returns, and the copy is ordered before the kernels that follow on
// that stream -- also on cudaStreamNonBlocking streams).  That first call allocates (cudaMalloc: not capturable in a CUDA graph --
// warm the sizes up before capturing).  The cache is bounded: beyond kMaxCachedTables distinct sizes it is flushed after a stream
// synchronisation (a service sees a handful of camera resolutions; ROI crops do not go through this cache at all).
constexpr size_t kMaxCachedTables = 64;
int get_resize_table(cre_ctx* ctx, int in, int out, DevTable* t, cudaStream_t stream) {
    auto key = std::make_pair(in, out);
    auto it = ctx->resize_tables.find(key);
    if (it != ctx->resize_tables.end()) {
        *t = it->second;
        return 0;
    }
    if (ctx->resize_tables.size() >= kMaxCachedTables) {
        CRE_CUDA_OK(cudaStreamSynchronize(stream));       // no queued kernel may still read a table that is about to be freed
        for (auto& kv : ctx->resize_tables) {
            cudaFree(kv.second.lo);
            cudaFree(kv.second.cnt);
            cudaFree(kv.second.w);
        }
        ctx->resize_tables.clear();
    }
    std::vector<int32_t> lo, cnt;
    std::vector<float> w;
    int kmax = 0;
    build_resize_table(in, out, lo, cnt, w, &kmax);
    DevTable d;
    d.kmax = kmax;
    d.in = in;
    d.out = out;
    CRE_CUDA_OK(cudaMalloc(&d.lo, lo.size() * 4));
    CRE_CUDA_OK(cudaMalloc(&d.cnt, cnt.size() * 4));
    CRE_CUDA_OK(cudaMalloc(&d.w, w.size() * 4));
    CRE_CUDA_OK(cudaMemcpyAsync(d.lo, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice, stream));
    CRE_CUDA_OK(cudaMemcpyAsync(d.cnt, cnt.data(), cnt.size() * 4, cudaMemcpyHostToDevice, stream));
    CRE_CUDA_OK(cudaMemcpyAsync(d.w, w.data(), w.size() * 4, cudaMemcpyHostToDevice, stream));
    ctx->resize_tables[key] = d;
    *t = d;
    return 0;
}

// HF:modeling_dinov3_vit.py:95-121,153-200: patch-centre coordinates in [-1, 1], 16 inverse frequencies,
// angles = 2*pi*coord*inv_freq laid out [y*f0..f15, x*f0..f15] and tiled twice; fp32 cos / sin.
int get_rope_table(cre_ctx* ctx, int gh, int gw, RopeTable* t, cudaStream_t stream) {
    auto key = std::make_pair(gh, gw);
    auto it = ctx->rope_tables.find(key);
    if (it != ctx->rope_tables.end()) {
        *t = it->second;
        return 0;
    }
    CRE_REQUIRE((gh + gw) * kRopeRowFloats * 4 <= kRopeTableBytes, "rotary table for a %dx%d patch grid exceeds the kernel's %d bytes",
                gh, gw, kRopeTableBytes);
    // The HF table is cos/sin of [y*f0..f15, x*f0..f15] tiled twice: per token only the 16 y-angles of its patch row and
    // the 16 x-angles of its patch column are distinct, so one small per-axis table serves every token.
    std::vector<float> tab(static_cast<size_t>(gh + gw) * kRopeRowFloats, 0.0f);
    float inv_freq[16];
    for (int k = 0; k < 16; ++k) inv_freq[k] = 1.0f / powf(ctx->cfg.rope_theta, static_cast<float>(k) * (4.0f / 64.0f));
    const float two_pi = static_cast<float>(2.0 * M_PI);
    for (int a = 0; a < gh + gw; ++a) {
        const int pos = a < gh ? a : a - gh, n = a < gh ? gh : gw;
        const float coord = 2.0f * ((static_cast<float>(pos) + 0.5f) / static_cast<float>(n)) - 1.0f;
        float* row = tab.data() + static_cast<size_t>(a) * kRopeRowFloats;
        for (int k = 0; k < 16; ++k) {
            const float ang = two_pi * coord * inv_freq[k];
            row[k] = cosf(ang);
            row[16 + k] = sinf(ang);
            row[32 + k] = -row[16 + k];
        }
    }
    RopeTable r;
    CRE_CUDA_OK(cudaMalloc(&r.axis, tab.size() * 4));
    CRE_CUDA_OK(cudaMemcpyAsync(r.axis, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, stream));   // caller's stream, see get_resize_table
    ctx->rope_tables[key] = r;
    *t = r;
    return 0;
}

GemmParams base_params(int M, int N, int K) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M;
    p.N = N;
    p.K = K;
    p.b_k_extent = K;
    p.ldo = N;
    return p;
}

int g_default_cg = 2;  // CTA pairs: less smem traffic per FLOP, measured 3-5 % faster in the full step
constexpr float kQScale = 0.125f;   // head_dim^-0.5, head_dim = 64 (HF:modeling_dinov3_vit.py:284); folded into the q rows
int g_resid_ln_deep = 2;   // bit 0: attention-out projection, bit 1: MLP down projection use EPI_RESID_LN3 (1 x tile, 5 stages)
int g_resid_split = 1; // 1 (with ln_fold): the residual stream lives in HBM as two bf16 halves around the row pivot (EPI_RESID_SP);
                       // 0: fp32 stream + a bf16 copy (EPI_RESID_LN)
int g_ln_fold = 1;     // 1: LayerNorm folded into the GEMMs (no LayerNorm kernel in the blocks); 0: separate LayerNorm launches

// LN1 -> q/k/v and LN2 -> up projections of every layer, folded once
int build_folded_weights(cre_ctx* c) {
    const cre_model_cfg& m = c->cfg;
    const int64_t D = m.hidden, F = m.mlp;
    auto bytes_of = [&](int64_t out) { return align_up(out * D * 2, kAlign) + 2 * align_up(out * 4, kAlign); };
    const int64_t per_layer = bytes_of(3 * D) + bytes_of(F);
    CRE_CUDA_OK(cudaDeviceSynchronize());   // the caller's upload of the packed blob (any stream) is complete
    CRE_CUDA_OK(cudaMalloc(&c->fold_buf, per_layer * m.layers));
    c->fold_qkv.resize(m.layers);
    c->fold_up.resize(m.layers);
    uint8_t* p = c->fold_buf;
    auto take = [&](FoldedLinear& f, int64_t out) {
        f.w = reinterpret_cast<__nv_bfloat16*>(p); p += align_up(out * D * 2, kAlign);
        f.c1 = reinterpret_cast<float*>(p); p += align_up(out * 4, kAlign);
        f.c2 = reinterpret_cast<float*>(p); p += align_up(out * 4, kAlign);
    };
    for (int l = 0; l < m.layers; ++l) {
        take(c->fold_qkv[l], 3 * D);
        take(c->fold_up[l], F);
        int rc = launch_fold_ln_weights(c->w<__nv_bfloat16>(l, CRE_W_QKV), c->w<float>(l, CRE_LN1_G), c->w<float>(l, CRE_LN1_B),
                                        c->w<float>(l, CRE_B_QKV), static_cast<int>(3 * D), static_cast<int>(D), static_cast<int>(D), kQScale,
                                        c->fold_qkv[l].w, c->fold_qkv[l].c1, c->fold_qkv[l].c2, nullptr);
        if (rc) return rc;
        rc = launch_fold_ln_weights(c->w<__nv_bfloat16>(l, CRE_W_UP), c->w<float>(l, CRE_LN2_G), c->w<float>(l, CRE_LN2_B),
                                    c->w<float>(l, CRE_B_UP), static_cast<int>(F), static_cast<int>(D), 0, 1.0f, c->fold_up[l].w,
                                    c->fold_up[l].c1, c->fold_up[l].c2, nullptr);
        if (rc) return rc;
    }
    CRE_CUDA_OK(cudaStreamSynchronize(nullptr));
    return 0;
}

}  // namespace

extern "C" {

const char* cre_last_error(void) { return last_error(); }
int32_t cre_abi_version(void) { return CRE_ABI_VERSION; }

int64_t cre_packed_weights_bytes(const cre_model_cfg* cfg) {
    if (!cfg_ok(cfg)) return -1;
    return globals_bytes(cfg) + static_cast<int64_t>(cfg->layers) * layer_bytes(cfg);
}
int64_t cre_weight_offset(const cre_model_cfg* cfg, int32_t layer, int32_t kind) {
    if (!cfg_ok(cfg)) return -1;
    const int64_t off = weight_offset(cfg, layer, kind);
    if (off < 0) set_error("weight_offset: bad (layer=%d, kind=%d)", layer, kind);
    return off;
}
int64_t cre_weight_elems(const cre_model_cfg* cfg, int32_t layer, int32_t kind) {
    if (!cfg_ok(cfg)) return -1;
    if (weight_offset(cfg, layer, kind) < 0) {
        set_error("weight_elems: bad (layer=%d, kind=%d)", layer, kind);
        return -1;
    }
    int64_t e; int s;
    kind_shape(cfg, kind, &e, &s);
    return e;
}

int32_t cre_create(const cre_model_cfg* cfg, const void* packed_weights_dev, int32_t device, cre_ctx** out) {
    if (!cfg_ok(cfg)) return -1;
    CRE_REQUIRE(out != nullptr, "cre_create: out is NULL");
    CRE_REQUIRE((reinterpret_cast<uintptr_t>(packed_weights_dev) & 255) == 0, "cre_create: weights must be 256-byte aligned");
    cudaDeviceProp prop;
    CRE_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    CRE_REQUIRE(prop.major == 10, "cre_create: device %d is sm_%d%d; this library is sm_100a only (no fallback)", device,
                prop.major, prop.minor);
    CRE_CUDA_OK(cudaSetDevice(device));
    cre_ctx* c = new (std::nothrow) cre_ctx();
    CRE_REQUIRE(c != nullptr, "cre_create: out of host memory");
    c->cfg = *cfg;
    c->weights = static_cast<const uint8_t*>(packed_weights_dev);
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    int rc = build_folded_weights(c);   // the packed blob must already hold the weights (it does: see engine.py)
    if (rc) {
        cudaFree(c->fold_buf);
        delete c;
        return rc;
    }
    *out = c;
    return 0;
}

int32_t cre_destroy(cre_ctx* ctx) {
    if (ctx == nullptr) return 0;
    for (auto& kv : ctx->resize_tables) {
        cudaFree(kv.second.lo);
        cudaFree(kv.second.cnt);
        cudaFree(kv.second.w);
    }
    for (auto& kv : ctx->rope_tables) cudaFree(kv.second.axis);
    cudaFree(ctx->fold_buf);
    delete ctx;
    return 0;
}

int64_t cre_workspace_bytes(const cre_model_cfg* cfg, int32_t frames, int32_t grid_h, int32_t grid_w) {
    if (!cfg_ok(cfg)) return -1;
    if (frames <= 0 || grid_h <= 0 || grid_w <= 0) {
        set_error("workspace_bytes: frames=%d grid=%dx%d", frames, grid_h, grid_w);
        return -1;
    }
    return carve(cfg, frames, grid_h, grid_w, nullptr).total;
}

int32_t cre_preprocess_patchify(cre_ctx* ctx, const uint8_t* frames_dev, int32_t n, int32_t h, int32_t w,
                                int64_t row_pitch, int64_t frame_pitch, int32_t bgr, int32_t resize_h,
                                int32_t resize_w, const float mean[3], const float std_[3], void* out_patches_dev,
                                void* stream) {
    CRE_REQUIRE(ctx != nullptr && frames_dev != nullptr && out_patches_dev != nullptr, "preprocess: NULL argument");
    CRE_REQUIRE(n > 0 && h > 0 && w > 0 && resize_h >= 16 && resize_w >= 16, "preprocess: bad sizes n=%d %dx%d -> %dx%d", n, h,
                w, resize_h, resize_w);
    CRE_REQUIRE(ctx->cfg.patch == 16, "preprocess: patch size must be 16");
    PreprocArgs a;
    a.frames = frames_dev;
    a.n = n;
    a.h = h;
    a.w = w;
    a.row_pitch = row_pitch;
    a.frame_pitch = frame_pitch;
    a.bgr = bgr;
    a.gh = resize_h / 16;
    a.gw = resize_w / 16;
    for (int i = 0; i < 3; ++i) {
        a.mean[i] = mean[i];
        a.inv_std[i] = 1.0f / std_[i];
    }
    a.out = static_cast<__nv_bfloat16*>(out_patches_dev);
    DevTable ty, tx;
    int rc = get_resize_table(ctx, h, resize_h, &ty, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    rc = get_resize_table(ctx, w, resize_w, &tx, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    a.ty = {ty.lo, ty.cnt, ty.w, ty.kmax, ty.in, ty.out};
    a.tx = {tx.lo, tx.cnt, tx.w, tx.kmax, tx.in, tx.out};
    return launch_preprocess(a, static_cast<cudaStream_t>(stream));
}

namespace {
// widest tap count any crop of an `in`-pixel axis can need when resized to `out` (build_resize_table's kmax for the whole axis)
int roi_kmax(int in, int out) {
    const float scale = static_cast<float>(in) / static_cast<float>(out);
    return static_cast<int>(ceilf(scale >= 1.0f ? scale : 1.0f)) * 2 + 1;
}
struct RoiScratch {
    int32_t *ylo, *ycnt, *xlo, *xcnt;
    float *yw, *xw;
    int ykmax, xkmax;
    int64_t total;
};
RoiScratch carve_roi(int n_rois, int h, int w, int oh, int ow, void* base) {
    RoiScratch r;
    r.ykmax = roi_kmax(h, oh);
    r.xkmax = roi_kmax(w, ow);
    uint8_t* p = static_cast<uint8_t*>(base);
    int64_t off = 0;
    auto take = [&](int64_t bytes) { uint8_t* q = p + off; off += align_up(bytes, kAlign); return q; };
    r.ylo = reinterpret_cast<int32_t*>(take(4LL * n_rois * oh));
    r.ycnt = reinterpret_cast<int32_t*>(take(4LL * n_rois * oh));
    r.xlo = reinterpret_cast<int32_t*>(take(4LL * n_rois * ow));
    r.xcnt = reinterpret_cast<int32_t*>(take(4LL * n_rois * ow));
    r.yw = reinterpret_cast<float*>(take(4LL * n_rois * oh * r.ykmax));
    r.xw = reinterpret_cast<float*>(take(4LL * n_rois * ow * r.xkmax));
    r.total = off;
    return r;
}
}  // namespace

int64_t cre_roi_scratch_bytes(int32_t n_rois, int32_t h, int32_t w, int32_t resize_h, int32_t resize_w) {
    if (n_rois <= 0 || h <= 0 || w <= 0 || resize_h < 16 || resize_w < 16) {
        set_error("roi_scratch_bytes: n_rois=%d %dx%d -> %dx%d", n_rois, h, w, resize_h, resize_w);
        return -1;
    }
    return carve_roi(n_rois, h, w, resize_h / 16 * 16, resize_w / 16 * 16, nullptr).total;
}

int32_t cre_preprocess_patchify_roi(cre_ctx* ctx, const uint8_t* frames_dev, int32_t n_frames, int32_t h, int32_t w,
                                    int64_t row_pitch, int64_t frame_pitch, int32_t bgr, const int32_t* rois_dev, int32_t n_rois,
                                    int32_t resize_h, int32_t resize_w, const float mean[3], const float std_[3], void* scratch_dev,
                                    int64_t scratch_bytes, void* out_patches_dev, void* stream_) {
    CRE_REQUIRE(ctx != nullptr && frames_dev != nullptr && rois_dev != nullptr && scratch_dev != nullptr && out_patches_dev != nullptr,
                "preprocess_roi: NULL argument");
    CRE_REQUIRE(n_frames > 0 && n_rois > 0 && h > 0 && w > 0 && resize_h >= 16 && resize_w >= 16,
                "preprocess_roi: bad sizes frames=%d rois=%d %dx%d -> %dx%d", n_frames, n_rois, h, w, resize_h, resize_w);
    CRE_REQUIRE(ctx->cfg.patch == 16, "preprocess_roi: patch size must be 16");
    CRE_REQUIRE((reinterpret_cast<uintptr_t>(scratch_dev) & 255) == 0, "preprocess_roi: scratch must be 256-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int gh = resize_h / 16, gw = resize_w / 16;
    RoiScratch sc = carve_roi(n_rois, h, w, gh * 16, gw * 16, scratch_dev);
    CRE_REQUIRE(scratch_bytes >= sc.total, "preprocess_roi: scratch %lld < required %lld bytes", (long long)scratch_bytes, (long long)sc.total);
    int rc = launch_build_roi_tables(rois_dev, n_rois, gh * 16, gw * 16, sc.ykmax, sc.xkmax, sc.ylo, sc.ycnt, sc.yw, sc.xlo, sc.xcnt, sc.xw,
                                     stream);
    if (rc) return rc;
    PreprocArgs a;
    a.frames = frames_dev;
    a.n = n_frames;
    a.h = h;
    a.w = w;
    a.row_pitch = row_pitch;
    a.frame_pitch = frame_pitch;
    a.bgr = bgr;
    a.gh = gh;
    a.gw = gw;
    for (int i = 0; i < 3; ++i) {
        a.mean[i] = mean[i];
        a.inv_std[i] = 1.0f / std_[i];
    }
    a.out = static_cast<__nv_bfloat16*>(out_patches_dev);
    a.ty = {sc.ylo, sc.ycnt, sc.yw, sc.ykmax, h, gh * 16};      // in / out: the widest crop sizes the shared-memory band
    a.tx = {sc.xlo, sc.xcnt, sc.xw, sc.xkmax, w, gw * 16};
    a.rois = rois_dev;
    a.n_rois = n_rois;
    return launch_preprocess(a, stream);
}

int32_t cre_vit_forward(cre_ctx* ctx, const void* patches_dev, int32_t n, int32_t grid_h, int32_t grid_w,
                        void* workspace_dev, int64_t workspace_bytes, float* out_frame_emb_dev, float* out_tokens_dev,
                        void* stream_) {
    CRE_REQUIRE(ctx != nullptr && patches_dev != nullptr && workspace_dev != nullptr && out_frame_emb_dev != nullptr,
                "vit_forward: NULL argument");
    CRE_REQUIRE(n > 0 && grid_h > 0 && grid_w > 0, "vit_forward: bad sizes n=%d grid=%dx%d", n, grid_h, grid_w);
    CRE_REQUIRE((reinterpret_cast<uintptr_t>(workspace_dev) & 1023) == 0, "vit_forward: workspace must be 1024-byte aligned");
    const cre_model_cfg& c = ctx->cfg;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws = carve(&c, n, grid_h, grid_w, workspace_dev);
    CRE_REQUIRE(workspace_bytes >= ws.total, "vit_forward: workspace %lld < required %lld bytes", (long long)workspace_bytes,
                (long long)ws.total);
    const int P = grid_h * grid_w, prefix = 1 + c.registers, T = P + prefix;
    const int64_t M64 = static_cast<int64_t>(n) * T;
    CRE_REQUIRE(M64 < (1LL << 31) / 4, "vit_forward: too many token rows (%lld); split the batch", (long long)M64);
    const int M = static_cast<int>(M64), D = c.hidden, F = c.mlp, PK = 3 * c.patch * c.patch;
    const int cg = g_default_cg;
    RopeTable rope;
    int rc = get_rope_table(ctx, grid_h, grid_w, &rope, stream);
    if (rc) return rc;

    // attention overflow flags: one "any" slot per layer + the per-unit flags (which every layer leaves zeroed again)
    CRE_CUDA_OK(cudaMemsetAsync(ws.attn_flags, 0, static_cast<size_t>(ws.attn_flag_bytes), stream));
    {   // patch embedding: [n*P, 768] x [D, 768]^T + bias -> token rows prefix.. of x
        GemmParams p = base_params(n * P, D, PK);
        p.bias = ctx->w<float>(-1, CRE_B_PATCH);
        p.out_f32 = ws.x;
        p.ldo = D;
        p.tokens_per_frame = T;
        p.prefix_tokens = prefix;
        p.patches_per_frame = P;
        rc = launch_gemm(EPI_PATCH, cg, patches_dev, PK, ctx->w<void>(-1, CRE_W_PATCH), PK, p, ctx->num_sms, stream);
        if (rc) return rc;
    }
    rc = launch_fill_prefix(ws.x, ctx->w<float>(-1, CRE_PREFIX), n, T, prefix, D, stream);
    if (rc) return rc;
    const bool fold = g_ln_fold != 0;
    const bool split = fold && g_resid_split != 0;
    const int S = ln_slots(&c), SS = ln_stride(&c);
    if (fold) {   // seed of the folded chain: statistics + centred bf16 copy of the embedded tokens (split: + its low half)
        rc = launch_row_stats(ws.x, M, D, SS, ws.xb, split ? ws.xl : nullptr, ws.stats[0], stream);
        if (rc) return rc;
    }
    for (int l = 0; l < c.layers; ++l) {
        if (!fold) {
            rc = launch_layernorm_bf16(ws.x, ctx->w<float>(l, CRE_LN1_G), ctx->w<float>(l, CRE_LN1_B), M, D, c.ln_eps, ws.h, stream);
            if (rc) return rc;
        }
        {
            GemmParams p = base_params(M, 3 * D, D);
            p.bias = fold ? ctx->fold_qkv[l].c2 : ctx->w<float>(l, CRE_B_QKV);
            p.out_bf16 = ws.qkv;
            p.ldo = 3 * D;
            p.rope_axis = rope.axis;
            p.grid_h = grid_h;
            p.grid_w = grid_w;
            p.tokens_per_frame = T;
            p.prefix_tokens = prefix;
            p.hidden = D;
            p.q_scale = fold ? 1.0f : kQScale;   // the folded q rows are pre-scaled
            if (fold) {
                p.c1 = ctx->fold_qkv[l].c1;
                p.ln_stats_in = ws.stats[0];
                p.ln_slots = S;
                p.ln_stride = SS;
                p.ln_eps = c.ln_eps;
            }
            rc = launch_gemm(EPI_QKV, cg, fold ? ws.xb : ws.h, D, fold ? static_cast<const void*>(ctx->fold_qkv[l].w) : ctx->w<void>(l, CRE_W_QKV),
                             D, p, ctx->num_sms, stream);
            if (rc) return rc;
        }
        {
            AttnArgs a;
            a.qkv = ws.qkv;
            a.ld = 3 * D;
            a.k_col0 = D;
            a.v_col0 = 2 * D;
            a.n = n;
            a.t = T;
            a.heads = c.heads;
            a.out = ws.h;
            a.any_flag = ws.attn_flags + 64 * l;
            a.unit_flags = ws.attn_flags + 64 * c.layers;
            rc = launch_attention(a, stream);
            if (rc) return rc;
        }
        {   // attention output projection: x += ls1 * (h Wo^T + b); folded: also bf16(x - pivot) and the LN2 statistics
            GemmParams p = base_params(M, D, D);
            p.bias = ctx->w<float>(l, CRE_B_O);
            p.scale = ctx->w<float>(l, CRE_LS1);
            p.out_f32 = ws.x;
            p.ldo = D;
            if (fold) {
                p.out_bf16 = ws.xb;
                p.x_lo = ws.xl;
                p.ldo2 = D;
                p.ln_stats_in = ws.stats[0];
                p.ln_stats_out = ws.stats[1];
                p.ln_slots = S;
                p.ln_stride = SS;
            }
            const bool deep = (g_resid_ln_deep & 1) != 0;
            rc = launch_gemm(split ? (deep ? EPI_RESID_SP3 : EPI_RESID_SP) : fold ? (deep ? EPI_RESID_LN3 : EPI_RESID_LN) : EPI_RESID, cg, ws.h, D,
                             ctx->w<void>(l, CRE_W_O), D, p, ctx->num_sms, stream);
            if (rc) return rc;
        }
        if (!fold) {
            rc = launch_layernorm_bf16(ws.x, ctx->w<float>(l, CRE_LN2_G), ctx->w<float>(l, CRE_LN2_B), M, D, c.ln_eps, ws.h, stream);
            if (rc) return rc;
        }
        {
            GemmParams p = base_params(M, F, D);
            p.bias = fold ? ctx->fold_up[l].c2 : ctx->w<float>(l, CRE_B_UP);
            p.out_bf16 = ws.mlp;
            p.ldo = F;
            if (fold) {
                p.c1 = ctx->fold_up[l].c1;
                p.ln_stats_in = ws.stats[1];
                p.ln_slots = S;
                p.ln_stride = SS;
                p.ln_eps = c.ln_eps;
            }
            rc = launch_gemm(EPI_GELU, cg, fold ? ws.xb : ws.h, D, fold ? static_cast<const void*>(ctx->fold_up[l].w) : ctx->w<void>(l, CRE_W_UP),
                             D, p, ctx->num_sms, stream);
            if (rc) return rc;
        }
        {   // MLP down projection; the last block feeds the final norm (fp32 x only)
            const bool ln_out = split || (fold && l + 1 < c.layers);   // split: the final norm reads the halves too
            GemmParams p = base_params(M, D, F);
            p.bias = ctx->w<float>(l, CRE_B_DOWN);
            p.scale = ctx->w<float>(l, CRE_LS2);
            p.out_f32 = ws.x;
            p.ldo = D;
            if (ln_out) {
                p.out_bf16 = ws.xb;
                p.x_lo = ws.xl;
                p.ldo2 = D;
                p.ln_stats_in = ws.stats[1];
                p.ln_stats_out = ws.stats[0];
                p.ln_slots = S;
                p.ln_stride = SS;
            }
            const bool deep = (g_resid_ln_deep & 2) != 0;
            rc = launch_gemm(split ? (deep ? EPI_RESID_SP3 : EPI_RESID_SP) : ln_out ? (deep ? EPI_RESID_LN3 : EPI_RESID_LN) : EPI_RESID, cg, ws.mlp,
                             F, ctx->w<void>(l, CRE_W_DOWN), F, p, ctx->num_sms, stream);
            if (rc) return rc;
        }
    }
    return launch_final_norm_mean(split ? nullptr : ws.x, ws.xb, ws.xl, ws.stats[0], SS, ctx->w<float>(-1, CRE_LN_F_G),
                                  ctx->w<float>(-1, CRE_LN_F_B), n, T, D, c.ln_eps, out_frame_emb_dev, out_tokens_dev, stream);
}

int32_t cre_pool_clips(const float* frame_emb_dev, const int32_t* clip_offsets_dev, int32_t clips, int32_t dim,
                       float* out_mean_dev, float* out_unit_dev, void* stream) {
    CRE_REQUIRE(frame_emb_dev != nullptr && clip_offsets_dev != nullptr, "pool_clips: NULL argument");
    return launch_pool_clips(frame_emb_dev, clip_offsets_dev, clips, dim, out_mean_dev, out_unit_dev,
                             static_cast<cudaStream_t>(stream));
}

static const int kMaxSlots = 512;
static int g_scan_small = 1;   // tuning: 0 routes every query count through the tile GEMM
static int g_topk_stacked = 1; // tuning: 0 keeps two MMAs per k-step (hi and lo query tiles) also for <= 64 queries

int64_t cre_gallery_scratch_bytes(int32_t q, int32_t dim, int32_t k) {
    if (q <= 0 || dim <= 0 || k <= 0 || k > CRE_TOPK_LIMIT) {
        set_error("gallery_scratch_bytes: q=%d dim=%d k=%d (k <= %d)", q, dim, k, CRE_TOPK_LIMIT);
        return -1;
    }
    const int kp = k < CRE_TOPK_MAX ? k : CRE_TOPK_MAX;      // candidates kept per partial list and pass
    const int64_t a = align_up(static_cast<int64_t>(q) * 2 * dim * 2, 1024);
    const int64_t part = align_up(static_cast<int64_t>(q) * kMaxSlots * kp * 4, 1024);
    return a + 2 * part + 1024;      // + the "CTAs finished" counter of the serving-form scan (per call: two streams never share it)
}

// One pass of the gallery scan: the kp best candidates per query that come strictly AFTER (cut_s, cut_i) in the (score desc, index
// asc) order (cut_s == NULL: no cutoff), written to out_* with row stride out_stride.
static int gallery_topk_pass(cre_ctx* ctx, const float* queries_dev, int q, int dim, const void* gallery_dev, int rows, int row_base, int kp,
                             uint8_t* sp, int64_t counter_off, float* out_scores, int32_t* out_idx, int out_stride, const float* cut_s, const int32_t* cut_i,
                             float* dump_scores_dev, bool first_pass, cudaStream_t stream) {
    __nv_bfloat16* a_hilo = reinterpret_cast<__nv_bfloat16*>(sp);
    const int64_t a_bytes = align_up(static_cast<int64_t>(q) * 2 * dim * 2, 1024);
    const int64_t part_bytes = align_up(static_cast<int64_t>(q) * kMaxSlots * kp * 4, 1024);
    float* part_s = reinterpret_cast<float*>(sp + a_bytes);
    int32_t* part_i = reinterpret_cast<int32_t*>(sp + a_bytes + part_bytes);
    int* counter = reinterpret_cast<int*>(sp + counter_off);

    // serving form (Q <= 2: one message = one query): HBM-streaming scan on the CUDA cores, one partial list per CTA
    if (g_scan_small && q <= 2) {
        int small_slots = 2 * ctx->num_sms;
        if (small_slots > kMaxSlots) small_slots = kMaxSlots;
        if (first_pass) CRE_CUDA_OK(cudaMemsetAsync(counter, 0, 4, stream));   // the kernel's last CTA leaves it at 0 for the later passes
        const int took = launch_gallery_scan_small(queries_dev, q, dim, gallery_dev, rows, row_base, kp, part_s, part_i, small_slots,
                                                   dump_scores_dev, counter, out_scores, out_idx, out_stride, cut_s, cut_i, stream);
        if (took != 0) return took < 0 ? took : 0;     // the kernel's last CTA has merged the partial lists into the result
    }
    const int workers = gemm_workers(q, rows, 1, ctx->num_sms);
    const int slots = 2 * workers;
    CRE_REQUIRE(slots <= kMaxSlots, "gallery_topk: %d partial slots exceed %d", slots, kMaxSlots);
    // the hi / lo split of the queries survives from the first pass; later passes only refill the partial lists
    int rc = first_pass ? launch_topk_prepare(queries_dev, q, dim, a_hilo, part_s, part_i, static_cast<int64_t>(q) * slots * kp, stream)
                        : launch_fill_topk(part_s, part_i, static_cast<int64_t>(q) * slots * kp, stream);
    if (rc) return rc;
    GemmParams p = base_params(q, rows, 2 * dim);
    p.b_k_extent = dim;
    p.topk = kp;
    p.col_base = row_base;
    p.part_scores = part_s;
    p.part_idx = part_i;
    p.part_slots = slots;
    p.topk_stacked = (g_topk_stacked && q <= 64) ? 1 : 0;   // hi / lo halves as rows of ONE M = 128 tile: one MMA per k-step
    p.dump_scores = dump_scores_dev;
    p.cut_scores = cut_s;
    p.cut_idx = cut_i;
    p.cut_stride = out_stride;
    rc = launch_gemm(EPI_TOPK, 1, a_hilo, 2 * dim, gallery_dev, dim, p, ctx->num_sms, stream);
    if (rc) return rc;
    return launch_merge_topk(part_s, part_i, kp, static_cast<int64_t>(slots) * kp, slots, kp, q, kp, out_scores, out_idx, out_stride, stream);
}

int32_t cre_gallery_topk(cre_ctx* ctx, const float* queries_dev, int32_t q, int32_t dim, const void* gallery_dev,
                         int32_t rows, int32_t row_base, int32_t k, void* scratch_dev, int64_t scratch_bytes,
                         float* out_scores_dev, int32_t* out_idx_dev, float* dump_scores_dev, void* stream_) {
    CRE_REQUIRE(ctx != nullptr && queries_dev != nullptr && scratch_dev != nullptr && out_scores_dev != nullptr &&
                    out_idx_dev != nullptr, "gallery_topk: NULL argument");
    CRE_REQUIRE(q > 0 && dim > 0 && dim % 64 == 0 && rows >= 0, "gallery_topk: q=%d dim=%d rows=%d", q, dim, rows);
    CRE_REQUIRE(k >= 1 && k <= CRE_TOPK_LIMIT, "gallery_topk: k=%d out of range (1..%d)", k, CRE_TOPK_LIMIT);
    CRE_REQUIRE((reinterpret_cast<uintptr_t>(scratch_dev) & 255) == 0, "gallery_topk: scratch must be 256-byte aligned");
    const int64_t need = cre_gallery_scratch_bytes(q, dim, k);
    CRE_REQUIRE(scratch_bytes >= need, "gallery_topk: scratch %lld < required %lld bytes", (long long)scratch_bytes, (long long)need);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows == 0) return launch_fill_topk(out_scores_dev, out_idx_dev, static_cast<int64_t>(q) * k, stream);
    CRE_REQUIRE(gallery_dev != nullptr, "gallery_topk: NULL gallery");
    // k <= CRE_TOPK_MAX (the reference asks for 5 everywhere): one pass.  Larger k: ceil(k / 8) passes over the shard, pass j
    // keeping only candidates that rank strictly after entry 8 j - 1 of the result so far (read on the device: no host sync).
    uint8_t* sp = static_cast<uint8_t*>(scratch_dev);
    for (int done = 0; done < k; done += CRE_TOPK_MAX) {
        const int kp = k - done < CRE_TOPK_MAX ? k - done : CRE_TOPK_MAX;
        const int rc = gallery_topk_pass(ctx, queries_dev, q, dim, gallery_dev, rows, row_base, kp, sp, need - 1024, out_scores_dev + done, out_idx_dev + done,
                                         k, done ? out_scores_dev + done - 1 : nullptr, done ? out_idx_dev + done - 1 : nullptr,
                                         done ? nullptr : dump_scores_dev, done == 0, stream);
        if (rc) return rc;
    }
    return 0;
}

int32_t cre_merge_topk(const float* scores_dev, const int32_t* idx_dev, int32_t lists, int32_t q, int32_t k,
                       float* out_scores_dev, int32_t* out_idx_dev, void* stream) {
    CRE_REQUIRE(scores_dev != nullptr && idx_dev != nullptr && out_scores_dev != nullptr && out_idx_dev != nullptr,
                "merge_topk: NULL argument");
    return launch_merge_topk(scores_dev, idx_dev, static_cast<int64_t>(q) * k, k, lists, k, q, k, out_scores_dev, out_idx_dev, k,
                             static_cast<cudaStream_t>(stream));
}

int32_t cre_gallery_update_row(void* gallery_dev, float* master_dev, int32_t dim, int32_t row, const float* unit_query_dev,
                               float momentum, void* stream) {
    CRE_REQUIRE(gallery_dev != nullptr && unit_query_dev != nullptr, "gallery_update_row: NULL argument");
    return launch_gallery_update_row(static_cast<__nv_bfloat16*>(gallery_dev), master_dev, dim, row, unit_query_dev, momentum,
                                     static_cast<cudaStream_t>(stream));
}

int32_t cre_gemm_bf16(cre_ctx* ctx, const void* a_dev, const void* b_dev, int32_t m, int32_t n, int32_t k,
                      int32_t epilogue, const float* bias_dev, const float* scale_dev, void* out_dev, int32_t cta_group,
                      void* stream) {
    CRE_REQUIRE(ctx != nullptr && a_dev != nullptr && b_dev != nullptr && out_dev != nullptr, "gemm: NULL argument");
    CRE_REQUIRE(epilogue == CRE_EPI_BF16 || epilogue == CRE_EPI_F32 || epilogue == CRE_EPI_GELU || epilogue == CRE_EPI_RESID ||
                    epilogue == CRE_EPI_NONE, "gemm: epilogue %d is not exposed", epilogue);
    CRE_REQUIRE(epilogue != CRE_EPI_RESID || scale_dev != nullptr, "gemm: RESID epilogue needs scale");
    GemmParams p = base_params(m, n, k);
    p.bias = bias_dev;
    p.scale = scale_dev;
    p.out_f32 = static_cast<float*>(out_dev);
    p.out_bf16 = static_cast<__nv_bfloat16*>(out_dev);
    p.ldo = n;
    return launch_gemm(epilogue, cta_group, a_dev, k, b_dev, k, p, ctx->num_sms, static_cast<cudaStream_t>(stream));
}

int32_t cre_row_stats(const float* x_dev, int32_t rows, int32_t dim, void* out_xb_dev, float* out_stats_dev, void* stream) {
    CRE_REQUIRE(x_dev != nullptr && out_xb_dev != nullptr && out_stats_dev != nullptr, "row_stats: NULL argument");
    return launch_row_stats(x_dev, rows, dim, 2 * (dim / 128) + 4, static_cast<__nv_bfloat16*>(out_xb_dev), nullptr, out_stats_dev,
                            static_cast<cudaStream_t>(stream));
}

int32_t cre_row_stats_split(const float* x_dev, int32_t rows, int32_t dim, void* out_hi_dev, void* out_lo_dev, float* out_stats_dev,
                            void* stream) {
    CRE_REQUIRE(x_dev != nullptr && out_hi_dev != nullptr && out_lo_dev != nullptr && out_stats_dev != nullptr, "row_stats_split: NULL argument");
    return launch_row_stats(x_dev, rows, dim, 2 * (dim / 128) + 4, static_cast<__nv_bfloat16*>(out_hi_dev),
                            static_cast<__nv_bfloat16*>(out_lo_dev), out_stats_dev, static_cast<cudaStream_t>(stream));
}

int32_t cre_fold_ln_weights(const void* w_dev, const float* gamma_dev, const float* beta_dev, const float* bias_dev, int32_t n,
                            int32_t k, int32_t scaled_rows, float row_scale, void* out_w_dev, float* out_c1_dev, float* out_c2_dev,
                            void* stream) {
    CRE_REQUIRE(w_dev != nullptr && gamma_dev != nullptr && beta_dev != nullptr && out_w_dev != nullptr && out_c1_dev != nullptr &&
                    out_c2_dev != nullptr, "fold_ln_weights: NULL argument");
    return launch_fold_ln_weights(static_cast<const __nv_bfloat16*>(w_dev), gamma_dev, beta_dev, bias_dev, n, k, scaled_rows, row_scale,
                                  static_cast<__nv_bfloat16*>(out_w_dev), out_c1_dev, out_c2_dev, static_cast<cudaStream_t>(stream));
}

int32_t cre_gemm_ln(cre_ctx* ctx, const void* a_dev, const void* b_dev, int32_t m, int32_t n, int32_t k, int32_t epilogue,
                    const float* bias_dev, const float* c1_dev, const float* scale_dev, const float* stats_in_dev, int32_t ln_dim,
                    float ln_eps, void* out_dev, void* out_xb_dev, float* stats_out_dev, int32_t cta_group, void* stream) {
    CRE_REQUIRE(ctx != nullptr && a_dev != nullptr && b_dev != nullptr && out_dev != nullptr && stats_in_dev != nullptr,
                "gemm_ln: NULL argument");
    CRE_REQUIRE(epilogue == CRE_EPI_BF16 || epilogue == CRE_EPI_GELU || epilogue == CRE_EPI_RESID_LN || epilogue == CRE_EPI_RESID_LN3 ||
                    epilogue == CRE_EPI_RESID_SP || epilogue == CRE_EPI_RESID_SP3, "gemm_ln: epilogue %d is not exposed", epilogue);
    CRE_REQUIRE(ln_dim > 0 && ln_dim % 128 == 0 && ln_dim / 128 <= 8, "gemm_ln: ln_dim=%d", ln_dim);
    GemmParams p = base_params(m, n, k);
    p.bias = bias_dev;
    p.c1 = c1_dev;
    p.scale = scale_dev;
    p.ln_stats_in = stats_in_dev;
    p.ln_stats_out = stats_out_dev;
    p.ln_slots = ln_dim / 128;
    p.ln_stride = 2 * p.ln_slots + 4;
    p.ln_eps = ln_eps;
    p.ldo = n;
    if (epilogue == CRE_EPI_RESID_LN || epilogue == CRE_EPI_RESID_LN3) {
        CRE_REQUIRE(out_xb_dev != nullptr && stats_out_dev != nullptr && n == ln_dim, "gemm_ln: RESID_LN needs out_xb, stats_out and n == ln_dim");
        p.out_f32 = static_cast<float*>(out_dev);
        p.out_bf16 = static_cast<__nv_bfloat16*>(out_xb_dev);
        p.ldo2 = n;
    } else if (epilogue == CRE_EPI_RESID_SP || epilogue == CRE_EPI_RESID_SP3) {
        CRE_REQUIRE(out_xb_dev != nullptr && stats_out_dev != nullptr && n == ln_dim, "gemm_ln: RESID_SP needs both halves, stats_out and n == ln_dim");
        p.x_lo = static_cast<__nv_bfloat16*>(out_dev);          // low half, in place
        p.out_bf16 = static_cast<__nv_bfloat16*>(out_xb_dev);   // high half, in place
        p.ldo2 = n;
    } else {
        p.out_bf16 = static_cast<__nv_bfloat16*>(out_dev);
    }
    return launch_gemm(epilogue, cta_group, a_dev, k, b_dev, k, p, ctx->num_sms, static_cast<cudaStream_t>(stream));
}

int32_t cre_layernorm_bf16(const float* x_dev, const float* gamma_dev, const float* beta_dev, int32_t rows, int32_t dim,
                           float eps, void* out_dev, void* stream) {
    CRE_REQUIRE(x_dev != nullptr && gamma_dev != nullptr && beta_dev != nullptr && out_dev != nullptr, "layernorm: NULL argument");
    return launch_layernorm_bf16(x_dev, gamma_dev, beta_dev, rows, dim, eps, static_cast<__nv_bfloat16*>(out_dev),
                                 static_cast<cudaStream_t>(stream));
}

int64_t cre_attention_scratch_bytes(int32_t n, int32_t heads) {
    if (n <= 0 || heads <= 0) {
        set_error("attention_scratch_bytes: n=%d heads=%d", n, heads);
        return -1;
    }
    return align_up(attention_flag_ints(n, heads) * 4, kAlign);
}

int32_t cre_attention(cre_ctx* ctx, const void* qkv_dev, int32_t ld, int32_t k_col0, int32_t v_col0, int32_t n, int32_t t,
                      int32_t heads, void* out_dev, void* scratch_dev, int64_t scratch_bytes, void* stream) {
    CRE_REQUIRE(ctx != nullptr && qkv_dev != nullptr && out_dev != nullptr && scratch_dev != nullptr, "attention: NULL argument");
    CRE_REQUIRE(n > 0 && heads > 0, "attention: n=%d heads=%d", n, heads);
    CRE_REQUIRE((reinterpret_cast<uintptr_t>(scratch_dev) & 255) == 0, "attention: scratch must be 256-byte aligned");
    const int64_t need = cre_attention_scratch_bytes(n, heads);
    CRE_REQUIRE(scratch_bytes >= need, "attention: scratch %lld < required %lld bytes", (long long)scratch_bytes, (long long)need);
    CRE_CUDA_OK(cudaMemsetAsync(scratch_dev, 0, static_cast<size_t>(need), static_cast<cudaStream_t>(stream)));
    AttnArgs a;
    a.any_flag = static_cast<int*>(scratch_dev);
    a.unit_flags = static_cast<int*>(scratch_dev) + 64;
    a.qkv = qkv_dev;
    a.ld = ld;
    a.k_col0 = k_col0;
    a.v_col0 = v_col0;
    a.n = n;
    a.t = t;
    a.heads = heads;
    a.out = out_dev;
    return launch_attention(a, static_cast<cudaStream_t>(stream));
}

int64_t cre_kernel_launches(void) { return launch_count(); }
int32_t cre_profile_start(int32_t max_launches) { return profile_start(max_launches); }
int32_t cre_profile_stop(int32_t* ids_out, float* ms_out, double* work_out, int32_t cap) {
    CRE_REQUIRE(ids_out != nullptr && ms_out != nullptr && work_out != nullptr && cap >= 0, "profile_stop: NULL argument");
    return profile_stop(ids_out, ms_out, work_out, cap);
}

// generic process-wide tuning knobs (benchmark / tuning harness; defaults are what the parity tests cover)
int32_t cre_set_tuning(const char* key, int32_t value) {
    CRE_REQUIRE(key != nullptr, "set_tuning: NULL key");
    if (strcmp(key, "cta_group") == 0) return cre_set_cta_group(value);
    if (strcmp(key, "gemm_stages") == 0) {
        CRE_REQUIRE(value == 0 || (value >= 3 && value <= 6), "set_tuning: gemm_stages=%d", value);
        set_gemm_stages(value);
        return 0;
    }
    if (strcmp(key, "attention_fast") == 0) {
        set_attention_fast(value);
        return 0;
    }
    if (strcmp(key, "attention_split") == 0) {
        set_attention_split(value);
        return 0;
    }
    if (strcmp(key, "attention_poly") == 0) {
        CRE_REQUIRE(value >= 0 && value <= 2, "set_tuning: attention_poly=%d", value);
        set_attention_poly(value);
        return 0;
    }
    if (strcmp(key, "attention_long") == 0) {
        set_attention_long(value);
        return 0;
    }
    if (strcmp(key, "attention_split_mode") == 0) {
        set_attention_split_mode(value);
        return 0;
    }
    if (strcmp(key, "attention_split_delay") == 0) {
        CRE_REQUIRE(value >= 0 && value <= 100000, "set_tuning: attention_split_delay=%d cycles", value);
        set_attention_split_delay(value);
        return 0;
    }
    if (strcmp(key, "preprocess_tma") == 0) {
        set_preprocess_tma(value);
        return 0;
    }
    if (strcmp(key, "scan_small") == 0) {
        g_scan_small = value != 0;
        return 0;
    }
    if (strcmp(key, "preprocess_identity") == 0) {   // 0: frames that need no resize go through the filtering kernel anyway
        set_preprocess_identity(value);
        return 0;
    }
    if (strcmp(key, "resid_ln_deep") == 0) {
        g_resid_ln_deep = value & 3;
        return 0;
    }
    if (strcmp(key, "ln_fold") == 0) {
        g_ln_fold = value != 0;
        return 0;
    }
    if (strcmp(key, "topk_stacked") == 0) {
        g_topk_stacked = value != 0;
        return 0;
    }
    if (strcmp(key, "resid_split") == 0) {
        g_resid_split = value != 0;
        return 0;
    }
#ifdef CRE_TUNING
    if (strcmp(key, "gemm_debug") == 0) {   // tuning builds only (make EXTRA=-DCRE_TUNING): results are garbage
        set_gemm_debug(value);
        return 0;
    }
#endif
    set_error("set_tuning: unknown key '%s'", key);
    return -1;
}

#ifdef CRE_ATTN_TRACE
// tracing builds only (tools/attn_trace.py): a device buffer of 12 x 256 uint64 that CTA 0 of the split-S attention kernel fills
// with (event << 56 | clock64) records; not part of include/cre.h
int32_t cre_debug_set_attention_trace(void* buf_dev) {
    set_attention_trace(static_cast<unsigned long long*>(buf_dev));
    return 0;
}
#endif

// 1 = one CTA per tile (cta_group::1), 2 = CTA pairs (cta_group::2) for the ViT GEMMs
int32_t cre_set_cta_group(int32_t cg) {
    CRE_REQUIRE(cg == 1 || cg == 2, "set_cta_group: %d", cg);
    g_default_cg = cg;
    return 0;
}

}  // extern "C"
