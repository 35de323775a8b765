// K1: fused frame preprocessing.  uint8 NHWC (BGR or RGB) -> antialiased bilinear resize -> x/255 ->
// (x - mean)/std -> 16x16 patchify -> bf16 patch rows [n * gh * gw, 768] with K index c*256 + ky*16 + kx
// (= Conv2d weight .view(D, -1) order, HF:modeling_dinov3_vit.py:71-81).
//
// Replaces cv2.cvtColor + PIL + DINOv3ViTImageProcessor._preprocess
// (services/dinov3-pipeline/app/main.py:98-107; HF:image_processing_dinov3_vit.py:45-86 ->
// torchvision resize(antialias=True) -> aten _upsample_bilinear2d_aa).  The separable triangle-filter
// weights are the aten ones (see build_resize_table in api.cu).
//
// One CTA = one patch row (16 output rows) x a band of kBandPatches patches of one frame.
//   pass 1 (HBM-bound): thread (r, v) accumulates output row r over its <= kmax_y input rows for the
//           16-byte column group v of the band -- coalesced 16-byte loads, fp32 accumulators in
//           registers -- and parks the vertically filtered row in shared memory.
//   pass 2: thread = one output element in patch-row order (c, ky, kx) -> horizontal taps from shared
//           memory, normalise, bf16, fully coalesced 1536-byte stores per patch.
// Each input byte is read from DRAM once (neighbouring CTAs share at most the filter support via L2).
#include "common.cuh"
#include "gemm_tcgen05.cuh"
#include "internal.h"

namespace cre {

constexpr int kBandPatches = 2;
constexpr int kPreThreads = 448;   // x3 CTAs per SM (<= 48 registers)

struct PreParams {
    const uint8_t* frames;
    const uint8_t* frames_end;   // one past the last byte of the last frame's last row (the unaligned identity path never reads beyond it)
    int h, w;
    int64_t row_pitch, frame_pitch;
    int bgr, gh, gw;
    float mean[3], inv_std[3];
    __nv_bfloat16* out;
    const int32_t *ylo, *ycnt, *xlo, *xcnt;
    const float *yw, *xw;
    int ykmax, xkmax;
    int sstride;  // floats per vertically-filtered row in shared memory
    int vec;      // 1: 16-byte aligned rows, vector loads allowed
    // region-of-interest mode (cre_preprocess_patchify_roi): blockIdx.z = ROI, rois[5 * z] = {frame, x0, y0, x1, y1}; the six
    // tables above then hold one block per ROI (oh / ow entries, oh * ykmax / ow * xkmax weights), built by build_roi_tables_kernel
    const int32_t* rois;
};

// Vertical pass, integer form (every K1 variant evaluates exactly this, so they agree bit for bit):
//     w16_k = rint(w_k * 2^15)           the aten antialias weight of tap k as an unsigned 16-bit fixed-point number
//     acc   = sum_k w16_k * b_k          exact in int32 (<= 255 * (2^15 + taps / 2) < 2^23)
//     v     = float(acc) * (1 / float(sum_k w16_k))        one rounding; the weights are renormalised to sum to 1
// Quantising the weights to 2^-16 moves a filtered pixel by at most 255 * taps * 2^-16 = 0.04 grey levels (1.7e-4 of the range, 7e-4
// after normalisation: a tenth of the bf16 half-ulp of the output).  Two taps cost ONE IDP.2A per byte column on packed bytes --
// the bytes of rows k and k + 1 interleaved by two byte-permutes per word -- instead of a byte-permute + half an FFMA2 per byte
// and tap: 0.8 instead of 1.6 instructions per byte-tap in the stage that bounds the kernel.
__device__ __forceinline__ uint32_t aa_weight16(float w) { return __float2uint_rn(w * 32768.0f); }
__device__ __forceinline__ uint32_t dp2a_lo(uint32_t w2, uint32_t bytes, uint32_t acc) {
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w2), "r"(bytes), "r"(acc));
    return d;
}
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t w2, uint32_t bytes, uint32_t acc) {
    uint32_t d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w2), "r"(bytes), "r"(acc));
    return d;
}
// 16 byte columns x two taps: acc[j] += w16_a * row_a[j] + w16_b * row_b[j]   (w2 = w16_a | w16_b << 16)
__device__ __forceinline__ void vtaps2(const uint4& qa, const uint4& qb, uint32_t w2, uint32_t (&acc)[16]) {
    const uint32_t a[4] = {qa.x, qa.y, qa.z, qa.w}, b[4] = {qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t lo = __byte_perm(a[i], b[i], 0x5140u);     // (a0, b0, a1, b1)
        const uint32_t hi = __byte_perm(a[i], b[i], 0x7362u);     // (a2, b2, a3, b3)
        acc[4 * i] = dp2a_lo(w2, lo, acc[4 * i]);
        acc[4 * i + 1] = dp2a_hi(w2, lo, acc[4 * i + 1]);
        acc[4 * i + 2] = dp2a_lo(w2, hi, acc[4 * i + 2]);
        acc[4 * i + 3] = dp2a_hi(w2, hi, acc[4 * i + 3]);
    }
}
// acc (< 2^23) -> float without the XU-pipe conversion: 2^23 + acc is exact in fp32, so is the subtraction
__device__ __forceinline__ float acc_to_float(uint32_t acc) { return __uint_as_float(acc | 0x4B000000u) - 8388608.0f; }

// (x / 255 - mean) / std exactly as every K1 variant evaluates it (one definition so that the compiler contracts it the same way
// everywhere: the variants must agree bit for bit)
__device__ __forceinline__ float normalise_px(float s, float mean, float inv_std) { return (s * (1.0f / 255.0f) - mean) * inv_std; }

__global__ void __launch_bounds__(kPreThreads, 2) preprocess_kernel(const PreParams p) {
    extern __shared__ __align__(16) float vbuf[];  // [16][sstride]
    const int band = blockIdx.x, py = blockIdx.y;
    int frame = blockIdx.z, cx0 = 0, cy0 = 0;   // crop origin inside the frame (ROI mode)
    const int32_t *t_ylo = p.ylo, *t_ycnt = p.ycnt, *t_xlo = p.xlo, *t_xcnt = p.xcnt;
    const float *t_yw = p.yw, *t_xw = p.xw;
    if (p.rois != nullptr) {
        const int32_t* roi = p.rois + 5 * blockIdx.z;
        frame = roi[0];
        cx0 = roi[1];
        cy0 = roi[2];
        const size_t oh = static_cast<size_t>(p.gh) * 16, ow = static_cast<size_t>(p.gw) * 16;
        t_ylo += blockIdx.z * oh; t_ycnt += blockIdx.z * oh; t_yw += blockIdx.z * oh * p.ykmax;
        t_xlo += blockIdx.z * ow; t_xcnt += blockIdx.z * ow; t_xw += blockIdx.z * ow * p.xkmax;
    }
    const int px0 = band * kBandPatches;
    const int npatch = min(kBandPatches, p.gw - px0);
    const int ox0 = px0 * 16, ox1 = ox0 + npatch * 16;  // output column range [ox0, ox1)
    const int x_lo = t_xlo[ox0];
    const int x_hi = t_xlo[ox1 - 1] + t_xcnt[ox1 - 1];
    // first byte column staged, relative to the crop origin; 16-byte aligned in the FRAME row when vector loads are allowed
    const int b0 = p.vec ? ((((cx0 + x_lo) * 3) & ~15) - cx0 * 3) : x_lo * 3;
    const int nbytes = x_hi * 3 - b0;
    const int nvec = (nbytes + 15) >> 4;
    const int row_bytes = (p.w - cx0) * 3;      // readable bytes of a frame row from the crop origin on
    const uint8_t* fbase = p.frames + static_cast<int64_t>(frame) * p.frame_pitch + static_cast<int64_t>(cy0) * p.row_pitch + cx0 * 3;
    const int nthreads = blockDim.x;

    // ---- pass 1: vertical filter (each item = one output row x one 16-byte column group) ----
    for (int item = threadIdx.x; item < 16 * nvec; item += nthreads) {
        const int r = item / nvec, v = item - r * nvec;
        const int oy = py * 16 + r;
        const int y0 = t_ylo[oy], cnt = t_ycnt[oy];
        const float* wy = t_yw + static_cast<size_t>(oy) * p.ykmax;
        const int col = b0 + v * 16;
        float acc[16];
        uint32_t iacc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) iacc[i] = 0u;
        uint32_t wtot = 0u;
        const uint8_t* src = fbase + static_cast<int64_t>(y0) * p.row_pitch + col;
        if (p.vec && (col + 16 <= row_bytes)) {
#pragma unroll 3
            for (int k = 0; k < cnt; k += 2) {
                const bool two = k + 1 < cnt;
                const uint32_t wa = aa_weight16(__ldg(wy + k)), wb = two ? aa_weight16(__ldg(wy + k + 1)) : 0u;
                const uint4 qa = __ldg(reinterpret_cast<const uint4*>(src + static_cast<int64_t>(k) * p.row_pitch));
                const uint4 qb = two ? __ldg(reinterpret_cast<const uint4*>(src + static_cast<int64_t>(k + 1) * p.row_pitch)) : qa;
                wtot += wa + wb;
                vtaps2(qa, qb, wa | (wb << 16), iacc);
            }
        } else {
            // unaligned / ragged edge: byte loads, the same integer sums
            for (int k = 0; k < cnt; ++k) {
                const uint32_t wk = aa_weight16(__ldg(wy + k));
                wtot += wk;
                const uint8_t* s8 = src + static_cast<int64_t>(k) * p.row_pitch;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (col + i < row_bytes) iacc[i] += wk * static_cast<uint32_t>(__ldg(s8 + i));
            }
        }
        {
            const float inv = 1.0f / static_cast<float>(wtot);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = acc_to_float(iacc[i]) * inv;
        }
        float4* dst = reinterpret_cast<float4*>(vbuf + static_cast<size_t>(r) * p.sstride + v * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    }
    __syncthreads();

    // ---- pass 2: horizontal filter + normalise + patchify; one thread = one output pixel, all 3 channels ----
    const int nout = npatch * 256;
    for (int t = threadIdx.x; t < nout; t += nthreads) {
        const int pl = t >> 8, ky = (t >> 4) & 15, kx = t & 15;
        const int ox = (px0 + pl) * 16 + kx;
        const int x0 = t_xlo[ox], cnt = t_xcnt[ox];
        const float* wx = t_xw + static_cast<size_t>(ox) * p.xkmax;
        const float* src = vbuf + static_cast<size_t>(ky) * p.sstride + (x0 * 3 - b0);
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const float w = wx[k];
            s0 = fmaf(w, src[3 * k], s0);
            s1 = fmaf(w, src[3 * k + 1], s1);
            s2 = fmaf(w, src[3 * k + 2], s2);
        }
        const float sm[3] = {s0, s1, s2};   // memory channel order
        const size_t patch = (static_cast<size_t>(blockIdx.z) * p.gh + py) * p.gw + px0 + pl;
        __nv_bfloat16* o = p.out + patch * 768 + ky * 16 + kx;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int c = p.bgr ? 2 - j : j;  // output (RGB) channel of memory channel j
            o[c * 256] = __float2bfloat16_rn(normalise_px(sm[j], p.mean[c], p.inv_std[c]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// No-resize variant (frame size == model input size: the 224 x 224 clips of BASELINE configs[0] / [3] / [4], 518 / 592 inputs):
// the antialias tables degenerate to the identity (weights {1, 0}), so K1 is normalise + patchify and purely HBM-bound
// (3 B in, 6 B out per pixel).  Thread = one 16-pixel row of one patch: 48 contiguous input bytes (three 16-byte loads; a warp
// covers two neighbouring patches, i.e. 96 contiguous bytes per image row) -> per channel 32 contiguous output bytes (16 threads
// of a patch write 512 contiguous bytes).  Same values as the filtered path with identity taps -- 2^15 b / 2^15 = b exactly
// -> normalise_px -> bf16 -- so the result is bit-identical to preprocess_kernel on the same input.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) preprocess_identity_kernel(const PreParams p, int64_t total) {
    const int64_t gid = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (gid >= total) return;
    const int ky = static_cast<int>(gid & 15);
    const int64_t patch = gid >> 4;
    const int px = static_cast<int>(patch % p.gw);
    const int64_t t = patch / p.gw;
    const int py = static_cast<int>(t % p.gh);
    const int64_t frame = t / p.gh;
    const uint8_t* src = p.frames + frame * p.frame_pitch + static_cast<int64_t>(py * 16 + ky) * p.row_pitch + px * 48;
    uint32_t wds[12];
    if (p.vec) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + i);
            wds[4 * i] = q.x; wds[4 * i + 1] = q.y; wds[4 * i + 2] = q.z; wds[4 * i + 3] = q.w;
        }
    } else {
        // rows that are not 16-byte aligned (e.g. 518-pixel rows: pitch 1 554 bytes): the four aligned vectors that cover the 48 bytes,
        // shifted into place in registers (two select stages for the word offset, a funnel shift for the byte offset).  The fourth vector
        // is only touched when the row really reaches into it, and never beyond the end of the frame buffer.
        const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
        const uint32_t sh = static_cast<uint32_t>(addr & 15);
        const uint4* base = reinterpret_cast<const uint4*>(addr - sh);
        uint32_t w[17];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint4 q = __ldg(base + i);
            w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w;
        }
        uint4 q3 = make_uint4(0u, 0u, 0u, 0u);
        if (sh != 0) {                           // bytes [48 - sh, 48) of the row live in the fourth vector
            const uint8_t* v3 = reinterpret_cast<const uint8_t*>(base + 3);
            if (v3 + 16 <= p.frames_end) {
                q3 = __ldg(base + 3);
            } else {                             // the very last rows of the buffer: only the bytes that exist
                uint32_t t[4] = {0u, 0u, 0u, 0u};
                for (int j = 0; j < 16 && v3 + j < p.frames_end; ++j) t[j >> 2] |= static_cast<uint32_t>(__ldg(v3 + j)) << (8 * (j & 3));
                q3 = make_uint4(t[0], t[1], t[2], t[3]);
            }
        }
        w[12] = q3.x; w[13] = q3.y; w[14] = q3.z; w[15] = q3.w; w[16] = 0u;
        const bool b0 = (sh & 4) != 0, b1 = (sh & 8) != 0;
        uint32_t v1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v1[i] = b0 ? w[i + 1] : w[i];
        uint32_t v2[13];
#pragma unroll
        for (int i = 0; i < 13; ++i) v2[i] = b1 ? v1[i + 2] : v1[i];
        const uint32_t bits = (sh & 3) * 8;
#pragma unroll
        for (int i = 0; i < 12; ++i) wds[i] = __funnelshift_r(v2[i], v2[i + 1], bits);
    }
    float v[48];   // byte 3 i + j = pixel i, memory channel j
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        // byte j -> 2^23 + b (one byte-permute) -> b: what the integer vertical pass yields for identity taps {2^15, 0}
        v[4 * i] = __uint_as_float(__byte_perm(wds[i], 0x4B000000u, 0x7650u)) - 8388608.0f;
        v[4 * i + 1] = __uint_as_float(__byte_perm(wds[i], 0x4B000000u, 0x7651u)) - 8388608.0f;
        v[4 * i + 2] = __uint_as_float(__byte_perm(wds[i], 0x4B000000u, 0x7652u)) - 8388608.0f;
        v[4 * i + 3] = __uint_as_float(__byte_perm(wds[i], 0x4B000000u, 0x7653u)) - 8388608.0f;
    }
    __nv_bfloat16* o = p.out + patch * 768 + ky * 16;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int c = p.bgr ? 2 - j : j;   // output (RGB) channel of memory channel j
        const float mean = p.mean[c], inv_std = p.inv_std[c];
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            pk[i] = pack_bf16x2(normalise_px(v[6 * i + j], mean, inv_std), normalise_px(v[6 * i + 3 + j], mean, inv_std));
        uint4* dst = reinterpret_cast<uint4*>(o + c * 256);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
}

// Separable antialias tables for every region of interest: the arithmetic of build_resize_table (api.cu; aten
// _compute_indices_min_size_weights_aa, triangle filter) in the same operation order, without FMA contraction, so that a ROI
// covering the whole frame reproduces the cached full-frame tables bit for bit.  One thread per (ROI, output index) of both axes.
__global__ void build_roi_tables_kernel(const int32_t* __restrict__ rois, int n_rois, int oh, int ow, int ykmax, int xkmax,
                                        int32_t* __restrict__ ylo, int32_t* __restrict__ ycnt, float* __restrict__ yw,
                                        int32_t* __restrict__ xlo, int32_t* __restrict__ xcnt, float* __restrict__ xw) {
    const int per = oh + ow;
    const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= static_cast<int64_t>(n_rois) * per) return;
    const int r = static_cast<int>(gid / per), e = static_cast<int>(gid % per);
    const bool is_y = e < oh;
    const int i = is_y ? e : e - oh;
    const int32_t* roi = rois + 5 * r;
    const int in = is_y ? roi[4] - roi[2] : roi[3] - roi[1];
    const int out = is_y ? oh : ow, kmax = is_y ? ykmax : xkmax;
    const float scale = __fdiv_rn(static_cast<float>(in), static_cast<float>(out));
    const float support = scale >= 1.0f ? scale : 1.0f;
    const float invscale = scale >= 1.0f ? __fdiv_rn(1.0f, scale) : 1.0f;
    const float center = __fmul_rn(scale, __fadd_rn(static_cast<float>(i), 0.5f));
    int xmin = static_cast<int>(__fadd_rn(__fsub_rn(center, support), 0.5f));
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(__fadd_rn(__fadd_rn(center, support), 0.5f));
    if (xmax > in) xmax = in;
    const int xsize = xmax - xmin;
    float* wi = (is_y ? yw : xw) + (static_cast<size_t>(r) * out + i) * kmax;
    float total = 0.0f;
    for (int j = 0; j < kmax; ++j) {
        float v = 0.0f;
        if (j < xsize) {
            const float x = fabsf(__fmul_rn(__fadd_rn(__fsub_rn(static_cast<float>(j + xmin), center), 0.5f), invscale));
            v = x < 1.0f ? __fsub_rn(1.0f, x) : 0.0f;
            total = __fadd_rn(total, v);
        }
        wi[j] = v;
    }
    if (total != 0.0f)
        for (int j = 0; j < xsize && j < kmax; ++j) wi[j] = __fdiv_rn(wi[j], total);
    (is_y ? ylo : xlo)[static_cast<size_t>(r) * out + i] = xmin;
    (is_y ? ycnt : xcnt)[static_cast<size_t>(r) * out + i] = xsize < kmax ? xsize : kmax;
}

int launch_build_roi_tables(const int32_t* rois, int n_rois, int oh, int ow, int ykmax, int xkmax, int32_t* ylo, int32_t* ycnt,
                            float* yw, int32_t* xlo, int32_t* xcnt, float* xw, cudaStream_t stream) {
    const int64_t total = static_cast<int64_t>(n_rois) * (oh + ow);
    LaunchScope scope(CRE_K_ROI_TABLES, 4.0 * total * (ykmax + xkmax), stream);
    build_roi_tables_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(rois, n_rois, oh, ow, ykmax, xkmax, ylo, ycnt,
                                                                                             yw, xlo, xcnt, xw);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

// =====================================================================================================================
// TMA-staged, warp-specialised variant (16-byte aligned frames and pitches; the tile must fit shared memory twice).
// Persistent: one CTA per SM walks tiles (frame, patch row, patch) in raster order.
//
//   warp 0 (one lane)   tile scheduler + TMA producer: the raw uint8 rows the tile's 16 output rows depend on -- rows_tile rows
//                       x nbox 256-byte boxes -- land in a 2-deep ring; no thread ever waits on a global load of pixels.
//   warps 9..24         pass 1, warp 9 + r = output row r of the patch, lane = 16-byte column group: vertical taps from shared
//                       memory (same arithmetic and order as preprocess_kernel: results are bit-identical) -> vbuf ring.
//   warps 1..8          pass 2: horizontal taps + normalise + bf16 + patch-order stores, one pixel per thread.
//
// The three roles are connected by mbarriers only (raw full/empty, vbuf full/empty): no __syncthreads in the tile loop.
// =====================================================================================================================
constexpr int kPtRowWarps = 16, kPtColWarps = 8;
constexpr int kPtThreads = (kPtRowWarps + kPtColWarps + 1) * 32;
// warp roles by index: the scheduler is warp 0, the column warps follow, the row warps come LAST -- the SM's warp arbiter prefers
// the highest warp id among the eligible ones, and the row warps (ALU-pipe bound) are the stage everything else waits for
constexpr int kPtColWarp0 = 1, kPtRowWarp0 = 1 + kPtColWarps;

// wait with back-off: the waiting roles (column warps, scheduler) would otherwise spend the row warps' issue slots on polling
// (ncu: 28 % of the kernel's executed instructions were mbarrier polls)
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, unsigned ns = 200) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        if (mbar_try_wait(bar, parity)) return;
        __nanosleep(ns);
    }
    __trap();
}
constexpr int kPtParamInts = 8;   // per tile: frame, py, px, b0, y_first, nvec

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(kEvictFirst)
        : "memory");
}

__global__ void __launch_bounds__(kPtThreads, 1)
preprocess_tma_kernel(const __grid_constant__ CUtensorMap tmap, const PreParams p, int rows_tile, int nbox, int tiles) {
    extern __shared__ uint8_t smem_raw_[];
    const uint32_t base_u32 = (smem_u32(smem_raw_) + 127u) & ~127u;
    uint8_t* base = smem_raw_ + (base_u32 - smem_u32(smem_raw_));
    const int raw_bytes = nbox * rows_tile * 256;
    const int vb_floats = 16 * p.sstride;
    float* vb0 = reinterpret_cast<float*>(base + 2 * raw_bytes);
    int* params = reinterpret_cast<int*>(base + 2 * raw_bytes + 2 * vb_floats * 4);          // [4][kPtParamInts]
    const uint32_t bar0 = base_u32 + 2 * raw_bytes + 2 * vb_floats * 4 + 4 * kPtParamInts * 4;
    // resize tables, copied once per CTA: every tap weight / window lookup in the tile loop is a shared-memory read
    const int oh = p.gh * 16, ow = p.gw * 16;
    float* s_yw = reinterpret_cast<float*>(base + 2 * raw_bytes + 2 * vb_floats * 4 + 4 * kPtParamInts * 4 + 64);
    float* s_xw = s_yw + oh * p.ykmax;
    int* s_ylo = reinterpret_cast<int*>(s_xw + ow * p.xkmax);
    int* s_ycnt = s_ylo + oh;
    int* s_xlo = s_ycnt + oh;
    int* s_xcnt = s_xlo + ow;
    // vertical taps in the integer form (see vtaps2): per output row its taps as 16-bit pairs, and 1 / (their sum)
    const int ypairs = (p.ykmax + 1) >> 1;
    uint32_t* s_yw16 = reinterpret_cast<uint32_t*>(s_xcnt + ow);
    float* s_yinv = reinterpret_cast<float*>(s_yw16 + oh * ypairs);
    for (int oy = threadIdx.x; oy < oh; oy += kPtThreads) {
        const int cnt = __ldg(p.ycnt + oy);
        uint32_t tot = 0u;
        for (int kp = 0; kp < ypairs; ++kp) {
            const uint32_t wa = 2 * kp < cnt ? aa_weight16(__ldg(p.yw + oy * p.ykmax + 2 * kp)) : 0u;
            const uint32_t wb = 2 * kp + 1 < cnt ? aa_weight16(__ldg(p.yw + oy * p.ykmax + 2 * kp + 1)) : 0u;
            s_yw16[oy * ypairs + kp] = wa | (wb << 16);
            tot += wa + wb;
        }
        s_yinv[oy] = 1.0f / static_cast<float>(tot);
    }
    for (int i = threadIdx.x; i < oh * p.ykmax; i += kPtThreads) s_yw[i] = __ldg(p.yw + i);
    for (int i = threadIdx.x; i < ow * p.xkmax; i += kPtThreads) s_xw[i] = __ldg(p.xw + i);
    for (int i = threadIdx.x; i < oh; i += kPtThreads) { s_ylo[i] = __ldg(p.ylo + i); s_ycnt[i] = __ldg(p.ycnt + i); }
    for (int i = threadIdx.x; i < ow; i += kPtThreads) { s_xlo[i] = __ldg(p.xlo + i); s_xcnt[i] = __ldg(p.xcnt + i); }
    auto raw_full = [&](int b) { return bar0 + 8u * b; };
    auto raw_empty = [&](int b) { return bar0 + 16u + 8u * b; };
    auto vb_full = [&](int b) { return bar0 + 32u + 8u * b; };
    auto vb_empty = [&](int b) { return bar0 + 48u + 8u * b; };
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform role index
    const int per_frame = p.gh * p.gw;
    const int step = gridDim.x;

    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int b = 0; b < 2; ++b) {
            mbar_init(raw_full(b), 1);
            mbar_init(raw_empty(b), kPtRowWarps);
            mbar_init(vb_full(b), kPtRowWarps);
            mbar_init(vb_empty(b), kPtColWarps);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == 0) {
        // =============================== scheduler + TMA producer ===============================
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += step, ++it) {
                const int buf = it & 1;
                if (it >= 2) mbar_wait_backoff(raw_empty(buf), static_cast<uint32_t>((it >> 1) - 1) & 1u);
                const int frame = tile / per_frame, rem = tile - frame * per_frame;
                const int py = rem / p.gw, px = rem - py * p.gw;
                const int ox0 = px * 16;
                const int x_lo = s_xlo[ox0];
                const int x_hi = s_xlo[ox0 + 15] + s_xcnt[ox0 + 15];
                const int b0 = (x_lo * 3) & ~15;
                const int y_first = s_ylo[py * 16];
                int* pr = params + (it & 3) * kPtParamInts;
                pr[0] = frame; pr[1] = py; pr[2] = px; pr[3] = b0; pr[4] = y_first; pr[5] = (x_hi * 3 - b0 + 15) >> 4;
                mbar_arrive_expect_tx(raw_full(buf), raw_bytes);
                for (int j = 0; j < nbox; ++j)
                    tma_load_3d(&tmap, raw_full(buf), base_u32 + buf * raw_bytes + j * rows_tile * 256, b0 + j * 256, y_first, frame);
            }
        }
    } else if (warp >= kPtRowWarp0) {
        // =============================== pass 1: vertical filter, warp = output row ===============================
        const int r = warp - kPtRowWarp0;
        const int rot = (lane >> 1) & 3;   // store-order rotation: the four 16-byte stores of a lane hit all bank groups evenly
        int it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += step, ++it) {
            const int buf = it & 1;
            const uint32_t ph = static_cast<uint32_t>(it >> 1) & 1u;
            mbar_wait(raw_full(buf), ph);
            const int* pr = params + (it & 3) * kPtParamInts;
            const int py = pr[1], y_first = pr[4], nvec = pr[5];
            if (it >= 2) mbar_wait(vb_empty(buf), ph ^ 1u);
            const uint8_t* raw = base + buf * raw_bytes;
            float* vrow = vb0 + buf * vb_floats + static_cast<size_t>(r) * p.sstride;
            const int oy = py * 16 + r;
            const int y0 = s_ylo[oy] - y_first, cnt = s_ycnt[oy];
            const uint32_t* wy16 = s_yw16 + oy * ypairs;
            const float inv = s_yinv[oy];
            const int npair = (cnt + 1) >> 1;
            for (int v = lane; v < nvec; v += 32) {
                const uint8_t* src = raw + (v >> 4) * (rows_tile * 256) + y0 * 256 + (v & 15) * 16;
                uint32_t iacc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) iacc[i] = 0u;
#pragma unroll 3
                for (int kp = 0; kp < npair; ++kp) {
                    const uint4 qa = *reinterpret_cast<const uint4*>(src + (2 * kp) * 256);
                    // an odd tap count pairs the last tap with weight 0: any in-bounds row will do
                    const uint4 qb = *reinterpret_cast<const uint4*>(src + (2 * kp + (2 * kp + 1 < cnt ? 1 : 0)) * 256);
                    vtaps2(qa, qb, wy16[kp], iacc);
                }
                float4 f[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    f[i] = make_float4(acc_to_float(iacc[4 * i]) * inv, acc_to_float(iacc[4 * i + 1]) * inv,
                                       acc_to_float(iacc[4 * i + 2]) * inv, acc_to_float(iacc[4 * i + 3]) * inv);
                // rotate the store order by `rot` (two select stages) -- lane l writes unit (i + rot) & 3 in store i
                float4 g[4], h[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 a = f[i], b = f[(i + 1) & 3];
                    g[i] = (rot & 1) ? b : a;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 a = g[i], b = g[(i + 2) & 3];
                    h[i] = (rot & 2) ? b : a;
                }
                float4* dst = reinterpret_cast<float4*>(vrow + v * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[(i + rot) & 3] = h[i];
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(raw_empty(buf));
                mbar_arrive(vb_full(buf));
            }
        }
    } else {
        // =============================== pass 2: horizontal filter + normalise + patchify ===============================
        // one thread per pixel of the 16 x 16 patch: the horizontal pass is a chain of shared-memory loads per tap, i.e. latency
        // bound per warp -- with 4 warps x 2 pixels it was the stage the 16 row warps ended up waiting for (ncu: 28 % of the
        // kernel's executed instructions were their mbarrier polls)
        const int pix = tid - kPtColWarp0 * 32;   // 0..255
        const int ky = pix >> 4, kx = pix & 15;
        int it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += step, ++it) {
            const int buf = it & 1;
            mbar_wait_backoff(vb_full(buf), static_cast<uint32_t>(it >> 1) & 1u);
            const int* pr = params + (it & 3) * kPtParamInts;
            const int frame = pr[0], py = pr[1], px = pr[2], b0 = pr[3];
            const float* vbuf = vb0 + buf * vb_floats;
            const size_t patch = (static_cast<size_t>(frame) * p.gh + py) * p.gw + px;
            {
                const int ox = px * 16 + kx;
                const int x0 = s_xlo[ox], cnt = s_xcnt[ox];
                const float* wx = s_xw + ox * p.xkmax;
                const float* src = vbuf + static_cast<size_t>(ky) * p.sstride + (x0 * 3 - b0);
                float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
#pragma unroll 4
                for (int k = 0; k < cnt; ++k) {
                    const float w = wx[k];
                    s0 = fmaf(w, src[3 * k], s0);
                    s1 = fmaf(w, src[3 * k + 1], s1);
                    s2 = fmaf(w, src[3 * k + 2], s2);
                }
                const float sm[3] = {s0, s1, s2};   // memory channel order
                __nv_bfloat16* o = p.out + patch * 768 + ky * 16 + kx;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int c = p.bgr ? 2 - j : j;  // output (RGB) channel of memory channel j
                    o[c * 256] = __float2bfloat16_rn(normalise_px(sm[j], p.mean[c], p.inv_std[c]));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(vb_empty(buf));
        }
    }
}

// =====================================================================================================================
// TMA-staged variant, second form (the default): same roles, same arithmetic, same bits -- laid out for the shared-memory pipe, which
// is what bounds the first form (ncu r02c: 78.6 % of the pipe's wavefronts, 30 % of them bank conflicts; issue slots 66 %).
//
//   * pass 1, one warp = TWO adjacent output rows.  Their tap windows overlap by ~55 % (window 2 x support, row pitch = scale), so the
//     raw rows of the union are read, and byte-interleaved, once for both (-25 % raw reads).  The 16-bit weight pairs of both rows are
//     tabulated per output-row pair on the union's pair grid (exact integer accumulation: the grouping of taps cannot change a bit).
//     A lane owns the 4-byte columns lane + 32 j: four conflict-free 128-byte LDS.32 per raw row instead of one 512-byte LDS.128.
//   * the intermediate is a row-PAIR matrix of float2 (row 2 rp, row 2 rp + 1) per byte column; a lane's four columns = two 16-byte
//     stores whose halves swap (in registers) for lanes with bit 2 set: every quarter-warp then covers all 32 banks.
//   * pass 2, one thread = one output column of one row pair: one LDS.64 per tap and channel feeds two FMA chains (the weight is
//     shared), so the horizontal pass issues half the loads; a half-warp = the 16 columns of ONE row pair.
//   * the two row-warp octets / column-warp quartets each own every other tile (= one buffer of each ring) and run concurrently.
// =====================================================================================================================
constexpr int kP2ParamInts = 8, kP2ParamSlots = 8;
// 16 row warps = 8 output-row pairs x 2 column halves of the tile; 4 column warps = 8 row pairs x 16 output columns (a second quartet
// on alternate tiles was slower, 1.95 vs 1.87 us per 1080p frame: the kernel is bound by issue slots, which extra pollers take)
constexpr int kP2RowWarps = 16, kP2ColWarps = 4;
constexpr int kP2Threads = (kP2RowWarps + kP2ColWarps + 1) * 32;
constexpr int kP2ColWarp0 = 1, kP2RowWarp0 = 1 + kP2ColWarps;
__device__ __forceinline__ uint32_t lds32(const uint8_t* p) { return *reinterpret_cast<const uint32_t*>(p); }

__global__ void __launch_bounds__(kP2Threads, 1)
preprocess_tma2_kernel(const __grid_constant__ CUtensorMap tmap, const PreParams p, int rows_tile, int nbox, int tiles, int upairs, int nst, int nvb) {
    extern __shared__ uint8_t smem_raw_[];
    const uint32_t base_u32 = (smem_u32(smem_raw_) + 127u) & ~127u;
    uint8_t* base = smem_raw_ + (base_u32 - smem_u32(smem_raw_));
    const int raw_bytes = nbox * rows_tile * 256;
    const int vb_bytes = 8 * p.sstride * 8;                       // 8 row pairs x sstride float2 columns
    // raw ring: nst (2 or 3) tiles deep; intermediate ring: nvb (2 or 3) -- with the two passes at comparable ~0.6 us per tile and a
    // hand-off latency of a few hundred cycles, a two-deep intermediate ring made BOTH sides wait for each other (ncu r02d)
    uint8_t* vb0 = base + nst * raw_bytes;
    int* params = reinterpret_cast<int*>(base + nst * raw_bytes + nvb * vb_bytes);             // [kP2ParamSlots][kP2ParamInts]
    const uint32_t bar0 = base_u32 + nst * raw_bytes + nvb * vb_bytes + kP2ParamSlots * kP2ParamInts * 4;
    const int oh = p.gh * 16, ow = p.gw * 16;
    float* s_xw = reinterpret_cast<float*>(base + nst * raw_bytes + nvb * vb_bytes + kP2ParamSlots * kP2ParamInts * 4 + 128);
    float* s_yinv = s_xw + ow * p.xkmax;
    int* s_ylo = reinterpret_cast<int*>(s_yinv + oh);
    int* s_xlo = s_ylo + oh;
    int* s_xcnt = s_xlo + ow;
    int* s_unp = s_xcnt + ow;                                     // union tap pairs per output-row pair
    uint2* s_uw = reinterpret_cast<uint2*>(s_unp + ((oh / 2 + 1) & ~1));   // [oh / 2][upairs]: (row a, row b) weight pairs on the union's grid
    // per output-row pair P = (rows 2P, 2P + 1): union window = rows [ylo[2P], max(end_a, end_b)); pair jp covers union rows 2 jp, 2 jp + 1
    for (int P = threadIdx.x; P < oh / 2; P += kP2Threads) {
        const int ya = __ldg(p.ylo + 2 * P), ca = __ldg(p.ycnt + 2 * P);
        const int yb = __ldg(p.ylo + 2 * P + 1), cb = __ldg(p.ycnt + 2 * P + 1);
        const int d = yb - ya;                                   // >= 0: the windows move down with the output row
        const int ulen = max(ca, d + cb);
        const int np = (ulen + 1) >> 1;
        // pairs [0, ja) carry weight of row a, pairs [jb, np) of row b: three branch-free tap loops (a only | both | b only)
        const int ja = (ca + 1) >> 1, jb = min(d >> 1, np);
        s_unp[P] = np | (ja << 8) | (jb << 16);
        uint32_t ta = 0u, tb = 0u;
        for (int jp = 0; jp < upairs; ++jp) {
            uint32_t w[4];                                       // a(2 jp), a(2 jp + 1), b(2 jp), b(2 jp + 1)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int u = 2 * jp + (e & 1);
                const int k = e < 2 ? u : u - d;
                const int cnt = e < 2 ? ca : cb;
                const float* wy = p.yw + static_cast<size_t>(2 * P + (e >> 1)) * p.ykmax;
                w[e] = (k >= 0 && k < cnt) ? aa_weight16(__ldg(wy + k)) : 0u;
            }
            s_uw[P * upairs + jp] = make_uint2(w[0] | (w[1] << 16), w[2] | (w[3] << 16));
            ta += w[0] + w[1];
            tb += w[2] + w[3];
        }
        s_yinv[2 * P] = 1.0f / static_cast<float>(ta);
        s_yinv[2 * P + 1] = 1.0f / static_cast<float>(tb);
    }
    for (int i = threadIdx.x; i < ow * p.xkmax; i += kP2Threads) s_xw[i] = __ldg(p.xw + i);
    for (int i = threadIdx.x; i < oh; i += kP2Threads) s_ylo[i] = __ldg(p.ylo + i);
    for (int i = threadIdx.x; i < ow; i += kP2Threads) { s_xlo[i] = __ldg(p.xlo + i); s_xcnt[i] = __ldg(p.xcnt + i); }
    auto raw_full = [&](int b) { return bar0 + 8u * b; };                // b < nst <= 4
    auto raw_empty = [&](int b) { return bar0 + 32u + 8u * b; };
    auto vb_full = [&](int b) { return bar0 + 64u + 8u * b; };                 // b < nvb <= 4
    auto vb_empty = [&](int b) { return bar0 + 96u + 8u * b; };
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform role index
    const int per_frame = p.gh * p.gw;
    const int step = gridDim.x;

    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int b = 0; b < nst; ++b) {
            mbar_init(raw_full(b), 1);
            mbar_init(raw_empty(b), kP2RowWarps);
        }
        for (int b = 0; b < nvb; ++b) {
            mbar_init(vb_full(b), kP2RowWarps);
            mbar_init(vb_empty(b), kP2ColWarps);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == 0) {
        // =============================== scheduler + TMA producer ===============================
        if (lane == 0) {
            int it = 0, buf = 0, use = 0;                 // ring slot of tile `it` and how often the slot has been used before
            for (int tile = blockIdx.x; tile < tiles; tile += step, ++it) {
                if (use >= 1) mbar_wait_backoff(raw_empty(buf), static_cast<uint32_t>(use - 1) & 1u, 40);
                const int frame = tile / per_frame, rem = tile - frame * per_frame;
                const int py = rem / p.gw, px = rem - py * p.gw;
                const int ox0 = px * 16;
                const int x_lo = s_xlo[ox0];
                const int x_hi = s_xlo[ox0 + 15] + s_xcnt[ox0 + 15];
                const int b0 = (x_lo * 3) & ~15;
                const int y_first = s_ylo[py * 16];
                int* pr = params + (it & (kP2ParamSlots - 1)) * kP2ParamInts;
                pr[0] = frame; pr[1] = py; pr[2] = px; pr[3] = b0; pr[4] = y_first; pr[5] = (x_hi * 3 - b0 + 15) >> 4;
                mbar_arrive_expect_tx(raw_full(buf), raw_bytes);
                for (int j = 0; j < nbox; ++j)
                    tma_load_3d(&tmap, raw_full(buf), base_u32 + buf * raw_bytes + j * rows_tile * 256, b0 + j * 256, y_first, frame);
                if (++buf == nst) { buf = 0; ++use; }
            }
        }
    } else if (warp >= kP2RowWarp0) {
        // =============================== pass 1: vertical filter, warp = one output-row pair x one column half ===============================
        const int rp = (warp - kP2RowWarp0) & 7, half = (warp - kP2RowWarp0) >> 3;
        const int sel = (lane >> 2) & 1;               // lanes whose two 16-byte stores go out in swapped order
        int it = 0, buf = 0, use = 0, vbuf = 0, vuse = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += step, ++it) {
            mbar_wait_backoff(raw_full(buf), static_cast<uint32_t>(use) & 1u, 20);
            const int* pr = params + (it & (kP2ParamSlots - 1)) * kP2ParamInts;
            const int py = pr[1], y_first = pr[4], ncol4 = pr[5] * 4;
            if (vuse >= 1) mbar_wait_backoff(vb_empty(vbuf), static_cast<uint32_t>(vuse - 1) & 1u, 100);
            const uint8_t* raw = base + buf * raw_bytes;
            float* vrow = reinterpret_cast<float*>(vb0 + vbuf * vb_bytes) + static_cast<size_t>(rp) * p.sstride * 2;
            const int P = py * 8 + rp;
            const uint2* uw = s_uw + P * upairs;
            const int unp = s_unp[P];
            const int np = unp & 0xff, ja = (unp >> 8) & 0xff, jb = unp >> 16;
            const int y0 = s_ylo[2 * P] - y_first;
            const float inv_a = s_yinv[2 * P], inv_b = s_yinv[2 * P + 1];
            // one pass = 512 byte columns of the tile; this warp's half of it = the 4-byte columns 64 half + lane + 32 j, j = 0, 1
            for (int c4_0 = 64 * half + lane; c4_0 < ncol4; c4_0 += 128) {
                uint32_t acc_a[8], acc_b[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { acc_a[i] = 0u; acc_b[i] = 0u; }
                // byte offset of the lane's column group j inside a raw row of the ring; a group beyond the staged boxes (only possible
                // beyond the tile's last column: never stored) reads group 0 instead of branching
                int off0, off1;
                {
                    const int cb0 = 4 * c4_0, cb1 = 4 * (c4_0 + 32);
                    off0 = (cb0 >> 8) * (rows_tile * 256) + (cb0 & 255);
                    off1 = cb1 < nbox * 256 ? (cb1 >> 8) * (rows_tile * 256) + (cb1 & 255) : off0;
                }
                const uint8_t* rowp = raw + y0 * 256;                       // union row 2 jp
                const int last = (rows_tile - 1 - y0) * 256;                // byte offset of the tile's last row from rowp(jp = 0)
                auto taps = [&](int jp, bool do_a, bool do_b) {
                    const uint2 w = uw[jp];
                    // an odd union length pairs its last row with weight 0: any in-bounds row will do
                    const int o0 = jp * 512, o1 = min(o0 + 256, last);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int off = j == 0 ? off0 : off1;
                        const uint32_t a = lds32(rowp + o0 + off), b = lds32(rowp + o1 + off);
                        const uint32_t lo = __byte_perm(a, b, 0x5140u);     // (a0, b0, a1, b1)
                        const uint32_t hi = __byte_perm(a, b, 0x7362u);     // (a2, b2, a3, b3)
                        if (do_a) {
                            acc_a[4 * j] = dp2a_lo(w.x, lo, acc_a[4 * j]);
                            acc_a[4 * j + 1] = dp2a_hi(w.x, lo, acc_a[4 * j + 1]);
                            acc_a[4 * j + 2] = dp2a_lo(w.x, hi, acc_a[4 * j + 2]);
                            acc_a[4 * j + 3] = dp2a_hi(w.x, hi, acc_a[4 * j + 3]);
                        }
                        if (do_b) {
                            acc_b[4 * j] = dp2a_lo(w.y, lo, acc_b[4 * j]);
                            acc_b[4 * j + 1] = dp2a_hi(w.y, lo, acc_b[4 * j + 1]);
                            acc_b[4 * j + 2] = dp2a_lo(w.y, hi, acc_b[4 * j + 2]);
                            acc_b[4 * j + 3] = dp2a_hi(w.y, hi, acc_b[4 * j + 3]);
                        }
                    }
                };
                int jp = 0;
                const int j_both0 = min(jb, ja);
                for (; jp < j_both0; ++jp) taps(jp, true, false);           // rows only row a reaches
#pragma unroll 2
                for (; jp < ja; ++jp) taps(jp, true, true);                 // the overlap of the two windows
                for (; jp < jb; ++jp) { }                                   // (windows that do not touch: nothing to add)
                for (; jp < np; ++jp) taps(jp, false, true);                // rows only row b reaches
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c4 = c4_0 + 32 * j;
                    if (c4 >= ncol4) continue;
                    float fa[4], fb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        fa[i] = acc_to_float(acc_a[4 * j + i]) * inv_a;
                        fb[i] = acc_to_float(acc_b[4 * j + i]) * inv_b;
                    }
                    // columns (0, 1) and (2, 3) as float2 pairs at their natural places; lanes with bit 2 set store them in the other order
                    const float4 q0 = make_float4(fa[0], fb[0], fa[1], fb[1]), q1 = make_float4(fa[2], fb[2], fa[3], fb[3]);
                    float4* dst = reinterpret_cast<float4*>(vrow + 8 * c4);          // 4 columns x float2 = 32 bytes
                    dst[sel] = sel ? q1 : q0;
                    dst[sel ^ 1] = sel ? q0 : q1;
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(raw_empty(buf));
                mbar_arrive(vb_full(vbuf));
            }
            if (++buf == nst) { buf = 0; ++use; }
            if (++vbuf == nvb) { vbuf = 0; ++vuse; }
        }
    } else {
        // =============================== pass 2: horizontal filter + normalise + patchify, thread = output column of a row pair ===============================
        const int wq = warp - kP2ColWarp0;
        const int kx = lane & 15, rp = 2 * wq + (lane >> 4);
        int it = 0, grp = 0, vuse = 0;                    // grp = ring slot of tile `it`
        for (int tile = blockIdx.x; tile < tiles; tile += step, ++it) {
            mbar_wait_backoff(vb_full(grp), static_cast<uint32_t>(vuse) & 1u, 40);
            const int* pr = params + (it & (kP2ParamSlots - 1)) * kP2ParamInts;
            const int frame = pr[0], py = pr[1], px = pr[2], b0 = pr[3];
            const size_t patch = (static_cast<size_t>(frame) * p.gh + py) * p.gw + px;
            {
                const int ox = px * 16 + kx;
                const int x0 = s_xlo[ox], cnt = s_xcnt[ox];
                const float* wx = s_xw + ox * p.xkmax;
                const float2* src = reinterpret_cast<const float2*>(vb0 + grp * vb_bytes) + static_cast<size_t>(rp) * p.sstride + (x0 * 3 - b0);
                float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
#pragma unroll 4
                for (int k = 0; k < cnt; ++k) {
                    const float w = wx[k];
                    const float2 v0 = src[3 * k], v1 = src[3 * k + 1], v2 = src[3 * k + 2];
                    a0 = fmaf(w, v0.x, a0); c0 = fmaf(w, v0.y, c0);
                    a1 = fmaf(w, v1.x, a1); c1 = fmaf(w, v1.y, c1);
                    a2 = fmaf(w, v2.x, a2); c2 = fmaf(w, v2.y, c2);
                }
                const float sa[3] = {a0, a1, a2}, sc[3] = {c0, c1, c2};   // memory channel order; rows 2 rp and 2 rp + 1
                __nv_bfloat16* o = p.out + patch * 768 + (2 * rp) * 16 + kx;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int c = p.bgr ? 2 - j : j;  // output (RGB) channel of memory channel j
                    o[c * 256] = __float2bfloat16_rn(normalise_px(sa[j], p.mean[c], p.inv_std[c]));
                    o[c * 256 + 16] = __float2bfloat16_rn(normalise_px(sc[j], p.mean[c], p.inv_std[c]));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(vb_empty(grp));
            if (++grp == nvb) { grp = 0; ++vuse; }
        }
    }
}

static int g_pre_tma = 2, g_pre_identity = 1;   // 0: direct-load kernel, 1: TMA-staged (first form), 2: TMA-staged, row pairs (default)
void set_preprocess_tma(int on) { g_pre_tma = on; }
void set_preprocess_identity(int on) { g_pre_identity = on; }

// returns 1 if the TMA variant was launched, 0 if the input does not qualify, negative on error
static int try_launch_preprocess_tma(const PreprocArgs& a, PreParams p, cudaStream_t stream) {
    if (!g_pre_tma || !p.vec) return 0;
    // tile extent: input rows / bytes one 16 x 16 output patch depends on
    const double sy = static_cast<double>(a.ty.in) / a.ty.out, sx = static_cast<double>(a.tx.in) / a.tx.out;
    const double supy = sy >= 1.0 ? sy : 1.0, supx = sx >= 1.0 ? sx : 1.0;
    int rows_tile = static_cast<int>(16 * sy + 2 * supy + 4);
    if (rows_tile > a.h) rows_tile = a.h;
    const int span_px = static_cast<int>(16 * sx + 2 * supx + 4);
    const int nvec_max = (span_px * 3 + 15 + 15) / 16;
    const int nbox = (nvec_max * 16 + 255) / 256;
    if (rows_tile > 256) return 0;
    // small tiles are bound by the per-tile hand-offs, not by bytes: this variant costs ~1.45 us + 0.16 us/MB per frame, the direct-load
    // kernel ~1 us/MB.  Measured cross-over (600 frames; direct vs TMA us/frame): 540x960 1.67 / 1.75, 720x1280 2.70 / 1.95,
    // 1080x1920 4.11 / 2.42 -> at about 12x fewer output than input pixels
    if (sy * sx < 12.0) return 0;
    p.sstride = nvec_max * 16 + 4;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles64 = static_cast<int64_t>(a.n) * a.gh * a.gw;
    if (tiles64 > 0x7fffffff) return 0;
    const int tiles = static_cast<int>(tiles64);
    const int grid = tiles < sms ? tiles : sms;
    CUtensorMap tmap;
    if (g_pre_tma >= 2 && (a.gh * 16) % 2 == 0) {
        // second form: row-pair intermediate (float2 per byte column), union-window weight table
        const int upairs = (a.ty.kmax + static_cast<int>(ceil(sy)) + 2) / 2;
        const int oh = a.gh * 16, ow = a.gw * 16;
        // input rows 16 output rows depend on: windows [int(c - sup + .5), int(c + sup + .5)), c = sy (i + .5)  ->  <= 15 sy + 2 sup + 2
        int rows2 = static_cast<int>(15 * sy + 2 * supy + 1) + 1;
        if (rows2 > a.h) rows2 = a.h;
        const size_t vb1 = static_cast<size_t>(8) * p.sstride * 8;
        const size_t fixed = kP2ParamSlots * kP2ParamInts * 4 + 128 + 128 + 128 +
                             (static_cast<size_t>(ow) * a.tx.kmax + oh + oh + 2 * ow + oh / 2 + 2) * 4 + static_cast<size_t>(oh / 2) * upairs * 8 + 16;
        const size_t raw1 = static_cast<size_t>(nbox) * rows2 * 256;
        // ring depths: the raw ring first (measured at 1080p, where only one of the two fits three deep: 1.87 vs 1.96 us per frame)
        const size_t cap = 227 * 1024;
        const int nst = fixed + 3 * raw1 + 2 * vb1 <= cap ? 3 : 2;
        const int nvb = fixed + nst * raw1 + 3 * vb1 <= cap ? 3 : 2;
        const size_t smem2 = fixed + nst * raw1 + nvb * vb1;
        if (smem2 <= 227 * 1024) {
            int rc = make_tmap_u8_3d(&tmap, a.frames, 3LL * a.w, a.h, a.n, a.row_pitch, a.frame_pitch, 256, rows2);
            if (rc) return rc;
            CRE_SMEM_ATTR_ONCE(preprocess_tma2_kernel, smem2);
            LaunchScope scope(CRE_K_PREPROCESS, static_cast<double>(a.n) * (3.0 * a.h * a.w + 1536.0 * a.gh * a.gw), stream);
            preprocess_tma2_kernel<<<grid, kP2Threads, smem2, stream>>>(tmap, p, rows2, nbox, tiles, upairs, nst, nvb);
            CRE_CUDA_OK(cudaGetLastError());
            return 1;
        }
    }
    const size_t smem = 2 * static_cast<size_t>(nbox) * rows_tile * 256 + 2 * static_cast<size_t>(16) * p.sstride * 4 +
                        4 * kPtParamInts * 4 + 64 + 128 +
                        (static_cast<size_t>(a.gh) * 16 * (a.ty.kmax + 2) + static_cast<size_t>(a.gw) * 16 * (a.tx.kmax + 2)) * 4 +
                        static_cast<size_t>(a.gh) * 16 * ((a.ty.kmax + 1) / 2 + 1) * 4;
    if (smem > 227 * 1024) return 0;
    int rc = make_tmap_u8_3d(&tmap, a.frames, 3LL * a.w, a.h, a.n, a.row_pitch, a.frame_pitch, 256, rows_tile);
    if (rc) return rc;
    CRE_SMEM_ATTR_ONCE(preprocess_tma_kernel, smem);
    LaunchScope scope(CRE_K_PREPROCESS, static_cast<double>(a.n) * (3.0 * a.h * a.w + 1536.0 * a.gh * a.gw), stream);
    preprocess_tma_kernel<<<grid, kPtThreads, smem, stream>>>(tmap, p, rows_tile, nbox, tiles);
    CRE_CUDA_OK(cudaGetLastError());
    return 1;
}

int launch_preprocess(const PreprocArgs& a, cudaStream_t stream) {
    CRE_REQUIRE(a.n > 0 && a.gh > 0 && a.gw > 0, "preprocess: empty problem");
    CRE_REQUIRE(a.gh * 16 <= a.ty.out && a.gw * 16 <= a.tx.out, "preprocess: patch grid exceeds the resized image");
    CRE_REQUIRE(a.row_pitch >= 3LL * a.w, "preprocess: row_pitch %lld < 3*w", (long long)a.row_pitch);
    PreParams p;
    p.frames = a.frames;
    p.frames_end = a.frames + static_cast<int64_t>(a.n - 1) * a.frame_pitch + static_cast<int64_t>(a.h - 1) * a.row_pitch + 3LL * a.w;
    p.h = a.h;
    p.w = a.w;
    p.row_pitch = a.row_pitch;
    p.frame_pitch = a.frame_pitch;
    p.bgr = a.bgr;
    p.gh = a.gh;
    p.gw = a.gw;
    for (int i = 0; i < 3; ++i) {
        p.mean[i] = a.mean[i];
        p.inv_std[i] = a.inv_std[i];
    }
    p.out = a.out;
    p.ylo = a.ty.lo;
    p.ycnt = a.ty.cnt;
    p.yw = a.ty.w;
    p.ykmax = a.ty.kmax;
    p.xlo = a.tx.lo;
    p.xcnt = a.tx.cnt;
    p.xw = a.tx.w;
    p.xkmax = a.tx.kmax;
    p.vec = ((reinterpret_cast<uintptr_t>(a.frames) & 15) == 0 && a.row_pitch % 16 == 0 && a.frame_pitch % 16 == 0) ? 1 : 0;
    p.rois = a.rois;
    if (a.rois == nullptr && g_pre_identity && a.ty.in == a.ty.out && a.tx.in == a.tx.out) {
        // no resize: normalise + patchify only
        const int64_t total = static_cast<int64_t>(a.n) * a.gh * a.gw * 16;
        LaunchScope scope(CRE_K_PREPROCESS, static_cast<double>(a.n) * (3.0 * a.h * a.w + 1536.0 * a.gh * a.gw), stream);
        preprocess_identity_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(p, total);
        CRE_CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (a.rois == nullptr) {
        const int rc = try_launch_preprocess_tma(a, p, stream);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    // widest band: (band output pixels) * scale + 2 * support, rounded up generously
    const double scale = static_cast<double>(a.tx.in) / a.tx.out;
    const double support = scale >= 1.0 ? scale : 1.0;
    const int span_px = static_cast<int>(kBandPatches * 16 * scale + 2 * support + 4);
    int sstride = ((span_px * 3 + 15 + 15) / 16) * 16 + 16 + (a.rois != nullptr ? 16 : 0);   // ROI: the crop origin shifts the alignment
    sstride += 4;  // de-phase the 16 rows across shared-memory banks
    p.sstride = sstride;
    const size_t smem = static_cast<size_t>(16) * sstride * sizeof(float);
    CRE_REQUIRE(smem <= 220 * 1024, "preprocess: input too wide for one band (%zu bytes of shared memory)", smem);
    CRE_SMEM_ATTR_ONCE(preprocess_kernel, smem);
    dim3 grid((a.gw + kBandPatches - 1) / kBandPatches, a.gh, a.rois != nullptr ? a.n_rois : a.n);
    // two balanced rounds of the vertical pass: 16 rows x nvec column groups over the block
    const int nvec_max = (span_px * 3 + 15 + 15) / 16;
    int threads = ((16 * nvec_max + 1) / 2 + 31) / 32 * 32;
    threads = threads < 128 ? 128 : (threads > kPreThreads ? kPreThreads : threads);
    LaunchScope scope(CRE_K_PREPROCESS, a.rois != nullptr ? static_cast<double>(a.n_rois) * 1536.0 * a.gh * a.gw
                                                          : static_cast<double>(a.n) * (3.0 * a.h * a.w + 1536.0 * a.gh * a.gw), stream);
    preprocess_kernel<<<grid, threads, smem, stream>>>(p);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cre
