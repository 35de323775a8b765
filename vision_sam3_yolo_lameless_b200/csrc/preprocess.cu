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
#include "internal.h"

namespace cre {

constexpr int kBandPatches = 2;
constexpr int kPreThreads = 512;

struct PreParams {
    const uint8_t* frames;
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
};

__global__ void __launch_bounds__(kPreThreads) preprocess_kernel(const PreParams p) {
    extern __shared__ __align__(16) float vbuf[];  // [16][sstride]
    const int band = blockIdx.x, py = blockIdx.y, frame = blockIdx.z;
    const int px0 = band * kBandPatches;
    const int npatch = min(kBandPatches, p.gw - px0);
    const int ox0 = px0 * 16, ox1 = ox0 + npatch * 16;  // output column range [ox0, ox1)
    const int x_lo = p.xlo[ox0];
    const int x_hi = p.xlo[ox1 - 1] + p.xcnt[ox1 - 1];
    const int b0 = p.vec ? ((x_lo * 3) & ~15) : x_lo * 3;  // first byte column staged
    const int nbytes = x_hi * 3 - b0;
    const int nvec = (nbytes + 15) >> 4;
    const int row_bytes = p.w * 3;
    const uint8_t* fbase = p.frames + static_cast<int64_t>(frame) * p.frame_pitch;

    // ---- pass 1: vertical filter ----
    for (int item = threadIdx.x; item < 16 * nvec; item += kPreThreads) {
        const int r = item / nvec, v = item - r * nvec;
        const int oy = py * 16 + r;
        const int y0 = p.ylo[oy], cnt = p.ycnt[oy];
        const float* wy = p.yw + static_cast<size_t>(oy) * p.ykmax;
        const int col = b0 + v * 16;
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
        const bool full = p.vec && (col + 16 <= row_bytes);
        for (int k = 0; k < cnt; ++k) {
            const float wk = __ldg(wy + k);
            const uint8_t* src = fbase + static_cast<int64_t>(y0 + k) * p.row_pitch + col;
            if (full) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(src));
                const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[4 * i + 0] = fmaf(wk, static_cast<float>(wds[i] & 0xffu), acc[4 * i + 0]);
                    acc[4 * i + 1] = fmaf(wk, static_cast<float>((wds[i] >> 8) & 0xffu), acc[4 * i + 1]);
                    acc[4 * i + 2] = fmaf(wk, static_cast<float>((wds[i] >> 16) & 0xffu), acc[4 * i + 2]);
                    acc[4 * i + 3] = fmaf(wk, static_cast<float>(wds[i] >> 24), acc[4 * i + 3]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (col + i < row_bytes) acc[i] = fmaf(wk, static_cast<float>(__ldg(src + i)), acc[i]);
            }
        }
        float4* dst = reinterpret_cast<float4*>(vbuf + static_cast<size_t>(r) * p.sstride + v * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    }
    __syncthreads();

    // ---- pass 2: horizontal filter + normalise + patchify ----
    const int nout = npatch * 768;
    for (int t = threadIdx.x; t < nout; t += kPreThreads) {
        const int pl = t / 768, kidx = t - pl * 768;
        const int c = kidx >> 8, ky = (kidx >> 4) & 15, kx = kidx & 15;
        const int ox = (px0 + pl) * 16 + kx;
        const int x0 = __ldg(p.xlo + ox), cnt = __ldg(p.xcnt + ox);
        const float* wx = p.xw + static_cast<size_t>(ox) * p.xkmax;
        const int cin = p.bgr ? 2 - c : c;
        const float* src = vbuf + static_cast<size_t>(ky) * p.sstride + (x0 * 3 + cin - b0);
        float s = 0.0f;
        for (int k = 0; k < cnt; ++k) s = fmaf(__ldg(wx + k), src[3 * k], s);
        const float val = (s * (1.0f / 255.0f) - p.mean[c]) * p.inv_std[c];
        const size_t patch = (static_cast<size_t>(frame) * p.gh + py) * p.gw + px0 + pl;
        p.out[patch * 768 + kidx] = __float2bfloat16_rn(val);
    }
}

int launch_preprocess(const PreprocArgs& a, cudaStream_t stream) {
    CRE_REQUIRE(a.n > 0 && a.gh > 0 && a.gw > 0, "preprocess: empty problem");
    CRE_REQUIRE(a.gh * 16 <= a.ty.out && a.gw * 16 <= a.tx.out, "preprocess: patch grid exceeds the resized image");
    CRE_REQUIRE(a.row_pitch >= 3LL * a.w, "preprocess: row_pitch %lld < 3*w", (long long)a.row_pitch);
    PreParams p;
    p.frames = a.frames;
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
    // widest band: (band output pixels) * scale + 2 * support, rounded up generously
    const double scale = static_cast<double>(a.tx.in) / a.tx.out;
    const double support = scale >= 1.0 ? scale : 1.0;
    const int span_px = static_cast<int>(kBandPatches * 16 * scale + 2 * support + 4);
    int sstride = ((span_px * 3 + 15 + 15) / 16) * 16 + 16;
    sstride += 4;  // de-phase the 16 rows across shared-memory banks
    p.sstride = sstride;
    const size_t smem = static_cast<size_t>(16) * sstride * sizeof(float);
    CRE_REQUIRE(smem <= 220 * 1024, "preprocess: input too wide for one band (%zu bytes of shared memory)", smem);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        CRE_CUDA_OK(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        smem_set = smem;
    }
    dim3 grid((a.gw + kBandPatches - 1) / kBandPatches, a.gh, a.n);
    LaunchScope scope(CRE_K_PREPROCESS, static_cast<double>(a.n) * (3.0 * a.h * a.w + 1536.0 * a.gh * a.gw), stream);
    preprocess_kernel<<<grid, kPreThreads, smem, stream>>>(p);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cre
