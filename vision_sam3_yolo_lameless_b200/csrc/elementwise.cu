// Memory-bound row kernels around the GEMMs: LayerNorm, prefix-token fill, final norm + token mean,
// clip pooling + L2 normalisation, query hi/lo split, top-k list merge, gallery row update.
// All are one-warp-per-row (or one-CTA-per-row) with 16-byte loads and shuffle reductions.
#include <math.h>

#include "common.cuh"
#include "internal.h"

namespace cre {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (HF:modeling_dinov3_vit.py:411,416 nn.LayerNorm, eps 1e-5): fp32 in, bf16 out.
// One warp per row, the row lives in registers (DIM/128 float4 per lane), exact two-pass variance.
// ---------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                             const float* __restrict__ b, int rows, float eps,
                                                             __nv_bfloat16* __restrict__ out) {
    constexpr int V = DIM / 128;  // float4 per lane
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * DIM);
    float4 v[V];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        v[i] = xr[lane + 32 * i];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / DIM);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
        q += (a * a + c * c) + (d * d + e * e);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / DIM) + eps);
    uint2* o = reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * DIM);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + lane + 32 * i);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
        o[lane + 32 * i] = make_uint2(pack_bf16x2((v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y),
                                      pack_bf16x2((v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w));
    }
}

int launch_layernorm_bf16(const float* x, const float* g, const float* b, int rows, int dim, float eps,
                          __nv_bfloat16* out, cudaStream_t stream) {
    CRE_REQUIRE(rows > 0, "layernorm: no rows");
    const int grid = (rows + 7) / 8;
    LaunchScope scope(CRE_K_LAYERNORM, 6.0 * rows * dim, stream);
    if (dim == 768) layernorm_bf16_kernel<768><<<grid, 256, 0, stream>>>(x, g, b, rows, eps, out);
    else if (dim == 1024) layernorm_bf16_kernel<1024><<<grid, 256, 0, stream>>>(x, g, b, rows, eps, out);
    else {
        set_error("layernorm: unsupported dim %d (768 or 1024)", dim);
        return -3;
    }
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm folding.  LN(x) W^T = rstd * ((x - pivot) W'^T - (mean - pivot) c1) + c2 with
// W' = W * gamma (per input column), c1[n] = sum_k W'[n, k], c2[n] = bias[n] + sum_k beta[k] W[n, k]: the GEMM reads the
// bf16 residual stream (minus a per-row pivot ~ the row mean, so bf16 rounding is relative to the centred value exactly as
// for a LayerNorm output) and its epilogue applies the row statistics.  Statistics row layout (fp32):
// [pivot, -, -, -, (mean_i, M2_i) for every 128-column slot i].
//
// row_stats_kernel seeds the chain after the patch embedding: exact two-pass mean / M2 of each row, pivot = mean.
// ---------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ x, int rows, int stride,
                                                        __nv_bfloat16* __restrict__ xb, __nv_bfloat16* __restrict__ xl,
                                                        float* __restrict__ stats) {
    constexpr int V = DIM / 128;
    constexpr int S = DIM / 128;   // statistics slots
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * DIM);
    float4 v[V];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        v[i] = xr[lane + 32 * i];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / DIM);
    float q = 0.0f;
    uint2* o = reinterpret_cast<uint2*>(xb + static_cast<size_t>(row) * DIM);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
        q += (a * a + c * c) + (d * d + e * e);
        const uint32_t h0 = pack_bf16x2(a, c), h1 = pack_bf16x2(d, e);
        o[lane + 32 * i] = make_uint2(h0, h1);
        if (xl != nullptr)    // split residual stream: the low half, bf16((x - mean) - hi)
            reinterpret_cast<uint2*>(xl + static_cast<size_t>(row) * DIM)[lane + 32 * i] =
                make_uint2(pack_bf16x2(a - __uint_as_float(h0 << 16), c - __uint_as_float(h0 & 0xffff0000u)),
                           pack_bf16x2(d - __uint_as_float(h1 << 16), e - __uint_as_float(h1 & 0xffff0000u)));
    }
    const float m2 = warp_sum(q);
    float* sr = stats + static_cast<size_t>(row) * stride;
    if (lane == 0) sr[0] = mean;
    if (lane < S) *reinterpret_cast<float2*>(sr + 4 + 2 * lane) = make_float2(mean, m2 * (1.0f / S));
}

int launch_row_stats(const float* x, int rows, int dim, int stride, __nv_bfloat16* xb, __nv_bfloat16* xl, float* stats,
                     cudaStream_t stream) {
    CRE_REQUIRE(rows > 0, "row_stats: no rows");
    CRE_REQUIRE(stride >= 2 * (dim / 128) + 4 && stride % 2 == 0, "row_stats: statistics stride %d too small for dim %d", stride, dim);
    const int grid = (rows + 7) / 8;
    LaunchScope scope(CRE_K_ROW_STATS, (xl != nullptr ? 8.0 : 6.0) * rows * dim, stream);
    if (dim == 768) row_stats_kernel<768><<<grid, 256, 0, stream>>>(x, rows, stride, xb, xl, stats);
    else if (dim == 1024) row_stats_kernel<1024><<<grid, 256, 0, stream>>>(x, rows, stride, xb, xl, stats);
    else {
        set_error("row_stats: unsupported dim %d (768 or 1024)", dim);
        return -3;
    }
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

// One warp per output row n of a Linear that follows a LayerNorm: W' = bf16(W * gamma), c1 = row sum of the ROUNDED W'
// (what the tensor core will multiply), c2 = bias + W beta.  Rows [0, scaled_rows) are additionally multiplied by row_scale
// (the attention's head_dim^-0.5 on the q rows; a power of two, so exact).  Runs once per weight matrix at cre_create.
__global__ void __launch_bounds__(256) fold_ln_weights_kernel(const __nv_bfloat16* __restrict__ w, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ bias,
                                                              int n_rows, int k, int scaled_rows, float row_scale,
                                                              __nv_bfloat16* __restrict__ wf,
                                                              float* __restrict__ c1, float* __restrict__ c2) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const __nv_bfloat16* wr = w + static_cast<size_t>(n) * k;
    __nv_bfloat16* wo = wf + static_cast<size_t>(n) * k;
    const float rs = n < scaled_rows ? row_scale : 1.0f;
    float s1 = 0.0f, s2 = 0.0f;
    for (int j = lane; j < k; j += 32) {
        const float wv = __bfloat162float(wr[j]);
        const __nv_bfloat16 f = __float2bfloat16_rn(wv * gamma[j] * rs);
        wo[j] = f;
        s1 += __bfloat162float(f);
        s2 = fmaf(beta[j], wv, s2);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
        c1[n] = s1;
        c2[n] = ((bias != nullptr ? bias[n] : 0.0f) + s2) * rs;
    }
}

int launch_fold_ln_weights(const __nv_bfloat16* w, const float* gamma, const float* beta, const float* bias, int n_rows, int k,
                           int scaled_rows, float row_scale, __nv_bfloat16* wf, float* c1, float* c2, cudaStream_t stream) {
    CRE_REQUIRE(n_rows > 0 && k > 0, "fold_ln_weights: empty matrix");
    LaunchScope scope(CRE_K_FOLD_LN, 4.0 * n_rows * k, stream);
    fold_ln_weights_kernel<<<(n_rows + 7) / 8, 256, 0, stream>>>(w, gamma, beta, bias, n_rows, k, scaled_rows, row_scale, wf, c1, c2);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Final LayerNorm + mean over all tokens of a frame (HF:modeling_dinov3_vit.py:547 self.norm, then
// services/dinov3-pipeline/app/main.py:113 last_hidden_state.mean(dim=1)).  One CTA per frame,
// warps stride over the frame's tokens and keep a per-warp partial sum in registers.
// ---------------------------------------------------------------------------------------------
// SPLIT: the residual stream arrives as its two bf16 halves + the pivot in slot 0 of the row's statistics (EPI_RESID_SP):
// x = pivot + (hi + lo); the same 4 bytes per element as the fp32 stream.
template <int DIM, bool SPLIT>
__global__ void __launch_bounds__(256) final_norm_mean_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ xh,
                                                              const __nv_bfloat16* __restrict__ xl, const float* __restrict__ stats,
                                                              int stats_stride, const float* __restrict__ g,
                                                              const float* __restrict__ b, int t, float eps,
                                                              float* __restrict__ frame_emb,
                                                              float* __restrict__ tokens_out) {
    constexpr int V = DIM / 128;
    __shared__ float4 part[8][DIM / 4];
    const int frame = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int tok = warp; tok < t; tok += 8) {
        const size_t row = static_cast<size_t>(frame) * t + tok;
        float4 v[V];
        float s = 0.0f;
        if constexpr (SPLIT) {
            const uint2* hr = reinterpret_cast<const uint2*>(xh + row * DIM);
            const uint2* lr = reinterpret_cast<const uint2*>(xl + row * DIM);
            const float pivot = __ldg(stats + row * stats_stride);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const uint2 h = hr[lane + 32 * i], l = lr[lane + 32 * i];
                v[i].x = pivot + (__uint_as_float(h.x << 16) + __uint_as_float(l.x << 16));
                v[i].y = pivot + (__uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u));
                v[i].z = pivot + (__uint_as_float(h.y << 16) + __uint_as_float(l.y << 16));
                v[i].w = pivot + (__uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u));
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        } else {
            const float4* xr = reinterpret_cast<const float4*>(x + row * DIM);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                v[i] = xr[lane + 32 * i];
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
        const float mean = warp_sum(s) * (1.0f / DIM);
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
            q += (a * a + c * c) + (d * d + e * e);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / DIM) + eps);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + lane + 32 * i);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
            float4 y;
            y.x = (v[i].x - mean) * rstd * gg.x + bb.x;
            y.y = (v[i].y - mean) * rstd * gg.y + bb.y;
            y.z = (v[i].z - mean) * rstd * gg.z + bb.z;
            y.w = (v[i].w - mean) * rstd * gg.w + bb.w;
            if (tokens_out != nullptr) reinterpret_cast<float4*>(tokens_out + row * DIM)[lane + 32 * i] = y;
            acc[i].x += y.x; acc[i].y += y.y; acc[i].z += y.z; acc[i].w += y.w;
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) part[warp][lane + 32 * i] = acc[i];
    __syncthreads();
    const float inv_t = 1.0f / static_cast<float>(t);
    for (int c = threadIdx.x; c < DIM / 4; c += blockDim.x) {
        float4 s = part[0][c];
#pragma unroll
        for (int w = 1; w < 8; ++w) {
            const float4 o = part[w][c];
            s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
        }
        reinterpret_cast<float4*>(frame_emb + static_cast<size_t>(frame) * DIM)[c] =
            make_float4(s.x * inv_t, s.y * inv_t, s.z * inv_t, s.w * inv_t);
    }
}

int launch_final_norm_mean(const float* x, const __nv_bfloat16* xh, const __nv_bfloat16* xl, const float* stats, int stats_stride,
                           const float* g, const float* b, int frames, int t, int dim, float eps, float* frame_emb, float* tokens_out,
                           cudaStream_t stream) {
    CRE_REQUIRE(frames > 0 && t > 0, "final_norm_mean: empty input");
    const bool split = x == nullptr;
    CRE_REQUIRE(!split || (xh != nullptr && xl != nullptr && stats != nullptr), "final_norm_mean: neither x nor its split halves given");
    LaunchScope scope(CRE_K_FINAL_NORM_MEAN, 4.0 * frames * t * dim * (tokens_out != nullptr ? 2.0 : 1.0), stream);
    if (dim == 768 && split) final_norm_mean_kernel<768, true><<<frames, 256, 0, stream>>>(x, xh, xl, stats, stats_stride, g, b, t, eps, frame_emb, tokens_out);
    else if (dim == 768) final_norm_mean_kernel<768, false><<<frames, 256, 0, stream>>>(x, xh, xl, stats, stats_stride, g, b, t, eps, frame_emb, tokens_out);
    else if (dim == 1024 && split) final_norm_mean_kernel<1024, true><<<frames, 256, 0, stream>>>(x, xh, xl, stats, stats_stride, g, b, t, eps, frame_emb, tokens_out);
    else if (dim == 1024) final_norm_mean_kernel<1024, false><<<frames, 256, 0, stream>>>(x, xh, xl, stats, stats_stride, g, b, t, eps, frame_emb, tokens_out);
    else {
        set_error("final_norm_mean: unsupported dim %d", dim);
        return -3;
    }
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

// cls + register tokens copied into rows [0, prefix) of every frame (HF:modeling_dinov3_vit.py:88-90)
__global__ void fill_prefix_kernel(float* __restrict__ x, const float* __restrict__ prefix, int t, int prefix_tokens,
                                   int dim4, int64_t total4) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c = static_cast<int>(i % dim4);
    const int64_t r = i / dim4;
    const int tok = static_cast<int>(r % prefix_tokens);
    const int64_t frame = r / prefix_tokens;
    reinterpret_cast<float4*>(x)[(frame * t + tok) * dim4 + c] = __ldg(reinterpret_cast<const float4*>(prefix) + tok * dim4 + c);
}

int launch_fill_prefix(float* x, const float* prefix, int frames, int t, int prefix_tokens, int dim,
                       cudaStream_t stream) {
    const int64_t total4 = static_cast<int64_t>(frames) * prefix_tokens * (dim / 4);
    LaunchScope scope(CRE_K_FILL_PREFIX, 16.0 * total4, stream);
    fill_prefix_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, stream>>>(x, prefix, t, prefix_tokens,
                                                                                         dim / 4, total4);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Clip pooling (services/dinov3-pipeline/app/main.py:204-208 np.mean over the clip's frame
// embeddings) + L2 normalisation (services/tracking-service/app/reid/matcher.py:124, +1e-8).
// One CTA per clip.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pool_clips_kernel(const float* __restrict__ emb, const int32_t* __restrict__ offs,
                                                          int dim, float* __restrict__ out_mean,
                                                          float* __restrict__ out_unit) {
    // thread = (frame lane fl, float4 column group cg): dim / 4 <= 256 column groups, 1024 / (dim / 4) frame lanes that stride
    // over the clip's frames with independent 16-byte loads; the frame lanes are then summed IN A FIXED ORDER through shared
    // memory (bit-reproducible), and the first dim / 4 threads finish the mean, the norm and both stores.
    __shared__ float4 part[1024];
    __shared__ float red[32];
    const int clip = blockIdx.x;
    const int f0 = offs[clip], f1 = offs[clip + 1];
    const int groups = dim >> 2;                     // float4 column groups (dim % 4 == 0)
    const int lanes = 1024 / groups;                 // frame lanes
    const int cg = threadIdx.x % groups, fl = threadIdx.x / groups;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (fl < lanes) {
        for (int f = f0 + fl; f < f1; f += lanes) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(emb + static_cast<size_t>(f) * dim) + cg);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    const float inv = f1 > f0 ? 1.0f / static_cast<float>(f1 - f0) : 0.0f;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    float ss = 0.0f;
    if (threadIdx.x < groups) {
        for (int l = 0; l < lanes; ++l) {
            const float4 v = part[l * groups + threadIdx.x];
            m.x += v.x; m.y += v.y; m.z += v.z; m.w += v.w;
        }
        m.x *= inv; m.y *= inv; m.z *= inv; m.w *= inv;
        ss = m.x * m.x + m.y * m.y + m.z * m.z + m.w * m.w;
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < groups) {
        float tot = 0.0f;
        for (int w = 0; w < (groups + 31) / 32; ++w) tot += red[w];
        const float scale = 1.0f / (sqrtf(tot) + 1e-8f);
        if (out_mean != nullptr) reinterpret_cast<float4*>(out_mean + static_cast<size_t>(clip) * dim)[threadIdx.x] = m;
        if (out_unit != nullptr)
            reinterpret_cast<float4*>(out_unit + static_cast<size_t>(clip) * dim)[threadIdx.x] =
                make_float4(m.x * scale, m.y * scale, m.z * scale, m.w * scale);
    }
}

int launch_pool_clips(const float* frame_emb, const int32_t* offs, int clips, int dim, float* out_mean,
                      float* out_unit, cudaStream_t stream) {
    CRE_REQUIRE(clips > 0, "pool_clips: no clips");
    CRE_REQUIRE(dim > 0 && dim <= 1024 && dim % 4 == 0, "pool_clips: dim %d out of range (multiple of 4, <= 1024)", dim);
    CRE_REQUIRE((reinterpret_cast<uintptr_t>(frame_emb) & 15) == 0 && (out_mean == nullptr || (reinterpret_cast<uintptr_t>(out_mean) & 15) == 0) &&
                    (out_unit == nullptr || (reinterpret_cast<uintptr_t>(out_unit) & 15) == 0), "pool_clips: pointers must be 16-byte aligned");
    LaunchScope scope(CRE_K_POOL_CLIPS, 0.0, stream);   // bytes depend on the device-side offsets: the caller knows them
    pool_clips_kernel<<<clips, 1024, 0, stream>>>(frame_emb, offs, dim, out_mean, out_unit);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

// query f32 [rows, dim] -> bf16 [rows, 2*dim] = [hi | lo], hi = bf16(q), lo = bf16(q - hi): the gallery GEMM
// then accumulates hi.g + lo.g in fp32, i.e. an (almost) fp32 query against the bf16 gallery.
__global__ void split_hi_lo_kernel(const float* __restrict__ q, int dim, int64_t total, __nv_bfloat16* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t r = i / dim;
    const int c = static_cast<int>(i % dim);
    const float v = q[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    out[r * 2 * dim + c] = hi;
    out[r * 2 * dim + dim + c] = lo;
}

int launch_split_hi_lo(const float* q, int rows, int dim, __nv_bfloat16* out, cudaStream_t stream) {
    const int64_t total = static_cast<int64_t>(rows) * dim;
    LaunchScope scope(CRE_K_SPLIT_HI_LO, 8.0 * total, stream);
    split_hi_lo_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(q, dim, total, out);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void fill_topk_kernel(float* __restrict__ s, int32_t* __restrict__ idx, int64_t count) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) {
        s[i] = -INFINITY;
        idx[i] = 0x7fffffff;
    }
}

int launch_fill_topk(float* scores, int32_t* idx, int64_t count, cudaStream_t stream) {
    LaunchScope scope(CRE_K_FILL_TOPK, 8.0 * count, stream);
    fill_topk_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, stream>>>(scores, idx, count);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

// gallery_topk prologue in ONE launch: the hi / lo split of the queries and the -inf fill of the partial lists (the scan is a
// ~30 us kernel, so every extra launch in front of it shows)
__global__ void topk_prepare_kernel(const float* __restrict__ q, int dim, int64_t total, __nv_bfloat16* __restrict__ out,
                                    float* __restrict__ s, int32_t* __restrict__ idx, int64_t count) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < total) {
        const int64_t r = i / dim;
        const int c = static_cast<int>(i - r * dim);
        const float v = q[i];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        out[r * 2 * dim + c] = hi;
        out[r * 2 * dim + dim + c] = lo;
    }
    if (i < count) {
        s[i] = -INFINITY;
        idx[i] = 0x7fffffff;
    }
}

int launch_topk_prepare(const float* q, int rows, int dim, __nv_bfloat16* out, float* scores, int32_t* idx, int64_t count,
                        cudaStream_t stream) {
    const int64_t total = static_cast<int64_t>(rows) * dim;
    const int64_t n = total > count ? total : count;
    LaunchScope scope(CRE_K_SPLIT_HI_LO, 8.0 * total + 8.0 * count, stream);
    topk_prepare_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(q, dim, total, out, scores, idx, count);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Merge candidate lists under the total order (score desc, index asc).  One warp per query: each
// lane scans a strided share of the lists*per_list candidates into a private sorted top-k, then
// k rounds of a warp arg-best over the lane heads pop the global winners.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool better(float s, int i, float s2, int i2) { return s > s2 || (s == s2 && i < i2); }

// Block-wide (256 threads) selection of the k best of `total` <= 256 * kSelNC candidates under (score desc, index asc); equal
// (score, index) pairs are taken in candidate order, so duplicates come out as often as they went in.  Every thread holds
// kSelNC candidates in registers (all loads independent: one memory latency), then k rounds of thread-best -> warp arg-best ->
// 8-entry shared-memory arg-best; the owner of the winner retires it.  red_* : 8 entries each.  Thread 0 writes the result.
constexpr int kSelNC = 16;
template <typename LoadS, typename LoadI>
__device__ __forceinline__ void block_select_topk(int total, int k, LoadS load_s, LoadI load_i, float* out_s, int32_t* out_i,
                                                  float* red_s, int* red_i, int* red_p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float cs[kSelNC];
    int ci[kSelNC];
#pragma unroll
    for (int j = 0; j < kSelNC; ++j) {
        const int c = threadIdx.x + 256 * j;
        cs[j] = -INFINITY;
        ci[j] = 0x7fffffff;
        if (c < total) { cs[j] = load_s(c); ci[j] = load_i(c); }
    }
    for (int r = 0; r < k; ++r) {
        float bs = cs[0];
        int bi = ci[0], bp = threadIdx.x;
#pragma unroll
        for (int j = 1; j < kSelNC; ++j)          // positions ascend with j: strict comparison keeps the earlier one on equality
            if (better(cs[j], ci[j], bs, bi)) { bs = cs[j]; bi = ci[j]; bp = threadIdx.x + 256 * j; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (better(os, oi, bs, bi) || (os == bs && oi == bi && op < bp)) { bs = os; bi = oi; bp = op; }
        }
        if (lane == 0) { red_s[warp] = bs; red_i[warp] = bi; red_p[warp] = bp; }
        __syncthreads();
        float ws = red_s[0];
        int wi = red_i[0], wp = red_p[0];
#pragma unroll
        for (int w = 1; w < 8; ++w)
            if (better(red_s[w], red_i[w], ws, wi) || (red_s[w] == ws && red_i[w] == wi && red_p[w] < wp)) {
                ws = red_s[w]; wi = red_i[w]; wp = red_p[w];
            }
        if (threadIdx.x == 0) { out_s[r] = ws; out_i[r] = wi; }
        if ((wp & 255) == static_cast<int>(threadIdx.x)) {
#pragma unroll
            for (int j = 0; j < kSelNC; ++j)
                if (j == (wp >> 8)) { cs[j] = -INFINITY; ci[j] = 0x7fffffff; }
        }
        __syncthreads();
    }
}

// one CTA per query (merge_topk_kernel below keeps the one-warp-per-query form for candidate counts beyond 256 * kSelNC)
__global__ void __launch_bounds__(256) merge_topk_block_kernel(const float* __restrict__ scores, const int32_t* __restrict__ idx,
                                                               int64_t list_stride, int64_t query_stride, int lists, int per_list,
                                                               int k, float* __restrict__ out_scores, int32_t* __restrict__ out_idx,
                                                               int out_stride) {
    __shared__ float red_s[8];
    __shared__ int red_i[8], red_p[8];
    const int query = blockIdx.x;
    auto off = [&](int c) { const int l = c / per_list; return l * list_stride + query * query_stride + (c - l * per_list); };
    block_select_topk(lists * per_list, k, [&](int c) { return __ldg(scores + off(c)); }, [&](int c) { return __ldg(idx + off(c)); },
                      out_scores + static_cast<size_t>(query) * out_stride, out_idx + static_cast<size_t>(query) * out_stride, red_s, red_i,
                      red_p);
}

__global__ void __launch_bounds__(128) merge_topk_kernel(const float* __restrict__ scores, const int32_t* __restrict__ idx,
                                                         int64_t list_stride, int64_t query_stride, int lists,
                                                         int per_list, int q, int k, float* __restrict__ out_scores,
                                                         int32_t* __restrict__ out_idx, int out_stride) {
    const int query = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (query >= q) return;
    const int lane = threadIdx.x & 31;
    float bs[CRE_TOPK_MAX];
    int bi[CRE_TOPK_MAX];
#pragma unroll
    for (int j = 0; j < CRE_TOPK_MAX; ++j) { bs[j] = -INFINITY; bi[j] = 0x7fffffff; }
    const int total = lists * per_list;
    // candidates in batches of kBatch per lane: all loads of a batch are issued before the (dependent) insertions, so the
    // kernel pays one memory latency per batch instead of one per candidate (it is latency bound: a few KB per query)
    constexpr int kBatch = 8;
    for (int c0 = lane; c0 < total; c0 += 32 * kBatch) {
        float cs_[kBatch];
        int ci_[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const int c = c0 + 32 * b;
            cs_[b] = -INFINITY;
            ci_[b] = 0x7fffffff;
            if (c < total) {
                const int l = c / per_list, e = c - l * per_list;
                const int64_t o = l * list_stride + query * query_stride + e;
                cs_[b] = __ldg(scores + o);
                ci_[b] = __ldg(idx + o);
            }
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            float cs = cs_[b];
            int ci = ci_[b];
            if (better(cs, ci, bs[CRE_TOPK_MAX - 1], bi[CRE_TOPK_MAX - 1])) {
#pragma unroll
                for (int j = 0; j < CRE_TOPK_MAX; ++j) {
                    if (better(cs, ci, bs[j], bi[j])) {
                        const float ts = bs[j]; const int ti = bi[j];
                        bs[j] = cs; bi[j] = ci;
                        cs = ts; ci = ti;
                    }
                }
            }
        }
    }
    for (int r = 0; r < k; ++r) {
        float ws = bs[0];
        int wi = bi[0];
        int wl = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, ws, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
            if (better(os, oi, ws, wi) || (os == ws && oi == wi && ol < wl)) { ws = os; wi = oi; wl = ol; }
        }
        if (lane == 0) {
            out_scores[static_cast<size_t>(query) * out_stride + r] = ws;
            out_idx[static_cast<size_t>(query) * out_stride + r] = wi;
        }
        if (lane == wl) {  // pop the winner from its lane
#pragma unroll
            for (int j = 0; j < CRE_TOPK_MAX - 1; ++j) { bs[j] = bs[j + 1]; bi[j] = bi[j + 1]; }
            bs[CRE_TOPK_MAX - 1] = -INFINITY;
            bi[CRE_TOPK_MAX - 1] = 0x7fffffff;
        }
    }
}

int launch_merge_topk(const float* scores, const int32_t* idx, int64_t list_stride, int64_t query_stride,
                      int lists, int per_list, int q, int k, float* out_scores, int32_t* out_idx, int out_stride,
                      cudaStream_t stream) {
    CRE_REQUIRE(q > 0 && lists > 0, "merge_topk: empty input");
    CRE_REQUIRE(k >= 1 && k <= CRE_TOPK_LIMIT && per_list >= 1 && per_list <= CRE_TOPK_LIMIT && out_stride >= k,
                "merge_topk: k=%d per_list=%d out of range (1..%d)", k, per_list, CRE_TOPK_LIMIT);
    const bool block_form = static_cast<int64_t>(lists) * per_list <= 256 * kSelNC;
    if (!block_form && (k > CRE_TOPK_MAX || per_list > CRE_TOPK_MAX)) {
        set_error("merge_topk: %d lists x %d candidates with k=%d: more than %d candidates per query need k, per_list <= %d", lists, per_list, k,
                  256 * kSelNC, CRE_TOPK_MAX);
        return -3;
    }
    LaunchScope scope(CRE_K_MERGE_TOPK, 8.0 * lists * per_list * q, stream);
    if (block_form)
        merge_topk_block_kernel<<<q, 256, 0, stream>>>(scores, idx, list_stride, query_stride, lists, per_list, k, out_scores, out_idx,
                                                       out_stride);
    else
        merge_topk_kernel<<<(q + 3) / 4, 128, 0, stream>>>(scores, idx, list_stride, query_stride, lists, per_list, q, k,
                                                           out_scores, out_idx, out_stride);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K4, serving form: Q <= 2 queries against the whole gallery shard.  One pipeline.dinov3 message carries ONE clip embedding
// (services/tracking-service/app/main.py:318-334 -> matcher.py:127-132; dinov3 main.py:168-172), so this -- not the batched
// tile GEMM, whose 128-row MMA tile costs the same tensor time for 1 query as for 128 -- is the per-message re-ID step.
// Pure HBM stream: every warp walks rows w, w + W, ... (ascending), a lane owns the 16-byte chunks lane, lane + 32, ... of a
// row (three for D = 768), the fp32 query slices live in registers (no hi / lo split: fp32 query x bf16 row, fp32 accumulate),
// RU rows are in flight per lane, a 5-step shuffle tree finishes the dot products, and lane q keeps query q's running top-k.
// Block merge through shared memory -> one partial list per CTA; the last CTA to finish merges the lists into the result.
// ---------------------------------------------------------------------------------------------
template <int Q, int CH, int RU>
__global__ void __launch_bounds__(256, 2)
gallery_scan_small_kernel(const float* __restrict__ queries, const __nv_bfloat16* __restrict__ gallery, int rows, int row_base, int k,
                          float* part_s, int32_t* part_i, int slots, float* __restrict__ dump, int* done_counter,
                          float* __restrict__ out_scores, int32_t* __restrict__ out_idx, int out_stride,
                          const float* __restrict__ cut_scores, const int32_t* __restrict__ cut_idx) {
    constexpr int DIM = CH * 256;
    static_assert(CRE_TOPK_MAX == 8, "the block merge below maps 64 candidates onto 32 lanes x 2");
    __shared__ float sh_s[8][Q][CRE_TOPK_MAX];
    __shared__ int sh_i[8][Q][CRE_TOPK_MAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
    float qreg[Q][CH][8];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(queries + q * DIM + 8 * (lane + 32 * c)));
            const float4 b = __ldg(reinterpret_cast<const float4*>(queries + q * DIM + 8 * (lane + 32 * c)) + 1);
            qreg[q][c][0] = a.x; qreg[q][c][1] = a.y; qreg[q][c][2] = a.z; qreg[q][c][3] = a.w;
            qreg[q][c][4] = b.x; qreg[q][c][5] = b.y; qreg[q][c][6] = b.z; qreg[q][c][7] = b.w;
        }
    float ts[CRE_TOPK_MAX];     // lane q (< Q): running top list of query q, sorted under (score desc, index asc)
    int ti[CRE_TOPK_MAX];
#pragma unroll
    for (int j = 0; j < CRE_TOPK_MAX; ++j) { ts[j] = -INFINITY; ti[j] = 0x7fffffff; }
    // later passes of a k > CRE_TOPK_MAX request: only candidates strictly after the previous pass's last entry compete
    float cut_s = INFINITY;
    int cut_i = -1;
    if (cut_scores != nullptr && lane < Q) {
        cut_s = cut_scores[lane * out_stride];
        cut_i = cut_idx[lane * out_stride];
    }

    for (int r0 = gw; r0 < rows; r0 += nw * RU) {
        uint4 g[RU][CH];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = r0 + u * nw;
#pragma unroll
            for (int c = 0; c < CH; ++c)
                g[u][c] = r < rows ? __ldg(reinterpret_cast<const uint4*>(gallery + static_cast<size_t>(r) * DIM) + lane + 32 * c)
                                   : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = r0 + u * nw;          // ascending within the warp: on equal scores the smaller index is met first
            float acc[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) acc[q] = 0.0f;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const uint32_t w[4] = {g[u][c].x, g[u][c].y, g[u][c].z, g[u][c].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xffff0000u);
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        acc[q] = fmaf(lo, qreg[q][c][2 * i], acc[q]);
                        acc[q] = fmaf(hi, qreg[q][c][2 * i + 1], acc[q]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < Q; ++q)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
            if (r < rows) {
                float mine = acc[0];
#pragma unroll
                for (int q = 1; q < Q; ++q) mine = lane == q ? acc[q] : mine;
                if (lane < Q) {
                    if (dump != nullptr) dump[static_cast<size_t>(lane) * rows + r] = mine;
                    if (mine > ts[CRE_TOPK_MAX - 1] && (mine < cut_s || (mine == cut_s && row_base + r > cut_i))) {
                        // insertion + unconditional shift of everything behind (see the EPI_TOPK epilogue: displaced entries must
                        // pass entries of equal score)
                        float cs = mine;
                        int ci = row_base + r;
                        bool shifting = false;
#pragma unroll
                        for (int j = 0; j < CRE_TOPK_MAX; ++j) {
                            if (shifting || cs > ts[j]) {
                                const float t1 = ts[j]; const int t2 = ti[j];
                                ts[j] = cs; ti[j] = ci;
                                cs = t1; ci = t2;
                                shifting = true;
                            }
                        }
                    }
                }
            }
        }
    }
    if (lane < Q) {
#pragma unroll
        for (int j = 0; j < CRE_TOPK_MAX; ++j) { sh_s[warp][lane][j] = ts[j]; sh_i[warp][lane][j] = ti[j]; }
    }
    __syncthreads();
    if (warp < Q) {   // warp q merges the 8 lists of query q: 64 candidates, two per lane, k rounds of a warp arg-best
        const int q = warp;
        float h_s = sh_s[lane >> 3][q][lane & 7], n_s = sh_s[(lane >> 3) + 4][q][lane & 7];
        int h_i = sh_i[lane >> 3][q][lane & 7], n_i = sh_i[(lane >> 3) + 4][q][lane & 7];
        if (better(n_s, n_i, h_s, h_i)) {
            const float t1 = h_s; const int t2 = h_i;
            h_s = n_s; h_i = n_i; n_s = t1; n_i = t2;
        }
        for (int r = 0; r < k; ++r) {
            float ws = h_s;
            int wi = h_i, wl = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float os = __shfl_xor_sync(0xffffffffu, ws, o);
                const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
                const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
                if (better(os, oi, ws, wi) || (os == ws && oi == wi && ol < wl)) { ws = os; wi = oi; wl = ol; }
            }
            if (lane == 0) {
                const size_t o = (static_cast<size_t>(q) * slots + blockIdx.x) * k + r;
                part_s[o] = ws;
                part_i[o] = wi;
            }
            if (lane == wl) { h_s = n_s; h_i = n_i; n_s = -INFINITY; n_i = 0x7fffffff; }
        }
    }

    // ---- final merge by the LAST CTA to finish (no second launch: the whole call is one ~30 us kernel) ----
    __shared__ int is_last;
    __shared__ float red_s[8];
    __shared__ int red_i[8], red_p[8];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                   // this CTA's partial lists are visible device-wide ...
        is_last = atomicAdd(done_counter, 1) == static_cast<int>(gridDim.x) - 1;   // ... before it is counted
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int total = slots * k;                           // <= 256 * kSelNC (checked by the launcher)
    for (int q = 0; q < Q; ++q) {
        const float* ps = part_s + static_cast<size_t>(q) * total;
        const int32_t* pi = part_i + static_cast<size_t>(q) * total;
        block_select_topk(total, k, [&](int c) { return __ldcg(ps + c); }, [&](int c) { return __ldcg(pi + c); },
                          out_scores + q * out_stride, out_idx + q * out_stride, red_s, red_i, red_p);
    }
    if (threadIdx.x == 0) *done_counter = 0;               // ready for the next call on this context
}

// returns 1 if the serving-form scan was launched (then `slots` partial lists per query are complete), 0 if the problem does
// not qualify (the caller falls back to the tile GEMM), negative on error
int launch_gallery_scan_small(const float* queries, int q, int dim, const void* gallery, int rows, int row_base, int k, float* part_s,
                              int32_t* part_i, int slots, float* dump, int* done_counter, float* out_scores, int32_t* out_idx,
                              int out_stride, const float* cut_scores, const int32_t* cut_idx, cudaStream_t stream) {
    if (q < 1 || q > 2 || (dim != 768 && dim != 1024) || done_counter == nullptr || slots * k > 256 * kSelNC) return 0;
    if ((reinterpret_cast<uintptr_t>(gallery) & 15) != 0 || (reinterpret_cast<uintptr_t>(queries) & 15) != 0) return 0;
    LaunchScope scope(CRE_K_GEMM_TOPK, 2.0 * rows * dim, stream);
    const __nv_bfloat16* g = static_cast<const __nv_bfloat16*>(gallery);
#define CRE_SCAN(Q_, CH_, RU_) gallery_scan_small_kernel<Q_, CH_, RU_><<<slots, 256, 0, stream>>>(queries, g, rows, row_base, k, part_s, part_i, slots, dump, done_counter, out_scores, out_idx, out_stride, cut_scores, cut_idx)
    if (q == 1 && dim == 768) CRE_SCAN(1, 3, 4);
    else if (q == 1) CRE_SCAN(1, 4, 4);
    else if (dim == 768) CRE_SCAN(2, 3, 2);
    else CRE_SCAN(2, 4, 2);
#undef CRE_SCAN
    CRE_CUDA_OK(cudaGetLastError());
    return 1;
}

// gallery row <- normalise(momentum * row + (1 - momentum) * unit_q)   (matcher.py:281-285); one CTA.  `master` (optional) is the
// fp32 copy of the row -- what the reference keeps in Qdrant and blends into (matcher.py:267-301): the update reads and writes IT,
// and the bf16 row, the scan copy, is only ever a rounding of the master, so repeated updates do not accumulate bf16 error.
__global__ void __launch_bounds__(256) gallery_update_row_kernel(__nv_bfloat16* __restrict__ row, float* __restrict__ master, int dim,
                                                                 const float* __restrict__ uq, float momentum) {
    __shared__ float red[8];
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = threadIdx.x + 256 * i;
        if (c < dim) {
            const float old = momentum != 0.0f ? (master != nullptr ? master[c] : __bfloat162float(row[c])) : 0.0f;
            v[i] = momentum * old + (1.0f - momentum) * uq[c];
            ss += v[i] * v[i];
        }
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float scale = 1.0f / (sqrtf(tot) + 1e-8f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = threadIdx.x + 256 * i;
        if (c < dim) {
            row[c] = __float2bfloat16_rn(v[i] * scale);
            if (master != nullptr) master[c] = v[i] * scale;
        }
    }
}

int launch_gallery_update_row(__nv_bfloat16* gallery, float* master, int dim, int row, const float* unit_q, float momentum,
                              cudaStream_t stream) {
    CRE_REQUIRE(dim > 0 && dim <= 1024 && row >= 0, "gallery_update_row: bad dim/row");
    LaunchScope scope(CRE_K_GALLERY_UPDATE, 8.0 * dim, stream);
    gallery_update_row_kernel<<<1, 256, 0, stream>>>(gallery + static_cast<size_t>(row) * dim,
                                                     master != nullptr ? master + static_cast<size_t>(row) * dim : nullptr, dim, unit_q,
                                                     momentum);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cre
