// Persistent warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (both K-major
// bf16, fp32 accumulate in TMEM), with the ViT epilogues fused behind the accumulator read-back.
//
//   warp 0 lane 0 : TMA producer   (A tile 128x64, B tile (256/CG)x64, 128B swizzle, mbarrier ring)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma kind::f16, M = 128*CG, N = 256, K = 16 per instruction)
//   warp 2        : TMEM allocator (512 columns = two 256-column accumulator stages)
//   warps 4..11   : epilogue       (tcgen05.ld 32x32b -> fused bias/RoPE/GELU/LayerScale/residual -> global)
//
// CG = 1: one CTA per SM.  CG = 2: a CTA pair (cluster of 2, cta_group::2) shares one 256x256
// accumulator tile; each CTA loads its own 128 A rows and half of the B rows, the leader issues.
//
// Epilogues implement the arithmetic of HF DINOv3ViT (transformers 5.5.0,
// models/dinov3_vit/modeling_dinov3_vit.py): patch embedding :75-92, q/k/v projection + rotary on
// patch tokens :238-268,:288-310, GELU MLP :385-386, LayerScale + residual :337-343,:440-448.
#pragma once

#include "common.cuh"

namespace cre {

enum GemmEpi : int {
    EPI_BF16 = 0,   // out_bf16[m, n] = acc + bias[n]
    EPI_F32 = 1,    // out_f32[m, n]  = acc + bias[n]
    EPI_QKV = 2,    // bias, rotary on q/k columns of patch tokens, q * q_scale -> bf16
    EPI_GELU = 3,   // out_bf16 = gelu_erf(acc + bias)
    EPI_RESID = 4,  // out_f32[m, n] += scale[n] * (acc + bias[n])            (residual stream, in place)
    EPI_PATCH = 5,  // out_f32[token_row(m), n] = acc + bias[n]               (patch rows -> token rows)
    EPI_TOPK = 6,   // running per-row top-k over the columns this CTA visits  (gallery scan)
    EPI_NONE = 7,   // accumulators are dropped (main-loop tuning only)
};

constexpr int kTopKMax = 8;

struct GemmParams {
    int M, N, K;
    int b_k_extent;  // B's K extent; B's k coordinate is (k mod b_k_extent)  (hi/lo split queries)
    const float* bias;
    const float* scale;
    float* out_f32;
    __nv_bfloat16* out_bf16;
    int ldo;
    // EPI_QKV
    const float* rope_cos;  // [patches, 64] fp32, HF layout (angle j == angle j+32)
    const float* rope_sin;
    int tokens_per_frame, prefix_tokens, hidden;
    float q_scale;
    __nv_bfloat16* vt;   // [frames * heads * 64, t_pad]: v written transposed for the attention PV operand
    int t_pad;
    // EPI_PATCH
    int patches_per_frame;
    // EPI_TOPK
    int topk;            // k <= kTopKMax
    int col_base;        // global gallery index of column 0
    float* part_scores;  // [M, slots, k]
    int* part_idx;       // [M, slots, k]
    int part_slots;      // 2 * gridDim.x
    float* dump_scores;  // optional [M, N] full score matrix (parity tests)
    int debug_mode;      // tuning only: 1 = no TMA (MMA issue rate), 2 = no MMA (TMA rate); results are garbage
};

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;
constexpr int kEpiWarps = 8;

constexpr int default_stages(int cg) { return cg == 1 ? 4 : 6; }

template <int CG, int STAGES>
struct GemmCfg {
    static constexpr int kStages = STAGES;
    static constexpr int kSmemA = kBlockM * kBlockK * 2;          // 16 KB
    static constexpr int kSmemB = (kBlockN / CG) * kBlockK * 2;   // 32 KB / 16 KB
    static constexpr int kBarOff = kStages * (kSmemA + kSmemB);
    static constexpr int kSmemBytes = kBarOff + 256 + 1024;       // barriers + tmem ptr + align slack
};

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <int EPI, int CG, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmParams p) {
    using Cfg = GemmCfg<CG, STAGES>;
    constexpr int kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_base + kStages * Cfg::kSmemA;
    const uint32_t bar_base = smem_base + Cfg::kBarOff;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 1) ? 0u : cluster_ctarank();
    const bool is_leader = cta_rank == 0;

    if constexpr (CG > 1) cluster_sync_all();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), CG);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full_bar(s), 1);
            mbar_init(tmem_empty_bar(s), kEpiWarps * CG);
        }
        fence_barrier_init();
    } else if (warp == 2) {
        tmem_alloc<CG>(tmem_slot, 512);
    }
    tc_fence_before();
    if constexpr (CG > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // ---- tile schedule -------------------------------------------------------------------------
    const int rows_per_tile = kBlockM * CG;
    const int num_mt = (p.M + rows_per_tile - 1) / rows_per_tile;
    const int num_nt = (p.N + kBlockN - 1) / kBlockN;
    const int num_tiles = num_mt * num_nt;
    const int num_workers = gridDim.x / CG;
    const int worker = blockIdx.x / CG;
    int t_begin, t_end, t_step;
    if constexpr (EPI == EPI_TOPK) {  // contiguous range so a row's running top-k stays in registers
        const int per = (num_tiles + num_workers - 1) / num_workers;
        t_begin = worker * per;
        t_end = min(num_tiles, t_begin + per);
        t_step = 1;
    } else {
        t_begin = worker;
        t_end = num_tiles;
        t_step = num_workers;
    }
    const int num_kb = p.K / kBlockK;

    if (warp == 0 && lane == 0) {
        // =============================== TMA producer ===============================
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end && p.debug_mode != 1; t += t_step) {
            const int mt = t / num_nt, nt = t % num_nt;
            const int row0 = (mt * CG + static_cast<int>(cta_rank)) * kBlockM;
            const int col0 = nt * kBlockN + static_cast<int>(cta_rank) * (kBlockN / CG);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t fb = full_bar(stage);
                if constexpr (CG == 1) {
                    mbar_arrive_expect_tx(fb, Cfg::kSmemA + Cfg::kSmemB);
                } else {
                    // the leader's barrier tracks both halves of the stage: 2 arrivals + the bytes of both CTAs
                    if (is_leader) mbar_arrive_expect_tx(fb, 2 * (Cfg::kSmemA + Cfg::kSmemB));
                    else mbar_arrive_remote(fb, 0);
                }
                const int ka = kb * kBlockK;
                const int kbb = ka % p.b_k_extent;
                tma_load_2d<CG>(&tmap_a, fb, smem_a + stage * Cfg::kSmemA, ka, row0, kEvictNormal);
                tma_load_2d<CG>(&tmap_b, fb, smem_b + stage * Cfg::kSmemB, kbb, col0, kEvictLast);
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1 && lane == 0 && is_leader) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * CG, kBlockN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int t = t_begin; t < t_end; t += t_step, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(tmem_empty_bar(as), aphase ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * kBlockN;
            for (int kb = 0; kb < num_kb; ++kb) {
                if (p.debug_mode != 1) mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint64_t da = umma_desc_k_sw128(smem_a + stage * Cfg::kSmemA);
                const uint64_t db = umma_desc_k_sw128(smem_b + stage * Cfg::kSmemB);
                if (p.debug_mode != 2) {
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        // +32 B per K=16 step inside the 128 B swizzle atom (descriptor address unit = 16 B)
                        umma_bf16<CG>(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                }
                umma_commit<CG>(empty_bar(stage));
                if (kb == num_kb - 1) umma_commit<CG>(tmem_full_bar(as));
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // =============================== epilogue ===============================
        const int quarter = warp & 3;          // TMEM lane quarter this warp may read
        const int half = (warp - 4) >> 2;      // which 128 accumulator columns
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;

        // EPI_TOPK running state
        float tk_s[kTopKMax];
        int tk_i[kTopKMax];
        int tk_mt = -1;
        auto topk_reset = [&]() {
#pragma unroll
            for (int j = 0; j < kTopKMax; ++j) { tk_s[j] = -INFINITY; tk_i[j] = 0x7fffffff; }
        };
        auto topk_flush = [&](int mt) {
            const int row = mt * kBlockM + quarter * 32 + lane;
            if (row < p.M) {
                const size_t o = (static_cast<size_t>(row) * p.part_slots + worker * 2 + half) * p.topk;
#pragma unroll
                for (int j = 0; j < kTopKMax; ++j)
                    if (j < p.topk) { p.part_scores[o + j] = tk_s[j]; p.part_idx[o + j] = tk_i[j]; }
            }
        };
        if constexpr (EPI == EPI_TOPK) topk_reset();

        int it = 0;
        for (int t = t_begin; t < t_end; t += t_step, ++it) {
            const int mt = t / num_nt, nt = t % num_nt;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int row = (mt * CG + static_cast<int>(cta_rank)) * kBlockM + quarter * 32 + lane;
            const bool row_ok = row < p.M;
            mbar_wait(tmem_full_bar(as), aphase);
            __syncwarp();
            tc_fence_after();
            const uint32_t taddr = tmem_base + lane_off + as * kBlockN + half * 128;
            const int ncol0 = nt * kBlockN + half * 128;

            if constexpr (EPI == EPI_TOPK) {
                if (mt != tk_mt) {
                    if (tk_mt >= 0) { topk_flush(tk_mt); topk_reset(); }
                    tk_mt = mt;
                }
            }

            if constexpr (EPI == EPI_NONE) {
                (void)taddr; (void)ncol0; (void)row_ok;
            } else if constexpr (EPI == EPI_QKV) {
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    const int n0 = ncol0 + hh * 64;
                    if (n0 >= p.N) break;
                    uint32_t a[32], b[32];
                    tmem_ld32(taddr + hh * 64, a);
                    tmem_ld32(taddr + hh * 64 + 32, b);
                    tmem_ld_wait();
                    const int which = n0 / p.hidden;  // 0 = q, 1 = k, 2 = v
                    const int tok = row % p.tokens_per_frame;
                    const bool rot = which < 2 && tok >= p.prefix_tokens;
                    const float* cs = p.rope_cos + static_cast<size_t>(rot ? tok - p.prefix_tokens : 0) * 64;
                    const float* sn = p.rope_sin + static_cast<size_t>(rot ? tok - p.prefix_tokens : 0) * 64;
                    const float qs = which == 0 ? p.q_scale : 1.0f;
                    float x1[32], x2[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        x1[j] = __uint_as_float(a[j]) + __ldg(p.bias + n0 + j);
                        x2[j] = __uint_as_float(b[j]) + __ldg(p.bias + n0 + 32 + j);
                    }
                    if (rot) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs) + j4);
                            const float4 s4 = __ldg(reinterpret_cast<const float4*>(sn) + j4);
                            const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
                            const float ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int j = j4 * 4 + e;
                                const float u = x1[j], v = x2[j];
                                x1[j] = u * cc[e] - v * ss[e];   // q*cos + rotate_half(q)*sin, first half
                                x2[j] = v * cc[e] + u * ss[e];   // second half
                            }
                        }
                    }
                    if (row_ok && which < 2) {
                        // q and k: [row, 2*hidden], 128 contiguous bytes per (row, head)
                        uint4* o1 = reinterpret_cast<uint4*>(p.out_bf16 + static_cast<size_t>(row) * p.ldo + n0);
                        uint4* o2 = o1 + 4;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            o1[j] = make_uint4(pack_bf16x2(x1[8 * j] * qs, x1[8 * j + 1] * qs),
                                               pack_bf16x2(x1[8 * j + 2] * qs, x1[8 * j + 3] * qs),
                                               pack_bf16x2(x1[8 * j + 4] * qs, x1[8 * j + 5] * qs),
                                               pack_bf16x2(x1[8 * j + 6] * qs, x1[8 * j + 7] * qs));
                            o2[j] = make_uint4(pack_bf16x2(x2[8 * j] * qs, x2[8 * j + 1] * qs),
                                               pack_bf16x2(x2[8 * j + 2] * qs, x2[8 * j + 3] * qs),
                                               pack_bf16x2(x2[8 * j + 4] * qs, x2[8 * j + 5] * qs),
                                               pack_bf16x2(x2[8 * j + 6] * qs, x2[8 * j + 7] * qs));
                        }
                    } else if (row_ok) {
                        // v: transposed, vt[(frame*heads + head)*64 + d, tok]; a warp's 32 lanes are
                        // 32 consecutive tokens, so each store instruction writes one 64-byte run
                        const int frame = row / p.tokens_per_frame;
                        const int head = (n0 - 2 * p.hidden) >> 6;
                        const int heads = p.hidden >> 6;
                        __nv_bfloat16* o = p.vt + (static_cast<size_t>(frame) * heads + head) * 64 * p.t_pad + tok;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            o[static_cast<size_t>(j) * p.t_pad] = __float2bfloat16_rn(x1[j]);
                            o[static_cast<size_t>(j + 32) * p.t_pad] = __float2bfloat16_rn(x2[j]);
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int n0 = ncol0 + c * 32;
                    if (n0 >= p.N) break;
                    uint32_t a[32];
                    tmem_ld32(taddr + c * 32, a);
                    tmem_ld_wait();
                    float x[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(a[j]);

                    if constexpr (EPI == EPI_TOPK) {
                        if (row_ok) {
                            if (p.dump_scores != nullptr) {
                                for (int j = 0; j < 32; ++j)
                                    if (n0 + j < p.N) p.dump_scores[static_cast<size_t>(row) * p.N + n0 + j] = x[j];
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int n = n0 + j;
                                const float s = x[j];
                                // columns arrive in ascending index, so on ties the earlier (smaller) index stays ahead
                                if (n < p.N && s > tk_s[kTopKMax - 1]) {
                                    float cs_ = s;
                                    int ci_ = p.col_base + n;
#pragma unroll
                                    for (int q = 0; q < kTopKMax; ++q) {
                                        if (cs_ > tk_s[q]) {
                                            const float ts = tk_s[q]; const int ti = tk_i[q];
                                            tk_s[q] = cs_; tk_i[q] = ci_;
                                            cs_ = ts; ci_ = ti;
                                        }
                                    }
                                }
                            }
                        }
                    } else {
                        if (p.bias != nullptr) {
#pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j4);
                                x[4 * j4] += b4.x; x[4 * j4 + 1] += b4.y; x[4 * j4 + 2] += b4.z; x[4 * j4 + 3] += b4.w;
                            }
                        }
                        if constexpr (EPI == EPI_GELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
                        }
                        if (row_ok) {
                            if constexpr (EPI == EPI_BF16 || EPI == EPI_GELU) {
                                uint4* o = reinterpret_cast<uint4*>(p.out_bf16 + static_cast<size_t>(row) * p.ldo + n0);
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    o[j] = make_uint4(pack_bf16x2(x[8 * j], x[8 * j + 1]), pack_bf16x2(x[8 * j + 2], x[8 * j + 3]),
                                                      pack_bf16x2(x[8 * j + 4], x[8 * j + 5]), pack_bf16x2(x[8 * j + 6], x[8 * j + 7]));
                            } else {
                                size_t orow = static_cast<size_t>(row);
                                if constexpr (EPI == EPI_PATCH)
                                    orow = static_cast<size_t>(row / p.patches_per_frame) * p.tokens_per_frame +
                                           p.prefix_tokens + row % p.patches_per_frame;
                                float4* o = reinterpret_cast<float4*>(p.out_f32 + orow * p.ldo + n0);
                                if constexpr (EPI == EPI_RESID) {
#pragma unroll
                                    for (int j4 = 0; j4 < 8; ++j4) {
                                        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + n0) + j4);
                                        float4 r = o[j4];
                                        r.x += sc.x * x[4 * j4]; r.y += sc.y * x[4 * j4 + 1];
                                        r.z += sc.z * x[4 * j4 + 2]; r.w += sc.w * x[4 * j4 + 3];
                                        o[j4] = r;
                                    }
                                } else {
#pragma unroll
                                    for (int j4 = 0; j4 < 8; ++j4)
                                        o[j4] = make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
                                }
                            }
                        }
                    }
                }
            }
            // release this accumulator stage back to the MMA issuer (leader CTA owns the barrier)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 1) mbar_arrive(tmem_empty_bar(as));
                else mbar_arrive_remote(tmem_empty_bar(as), 0);
            }
        }
        if constexpr (EPI == EPI_TOPK) {
            if (tk_mt >= 0) topk_flush(tk_mt);
        }
    }

    tc_fence_before();
    if constexpr (CG > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, 512);
    }
}

}  // namespace cre
