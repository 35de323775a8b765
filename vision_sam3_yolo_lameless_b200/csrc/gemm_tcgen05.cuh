// Persistent warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (both K-major
// bf16, fp32 accumulate in TMEM), with the ViT epilogues fused behind the accumulator read-back.
//
//   warp 0 lane 0 : TMA producer   (A tile 128x64, B tile (256/CG)x64, 128B swizzle, mbarrier ring)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma kind::f16, M = 128*CG, N = 256, K = 16 per instruction)
//   warp 2        : TMEM allocator (512 columns = two 256-column accumulator stages)
//   warps 4..11   : epilogue       (tcgen05.ld 32x32b -> fused math -> swizzled smem staging -> TMA store / TMA reduce-add)
//
// CG = 1: one CTA per SM.  CG = 2: a CTA pair (cluster of 2, cta_group::2) shares one 256x256
// accumulator tile; each CTA loads its own 128 A rows and half of the B rows, the leader issues.
//
// Epilogue data path.  tcgen05.ld 32x32b hands every thread one accumulator ROW; writing rows straight to
// global memory touches 32 different 128-byte lines per instruction (LSU-wavefront bound, measured: the
// residual epilogue kept the tensor pipe 21 % busy).  Instead each epilogue warp parks its 32 rows x 128 bytes
// in a private 4 KB smem tile in the TMA 128B-swizzle layout (conflict-free 16-byte stores) and one lane issues
// a bulk tensor store; the residual update  x += lambda * (acc + bias)  is a TMA reduce-add, so the fp32
// residual stream is never read by the SMs at all.
//
// Epilogues implement the arithmetic of HF DINOv3ViT (transformers 5.5.0,
// models/dinov3_vit/modeling_dinov3_vit.py): patch embedding :75-92, q/k/v projection + rotary on
// patch tokens :238-268,:288-310, GELU MLP :385-386, LayerScale + residual :337-343,:440-448.
#pragma once

#include "common.cuh"

namespace cre {

enum GemmEpi : int {
    EPI_BF16 = 0,   // out_bf16[m, n] = acc + bias[n]                                   (TMA store)
    EPI_F32 = 1,    // out_f32[m, n]  = acc + bias[n]                                   (TMA store)
    EPI_QKV = 2,    // bias, rotary on q/k columns of patch tokens, q * q_scale -> bf16  (TMA store, [M, 3*hidden])
    EPI_GELU = 3,   // out_bf16 = gelu_erf(acc + bias)                                  (TMA store)
    EPI_RESID = 4,  // out_f32[m, n] += scale[n] * (acc + bias[n])                      (TMA reduce-add, in place)
    EPI_PATCH = 5,  // out_f32[token_row(m), n] = acc + bias[n]   (patch rows -> token rows; TMA stores into a [frame][token][col] map)
    EPI_TOPK = 6,   // running per-row top-k over the columns this CTA visits  (gallery scan)
    EPI_NONE = 7,   // accumulators are dropped (main-loop tuning only)
    EPI_RESID_LN = 8,  // v = x + scale * (acc + bias): x (TMA load -> store), bf16(v - pivot) and per-row LayerNorm partials
    EPI_RESID_LN3 = 9, // same with ONE x tile per epilogue warp and one pipeline stage more (long-K GEMMs: a tile's main loop
                       // is long enough to hide the x round trips; measured against three x tiles + one stage fewer: slower)
    EPI_RESID_SP = 10, // split residual stream: x = pivot + hi + lo with hi = bf16(x - pivot) (the A operand of the folded GEMMs) and
                       // lo = bf16(x - pivot - hi); v = x + scale * (acc + bias) is re-split around the row's new pivot.  Both halves
                       // come in and go out as 32 x 64 bf16 TMA boxes: 8 bytes per element instead of the 10 of RESID_LN
                       // (fp32 in, fp32 + bf16 out), no fp32 copy of x in HBM at all.  Two chunk slots per epilogue warp.
    EPI_RESID_SP3 = 11,// same results with ONE chunk slot per epilogue warp and more pipeline stages (long-K GEMMs)
};
constexpr bool epi_resid_ln(int epi) { return epi == EPI_RESID_LN || epi == EPI_RESID_LN3; }
constexpr bool epi_resid_sp(int epi) { return epi == EPI_RESID_SP || epi == EPI_RESID_SP3; }
constexpr bool epi_resid_x(int epi) { return epi_resid_ln(epi) || epi_resid_sp(epi); }   // x tiles fetched by the epilogue warps

constexpr int kTopKMax = 8;

struct GemmParams {
    int M, N, K;
    int b_k_extent;  // B's K extent; B's k coordinate is (k mod b_k_extent)  (hi/lo split queries)
    const float* bias;
    const float* scale;
    float* out_f32;
    __nv_bfloat16* out_bf16;
    int ldo;
    // LayerNorm folded into the GEMM (see "LayerNorm folding" below).  Consumer side (EPI_QKV / EPI_GELU / EPI_BF16 with
    // ln_stats_in != nullptr): out = rstd * (acc - (mean - pivot) * c1[n]) + bias[n].  Producer side (EPI_RESID_LN): reads
    // ln_stats_in (previous statistics of the row -> pivot), writes ln_stats_out, out_f32 (x) and out_bf16 (x - pivot).
    const float* ln_stats_in;
    float* ln_stats_out;
    const float* c1;     // [N] column sums of the folded weight (consumer side)
    int ln_slots;        // 128-column partials per row = 2 * hidden / 256
    int ln_stride;       // floats per statistics row = 2 * ln_slots + 4
    float ln_eps;
    int ldo2;            // EPI_RESID_LN: row stride of out_bf16
    __nv_bfloat16* x_lo; // EPI_RESID_SP: low half of the split residual stream (the high half is out_bf16); row stride ldo2
    // EPI_QKV
    const float* rope_axis;  // [(grid_h + grid_w), 52] fp32: per-axis {cos[16], sin[16], -sin[16]} rows (y positions, then x positions)
    int grid_h, grid_w;
    int tokens_per_frame, prefix_tokens, hidden;
    float q_scale;
    // EPI_PATCH
    int patches_per_frame;
    // EPI_TOPK
    int topk;            // k <= kTopKMax
    int topk_stacked;    // EPI_TOPK, M <= 64 queries: ONE A tile = rows [0, 64) hi halves, rows [64, 128) lo halves (see the epilogue)
    int col_base;        // global gallery index of column 0
    float* part_scores;  // [M, slots, k]
    int* part_idx;       // [M, slots, k]
    int part_slots;      // 2 * gridDim.x
    float* dump_scores;  // optional [M, N] full score matrix (parity tests)
    // later passes of a k > kTopKMax request: only candidates strictly AFTER (cut_scores[m * cut_stride], cut_idx[m * cut_stride])
    // in the (score desc, index asc) order compete.  NULL = first pass.
    const float* cut_scores;
    const int* cut_idx;
    int cut_stride;
    int debug_mode;      // -DCRE_TUNING builds only: 1 = no TMA (MMA issue rate), 2 = no MMA (TMA rate), 4 / 8 / 16 = epilogue cut
                         // short, 32 = L2 prefetch of A; results are garbage.  The shipped build compiles none of these paths.
};
#ifdef CRE_TUNING
__device__ __forceinline__ int gemm_dbg(const GemmParams& p) { return p.debug_mode; }
#else
__device__ __forceinline__ constexpr int gemm_dbg(const GemmParams&) { return 0; }
#endif

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kEpiStageBytes = 32 * 128;  // one epilogue warp's staging tile: 32 rows x 128 bytes

constexpr bool epi_tma_store(int epi) {
    return epi == EPI_BF16 || epi == EPI_F32 || epi == EPI_QKV || epi == EPI_GELU || epi == EPI_RESID || epi == EPI_PATCH || epi_resid_x(epi);
}
constexpr bool epi_out_bf16(int epi) { return epi == EPI_BF16 || epi == EPI_QKV || epi == EPI_GELU; }
constexpr int default_stages(int cg) { return cg == 1 ? 4 : 6; }
constexpr int kRopeRowFloats = 52;           // per axis position: cos[16] | sin[16] | -sin[16] | pad[4] (208 B: odd multiple of 16)
constexpr int kRopeTableBytes = 16 * 1024;   // up to 78 axis rows (grid_h + grid_w), e.g. 37 + 37 at 592 x 592
// the QKV kernel trades one pipeline stage for the rotary table in shared memory
// EPI_RESID_LN keeps 2 (LN3: 1) x tiles + the bf16 tile, 4 KB each, per epilogue warp
// EPI_RESID_SP keeps 2 (SP3: 1) chunk slots = (hi tile, lo tile) pairs, 8 KB each, per epilogue warp
constexpr int ln_x_slots(int epi) { return epi == EPI_RESID_LN3 || epi == EPI_RESID_SP3 ? 1 : 2; }
constexpr int ln_warp_bytes(int epi) { return epi_resid_sp(epi) ? ln_x_slots(epi) * 2 * kEpiStageBytes : (ln_x_slots(epi) + 1) * kEpiStageBytes; }
constexpr int kLnStatsPad = 4;               // floats before the partials in one statistics row: [pivot, -, -, -]
constexpr int default_stages_epi(int epi, int cg) {
    return epi == EPI_RESID_LN ? (cg == 1 ? 2 : 4) : epi == EPI_RESID_LN3 ? (cg == 1 ? 3 : 5)
         : epi == EPI_RESID_SP ? (cg == 1 ? 2 : 3) : epi == EPI_RESID_SP3 ? (cg == 1 ? 3 : 5)
                               : default_stages(cg) - (epi == EPI_QKV || epi == EPI_TOPK ? 1 : 0);
}

template <int EPI, int CG, int STAGES>
struct GemmCfg {
    static constexpr int kStages = STAGES;
    // EPI_TOPK stages TWO A boxes per k-block (the hi and the lo half of the fp32 queries, see the MMA issuer): 32 KB
    static constexpr int kSmemA = kBlockM * kBlockK * 2 * (EPI == EPI_TOPK ? 2 : 1);   // 16 KB
    static constexpr int kSmemB = (kBlockN / CG) * kBlockK * 2;   // 32 KB / 16 KB
    static constexpr int kEpiOff = kStages * (kSmemA + kSmemB);
    // EPI_TOPK: four 4 KB exchange tiles (stacked form: the lo-row warps hand their partial scores to the hi-row warps)
    static constexpr int kEpiBytes = epi_resid_x(EPI) ? kEpiWarps * ln_warp_bytes(EPI) : epi_tma_store(EPI) ? kEpiWarps * kEpiStageBytes
                                     : EPI == EPI_TOPK ? 4 * kEpiStageBytes : 0;
    static constexpr int kTableOff = kEpiOff + kEpiBytes;
    static constexpr int kTableBytes = EPI == EPI_QKV ? kRopeTableBytes : 0;
    static constexpr int kBarOff = kTableOff + kTableBytes;
    static constexpr int kSmemBytes = kBarOff + 512 + 1024;       // barriers (+ x-tile barriers) + tmem ptr + align slack
};

// ---------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (FFMA2 / FMUL2): two accumulator columns per instruction in the epilogues
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// gelu(x) = x * Phi(x) with the exact-erf Phi of nn.functional.gelu (HF:activations.py "gelu").
//
// The up-projection (K = 768) is bound by its epilogue, not by its MMAs (ncu: tensor pipe 63 % active, 3.4e8 warp instructions
// per launch), so the activation is priced in issue slots per element.  Default form (CRE_GELU_TANH = 1):
//     Phi(x) = 1/2 + 1/2 tanh(x * (c1 + c3 t + c5 t^2)),  t = min(x^2, 50)
// i.e. the exact identity erf(z) = tanh(atanh(erf(z))) with atanh(erf(x / sqrt 2)) / x fitted by a quadratic in x^2 (near-minimax
// on |x| <= 7.07; beyond that the argument is >= 12.5 and tanh has saturated): max |error| on gelu 2.5e-5 with an exact tanh.
// tanh is ONE MUFU op (tanh.approx.f32, relative error <= 2^-11), which adds at most 2.5e-4 |x| -- below the bf16 half-ulp of the
// output for x > 0 and below 1e-3 absolute for -4 < x < 0 (measured on the B200: tests/test_gpu_kernels.py::test_gelu_epilogue_error).
// Cost per element: 3 FMA-pipe + 1 ALU + 1 MUFU instruction slots instead of 6 + 2 + 0.
//
// CRE_GELU_TANH = 0: Phi(x) = 1/2 + xc * P(xc^2), xc = clamp(x, +-4.5), P a degree-8 near-minimax fit constrained so that
// Phi(4.5) = 1 exactly; max |error| 5.7e-5, all on the FMA pipe (erff() itself costs ~30 FMA-pipe instructions per element).
#ifndef CRE_GELU_TANH
#define CRE_GELU_TANH 1
#endif
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void gelu2(float& x0, float& x1) {
#if CRE_GELU_TANH
    const uint64_t x = pack2(x0, x1);
    float t0, t1, q0, q1;
    unpack2(mul2(x, x), t0, t1);
    const uint64_t t = pack2(fminf(t0, 50.0f), fminf(t1, 50.0f));
    uint64_t p = fma2(pack2(-3.519023885e-04f, -3.519023885e-04f), t, pack2(3.700802103e-02f, 3.700802103e-02f));
    p = fma2(p, t, pack2(7.975052595e-01f, 7.975052595e-01f));
    unpack2(mul2(p, x), q0, q1);
    const uint64_t th = pack2(tanh_approx(q0), tanh_approx(q1));
    const uint64_t hx = mul2(x, pack2(0.5f, 0.5f));
    unpack2(fma2(hx, th, hx), x0, x1);
#else
    const float a = fminf(fmaxf(x0, -4.5f), 4.5f), b = fminf(fmaxf(x1, -4.5f), 4.5f);
    const uint64_t xc = pack2(a, b);
    const uint64_t t = mul2(xc, xc);
    uint64_t p = pack2(2.999970758e-11f, 2.999970758e-11f);
    p = fma2(p, t, pack2(-3.283879391e-09f, -3.283879391e-09f));
    p = fma2(p, t, pack2(1.580244771e-07f, 1.580244771e-07f));
    p = fma2(p, t, pack2(-4.431659363e-06f, -4.431659363e-06f));
    p = fma2(p, t, pack2(8.120941493e-05f, 8.120941493e-05f));
    p = fma2(p, t, pack2(-1.036251313e-03f, -1.036251313e-03f));
    p = fma2(p, t, pack2(9.580645710e-03f, 9.580645710e-03f));
    p = fma2(p, t, pack2(-6.597725302e-02f, -6.597725302e-02f));
    p = fma2(p, t, pack2(3.987085521e-01f, 3.987085521e-01f));
    const uint64_t g = fma2(xc, p, pack2(0.5f, 0.5f));
    const uint64_t y = mul2(pack2(x0, x1), g);
    unpack2(y, x0, x1);
#endif
}

// ---------------------------------------------------------------------------------------------
// TMA store / reduce (smem tile in the 128B-swizzle layout -> global), bulk-group bookkeeping
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
constexpr int kPrefetchKb = 8;
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int EPI, int CG, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2,
               const GemmParams p) {
    using Cfg = GemmCfg<EPI, CG, STAGES>;
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

    // the warp index through a shuffle: provably warp-uniform, so that the role branches below are uniform control flow for the compiler
    // (uniform-datapath address / descriptor math inside them instead of R2UR broadcasts per TMA, MMA and parameter load)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 1) ? 0u : cluster_ctarank();
    const bool is_leader = cta_rank == 0;

    if constexpr (CG > 1) cluster_sync_all();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if constexpr (epi_tma_store(EPI)) tma_prefetch_desc(&tmap_out);
        if constexpr (epi_resid_x(EPI)) tma_prefetch_desc(&tmap_out2);
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
        if constexpr (epi_resid_x(EPI))
            for (int s = 0; s < 4 * kEpiWarps; ++s) mbar_init(bar_base + 256u + 8u * s, 1);
        fence_barrier_init();
    } else if (warp == 2) {
        tmem_alloc<CG>(tmem_slot, 512);
    }
    if constexpr (EPI == EPI_QKV) {
        // rotary per-axis table -> shared memory (read by every epilogue thread for every q / k chunk)
        float* tab = reinterpret_cast<float*>(smem_raw + (smem_base + Cfg::kTableOff - smem_u32(smem_raw)));
        const int nfl = (p.grid_h + p.grid_w) * kRopeRowFloats;
        for (int i = threadIdx.x; i < nfl; i += kGemmThreads) tab[i] = __ldg(p.rope_axis + i);
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
    // EPI_TOPK: A = [hi | lo] (K = 2 * b_k_extent columns); one k-block covers the same 64 gallery columns for both halves, so the
    // gallery tile (the HBM / L2 stream that bounds this kernel) is fetched ONCE and multiplied twice
    const int num_kb = (EPI == EPI_TOPK ? p.b_k_extent : p.K) / kBlockK;

    // Producer and MMA issuer: the WHOLE warp walks the loop (every lane polls the mbarrier) and one elected lane issues.  With the warp
    // converged and every operand derived from provably warp-uniform values (kernel parameters, loop counters, the shuffled TMEM base),
    // descriptors, coordinates and barrier addresses live in uniform registers: a tcgen05.mma / TMA load costs its own issue slot plus
    // a uniform add or two.  Issued from a divergent `lane == 0` branch instead, every one of them was preceded by an election loop and
    // four R2UR broadcasts (~15 dependent instructions on a scheduler the epilogue warps keep busy).
    if (warp == 0) {
        // =============================== TMA producer ===============================
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end && gemm_dbg(p) != 1; t += t_step) {
            const int mt = t / num_nt, nt = t % num_nt;
            const int row0 = (mt * CG + static_cast<int>(cta_rank)) * kBlockM;
            const int col0 = nt * kBlockN + static_cast<int>(cta_rank) * (kBlockN / CG);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    const uint32_t fb = full_bar(stage);
                    if constexpr (CG == 1) {
                        mbar_arrive_expect_tx(fb, (EPI == EPI_TOPK && p.topk_stacked ? kBlockM * kBlockK * 2 : Cfg::kSmemA) + Cfg::kSmemB);
                    } else {
                        // the leader's barrier tracks both halves of the stage: 2 arrivals + the bytes of both CTAs
                        if (is_leader) mbar_arrive_expect_tx(fb, 2 * (Cfg::kSmemA + Cfg::kSmemB));
                        else mbar_arrive_remote(fb, 0);
                    }
                    const int ka = kb * kBlockK;
                    const int kbb = ka % p.b_k_extent;
                    if (gemm_dbg(p) & 32) {
                        // tuning: pull the A box kPrefetchKb k-blocks ahead (next tile's first boxes at the tail) into L2
                        int pkb = kb + kPrefetchKb, prow = row0;
                        if (pkb >= num_kb) {
                            pkb -= num_kb;
                            const int tn = t + t_step;
                            prow = tn < t_end ? ((tn / num_nt) * CG + static_cast<int>(cta_rank)) * kBlockM : -1;
                        }
                        if (prow >= 0) tma_prefetch_2d(&tmap_a, pkb * kBlockK, prow);
                    }
                    tma_load_2d<CG>(&tmap_a, fb, smem_a + stage * Cfg::kSmemA, ka, row0, kEvictNormal);
                    if constexpr (EPI == EPI_TOPK) {
                        // stacked form: 64-row boxes, hi halves of the queries into tile rows [0, 64), lo halves into rows [64, 128)
                        const int lo_off = p.topk_stacked ? 64 * kBlockK * 2 : kBlockM * kBlockK * 2;
                        tma_load_2d<CG>(&tmap_a, fb, smem_a + stage * Cfg::kSmemA + lo_off, ka + p.b_k_extent, row0, kEvictNormal);
                    }
                    tma_load_2d<CG>(&tmap_b, fb, smem_b + stage * Cfg::kSmemB, kbb, col0, EPI == EPI_TOPK ? kEvictNormal : kEvictLast);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1 && is_leader) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * CG, kBlockN);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int t = t_begin; t < t_end; t += t_step, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(tmem_empty_bar(as), aphase ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_u + as * kBlockN;
            for (int kb = 0; kb < num_kb; ++kb) {
                if (gemm_dbg(p) != 1) mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = umma_desc_k_sw128(smem_a + stage * Cfg::kSmemA);
                    const uint64_t db = umma_desc_k_sw128(smem_b + stage * Cfg::kSmemB);
                    if (gemm_dbg(p) != 2) {
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            // +32 B per K=16 step inside the 128 B swizzle atom (descriptor address unit = 16 B)
                            umma_bf16<CG>(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                            if constexpr (EPI == EPI_TOPK) if (!p.topk_stacked) {   // + lo(query) . gallery into the same accumulator
                                const uint64_t da_lo = umma_desc_k_sw128(smem_a + stage * Cfg::kSmemA + kBlockM * kBlockK * 2);
                                umma_bf16<CG>(tmem_d, da_lo + 2 * k, db + 2 * k, idesc, 1u);
                            }
                        }
                    }
                    umma_commit<CG>(empty_bar(stage));
                    if (kb == num_kb - 1) umma_commit<CG>(tmem_full_bar(as));
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // =============================== epilogue ===============================
        // Vector-register copies of the per-column parameter pointers.  They are warp-uniform, and the compiler would otherwise
        // address every LDG through uniform registers at the price of two R2UR per load (10 % of the QKV kernel's instructions).
        uint32_t vzero;
        asm volatile("mov.u32 %0, 0;" : "=r"(vzero));
        const float* bias_v = p.bias + vzero;
        const float* scale_v = p.scale + vzero;
        const float* c1_v = (p.ln_stats_in != nullptr && p.c1 != nullptr ? p.c1 : p.bias) + vzero;   // unfolded: multiplied by nrd = 0
        const int quarter = warp & 3;          // TMEM lane quarter this warp may read
        const int half = (warp - 4) >> 2;      // which 128 accumulator columns
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        // this warp's staging tile (TMA-store epilogues): row r at r*128, 16-byte unit u at (u ^ (r & 7))
        constexpr int kWarpStage = epi_resid_x(EPI) ? ln_warp_bytes(EPI) : kEpiStageBytes;
        constexpr int XS = ln_x_slots(EPI);
        const uint32_t stage_u32 = smem_base + Cfg::kEpiOff + (warp - 4) * kWarpStage;
        uint8_t* stage_row = smem_raw + (stage_u32 - smem_u32(smem_raw)) + lane * 128;
        const uint32_t r7 = lane & 7;
        auto stage_store = [&](int u, uint4 v) {
            *reinterpret_cast<uint4*>(stage_row + ((static_cast<uint32_t>(u) ^ r7) << 4)) = v;
        };
        // wait until the previous bulk store has drained this warp's tile, then it may be overwritten
        auto stage_begin = [&]() {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
        };
        // hand the staged 32 x 128 B tile to the TMA
        auto stage_commit = [&](int col, int row) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if constexpr (EPI == EPI_RESID) tma_reduce_add_2d(&tmap_out, stage_u32, col, row);
                else tma_store_2d(&tmap_out, stage_u32, col, row);
                bulk_commit();
            }
        };

        // LayerNorm folding, consumer side: (rstd, -rstd * (mean - pivot)) of one row from its statistics row
        // [pivot, -, -, -, (mean_i, M2_i) x ln_slots]; Chan's combination of the 128-column partials.
        auto ln_row = [&](int r, bool ok, float& rstd, float& nrd) {
            rstd = 1.0f;
            nrd = 0.0f;
            if (p.ln_stats_in == nullptr || !ok) return;
            const float* so = p.ln_stats_in + static_cast<size_t>(r) * p.ln_stride;
            const float pivot = so[0];
            float mi[8];
            float ms = 0.0f, m2 = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                mi[i] = 0.0f;
                if (i < p.ln_slots) {
                    const float2 e = *reinterpret_cast<const float2*>(so + kLnStatsPad + 2 * i);
                    mi[i] = e.x;
                    ms += e.x;
                    m2 += e.y;
                }
            }
            const float inv_s = 1.0f / static_cast<float>(p.ln_slots);
            const float mean = ms * inv_s;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < p.ln_slots) {
                    const float d = mi[i] - mean;
                    m2 = fmaf(128.0f * d, d, m2);
                }
            rstd = rsqrtf(m2 * inv_s * (1.0f / 128.0f) + p.ln_eps);
            nrd = -rstd * (mean - pivot);
        };

        // EPI_RESID_LN: x-tile ring (XS 4 KB slots per warp, one mbarrier each), XS loads ahead of the chunk in hand
        [[maybe_unused]] int xg = 0;         // running chunk counter of this warp: chunk xg lives in slot xg % XS
        [[maybe_unused]] int sp_pend_slot = -1, sp_pend_col = 0, sp_pend_row = 0;   // EPI_RESID_SP, lane 0: slot refill not yet issued
        auto xbar = [&](int s) { return bar_base + 256u + 8u * static_cast<uint32_t>((warp - 4) * 4 + s); };
        auto x_issue = [&](int s, int col, int rowb) {   // lane 0: arm the slot's barrier and fetch one 32 x 32 fp32 box of x
            if constexpr (epi_resid_sp(EPI)) {           // ... or the 32 x 64 bf16 boxes of its two halves
                mbar_arrive_expect_tx(xbar(s), 2 * kEpiStageBytes);
                tma_load_2d<1>(&tmap_out, xbar(s), stage_u32 + s * 2 * kEpiStageBytes, col, rowb, kEvictNormal);
                tma_load_2d<1>(&tmap_out2, xbar(s), stage_u32 + s * 2 * kEpiStageBytes + kEpiStageBytes, col, rowb, kEvictNormal);
            } else {
                mbar_arrive_expect_tx(xbar(s), kEpiStageBytes);
                tma_load_2d<1>(&tmap_out, xbar(s), stage_u32 + s * kEpiStageBytes, col, rowb, kEvictNormal);
            }
        };
        if constexpr (epi_resid_x(EPI)) {
            if (lane == 0 && t_begin < t_end) {
                const int mt0 = t_begin / num_nt, nt0 = t_begin % num_nt;
                const int rb0 = (mt0 * CG + static_cast<int>(cta_rank)) * kBlockM + quarter * 32;
                const int nc0 = nt0 * kBlockN + half * 128;
                x_issue(0, nc0, rb0);
                if constexpr (XS > 1) x_issue(1, nc0 + (epi_resid_sp(EPI) ? 64 : 32), rb0);
            }
            __syncwarp();
        }

        // EPI_TOPK running state
        float tk_s[kTopKMax];
        int tk_i[kTopKMax];
        int tk_mt = -1;
        [[maybe_unused]] int tk_xc = 0;      // stacked form: chunks this (lo-row) warp has handed over so far
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
        [[maybe_unused]] float cut_s = INFINITY;     // this thread's row: cutoff of the previous pass (none: every finite score is after it)
        [[maybe_unused]] int cut_i = -1;

        int it = 0;
        for (int t = t_begin; t < t_end; t += t_step, ++it) {
            const int mt = t / num_nt, nt = t % num_nt;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int row_base = (mt * CG + static_cast<int>(cta_rank)) * kBlockM + quarter * 32;
            const int row = row_base + lane;
            const bool row_ok = row < p.M;
            const uint32_t taddr = tmem_base + lane_off + as * kBlockN + half * 128;
            const int ncol0 = nt * kBlockN + half * 128;

            if constexpr (EPI == EPI_TOPK) {
                if (mt != tk_mt) {
                    if (tk_mt >= 0) { topk_flush(tk_mt); topk_reset(); }
                    tk_mt = mt;
                    if (p.cut_scores != nullptr && row_ok) {
                        cut_s = p.cut_scores[static_cast<size_t>(row) * p.cut_stride];
                        cut_i = p.cut_idx[static_cast<size_t>(row) * p.cut_stride];
                    }
                }
            }
            // residual epilogues: the row's previous statistics (-> pivots) are fetched BEFORE waiting for the accumulator, so their
            // DRAM round trip hides under the tile's main loop (it was 7 % of the epilogue warps' samples in the attention-out projection)
            [[maybe_unused]] float ln_p_old = 0.0f, ln_pivot = 0.0f;
            if constexpr (epi_resid_x(EPI)) {
                if (row_ok) {
                    const float* so = p.ln_stats_in + static_cast<size_t>(row) * p.ln_stride;
                    ln_p_old = so[0];
                    float ms = 0.0f;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (i < p.ln_slots) ms += so[kLnStatsPad + 2 * i];
                    ln_pivot = ms / static_cast<float>(p.ln_slots);
                }
            }
            mbar_wait(tmem_full_bar(as), aphase);
            __syncwarp();
            tc_fence_after();

            if constexpr (EPI == EPI_NONE) {
                (void)taddr; (void)ncol0; (void)row_ok;
            } else if constexpr (EPI == EPI_QKV) {
                // one 64-column chunk = one head of q, k or v
                const int tok = row_ok ? row % p.tokens_per_frame : 0;
                const bool patch_row = tok >= p.prefix_tokens;
                const int pidx = patch_row ? tok - p.prefix_tokens : 0;
                const int py = pidx / p.grid_w, px = pidx - py * p.grid_w;
                const float* tab = reinterpret_cast<const float*>(smem_raw + (smem_base + Cfg::kTableOff - smem_u32(smem_raw)));
                const float4* ty = reinterpret_cast<const float4*>(tab + py * kRopeRowFloats);                 // cos[16] | sin[16]
                const float4* tx = reinterpret_cast<const float4*>(tab + (p.grid_h + px) * kRopeRowFloats);
                float rstd, nrd;
                ln_row(row, row_ok, rstd, nrd);
                const uint64_t rstd2 = pack2(rstd, rstd), nrd2 = pack2(nrd, nrd);
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    const int n0 = ncol0 + hh * 64;
                    if (n0 >= p.N) break;
                    uint32_t a[32], b[32];
                    tmem_ld32(taddr + hh * 64, a);
                    tmem_ld32(taddr + hh * 64 + 32, b);
                    const int which = n0 / p.hidden;  // 0 = q, 1 = k, 2 = v
                    const bool rot = which < 2 && patch_row;
                    tmem_ld_wait();
                    // packed column pairs: X1[i] = columns (2i, 2i+1) of the head's first half, X2[i] of its second half
                    uint64_t X1[16], X2[16];
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias_v + n0) + j4);
                        const float4 b2 = __ldg(reinterpret_cast<const float4*>(bias_v + n0 + 32) + j4);
                        const float4 k1 = __ldg(reinterpret_cast<const float4*>(c1_v + n0) + j4);
                        const float4 k2 = __ldg(reinterpret_cast<const float4*>(c1_v + n0 + 32) + j4);
                        // rstd * acc + (nrd * c1 + c2), two columns per FFMA2
                        X1[2 * j4] = fma2(rstd2, pack2(__uint_as_float(a[4 * j4]), __uint_as_float(a[4 * j4 + 1])),
                                          fma2(nrd2, pack2(k1.x, k1.y), pack2(b1.x, b1.y)));
                        X1[2 * j4 + 1] = fma2(rstd2, pack2(__uint_as_float(a[4 * j4 + 2]), __uint_as_float(a[4 * j4 + 3])),
                                              fma2(nrd2, pack2(k1.z, k1.w), pack2(b1.z, b1.w)));
                        X2[2 * j4] = fma2(rstd2, pack2(__uint_as_float(b[4 * j4]), __uint_as_float(b[4 * j4 + 1])),
                                          fma2(nrd2, pack2(k2.x, k2.y), pack2(b2.x, b2.y)));
                        X2[2 * j4 + 1] = fma2(rstd2, pack2(__uint_as_float(b[4 * j4 + 2]), __uint_as_float(b[4 * j4 + 3])),
                                              fma2(nrd2, pack2(k2.z, k2.w), pack2(b2.z, b2.w)));
                    }
                    if (rot) {
                        // element j of each half pairs with angle j: j < 16 -> y-angle j, j >= 16 -> x-angle j - 16;
                        // (u, v) -> (u cos - v sin, v cos + u sin)  =  q*cos + rotate_half(q)*sin
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4* t4 = j4 < 4 ? ty + j4 : tx + (j4 - 4);
                            const float4 c4 = t4[0], s4 = t4[4], n4 = t4[8];
                            const uint64_t u0 = X1[2 * j4], u1 = X1[2 * j4 + 1], v0 = X2[2 * j4], v1 = X2[2 * j4 + 1];
                            X1[2 * j4] = fma2(v0, pack2(n4.x, n4.y), mul2(u0, pack2(c4.x, c4.y)));
                            X1[2 * j4 + 1] = fma2(v1, pack2(n4.z, n4.w), mul2(u1, pack2(c4.z, c4.w)));
                            X2[2 * j4] = fma2(u0, pack2(s4.x, s4.y), mul2(v0, pack2(c4.x, c4.y)));
                            X2[2 * j4 + 1] = fma2(u1, pack2(s4.z, s4.w), mul2(v1, pack2(c4.z, c4.w)));
                        }
                    }
                    if (which == 0 && p.q_scale != 1.0f) {   // unfolded weights only: the folded q rows carry head_dim^-0.5 already
                        const uint64_t qs2 = pack2(p.q_scale, p.q_scale);
#pragma unroll
                        for (int i = 0; i < 16; ++i) { X1[i] = mul2(X1[i], qs2); X2[i] = mul2(X2[i], qs2); }
                    }
                    uint32_t o1[16], o2[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float lo, hi;
                        unpack2(X1[i], lo, hi);
                        o1[i] = pack_bf16x2(lo, hi);
                        unpack2(X2[i], lo, hi);
                        o2[i] = pack_bf16x2(lo, hi);
                    }
                    stage_begin();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        stage_store(j, make_uint4(o1[4 * j], o1[4 * j + 1], o1[4 * j + 2], o1[4 * j + 3]));
                        stage_store(j + 4, make_uint4(o2[4 * j], o2[4 * j + 1], o2[4 * j + 2], o2[4 * j + 3]));
                    }
                    stage_commit(n0, row_base);
                }
            } else if constexpr (epi_tma_store(EPI) && epi_out_bf16(EPI)) {
                // EPI_BF16 / EPI_GELU: 64-column chunks, bf16 out
                float rstd, nrd;
                ln_row(row, row_ok, rstd, nrd);
                const uint64_t rstd2 = pack2(rstd, rstd), nrd2 = pack2(nrd, nrd);
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    const int n0 = ncol0 + hh * 64;
                    if (n0 >= p.N) break;
                    if (gemm_dbg(p) & 4) break;          // tuning: accumulators are never read
                    uint32_t a[32], b[32];
                    tmem_ld32(taddr + hh * 64, a);
                    tmem_ld32(taddr + hh * 64 + 32, b);
                    tmem_ld_wait();
                    if (gemm_dbg(p) & 8) {               // tuning: TMEM read only
                        uint32_t acc_or = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc_or |= a[j] | b[j];
                        if (acc_or == 0x7fc12345u) p.out_bf16[0] = __float2bfloat16(0.f);
                        continue;
                    }
                    uint32_t o[32];
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;   // nrd * c1 + c2 for 8 columns (packed pairs)
                        if (p.bias != nullptr) {
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias_v + n0) + j4);
                            const float4 b2 = __ldg(reinterpret_cast<const float4*>(bias_v + n0 + 32) + j4);
                            const float4 k1 = __ldg(reinterpret_cast<const float4*>(c1_v + n0) + j4);
                            const float4 k2 = __ldg(reinterpret_cast<const float4*>(c1_v + n0 + 32) + j4);
                            t0 = fma2(nrd2, pack2(k1.x, k1.y), pack2(b1.x, b1.y));
                            t1 = fma2(nrd2, pack2(k1.z, k1.w), pack2(b1.z, b1.w));
                            t2 = fma2(nrd2, pack2(k2.x, k2.y), pack2(b2.x, b2.y));
                            t3 = fma2(nrd2, pack2(k2.z, k2.w), pack2(b2.z, b2.w));
                        }
                        float v0, v1, v2, v3, w0, w1, w2, w3;
                        unpack2(fma2(rstd2, pack2(__uint_as_float(a[4 * j4]), __uint_as_float(a[4 * j4 + 1])), t0), v0, v1);
                        unpack2(fma2(rstd2, pack2(__uint_as_float(a[4 * j4 + 2]), __uint_as_float(a[4 * j4 + 3])), t1), v2, v3);
                        unpack2(fma2(rstd2, pack2(__uint_as_float(b[4 * j4]), __uint_as_float(b[4 * j4 + 1])), t2), w0, w1);
                        unpack2(fma2(rstd2, pack2(__uint_as_float(b[4 * j4 + 2]), __uint_as_float(b[4 * j4 + 3])), t3), w2, w3);
                        if constexpr (EPI == EPI_GELU) {
                            gelu2(v0, v1); gelu2(v2, v3); gelu2(w0, w1); gelu2(w2, w3);
                        }
                        o[2 * j4] = pack_bf16x2(v0, v1);
                        o[2 * j4 + 1] = pack_bf16x2(v2, v3);
                        o[16 + 2 * j4] = pack_bf16x2(w0, w1);
                        o[16 + 2 * j4 + 1] = pack_bf16x2(w2, w3);
                    }
                    stage_begin();
#pragma unroll
                    for (int u = 0; u < 8; ++u) stage_store(u, make_uint4(o[4 * u], o[4 * u + 1], o[4 * u + 2], o[4 * u + 3]));
                    if (gemm_dbg(p) & 16) continue;      // tuning: staged, never stored
                    stage_commit(n0, row_base);
                }
            } else if constexpr (epi_resid_ln(EPI)) {
                // v = x + scale * (acc + bias) per 32-column chunk: x arrives by TMA in this warp's slot ring, v goes
                // back through the same slot (TMA store), bf16(v - pivot) through the 64-column tile (TMA store every
                // second chunk), and the row's (mean, M2) over this warp's 128 columns is merged chunk by chunk.
                const int slot_id = nt * 2 + half;
                const float pivot = ln_pivot;   // the row's mean at the previous LayerNorm (any value near the mean would do)
                float mean_r = 0.0f, m2_r = 0.0f;
                uint8_t* orow = stage_row + XS * kEpiStageBytes;
                // lane 0: fetch the x tile XS chunks ahead (this tile, or the head of this worker's next tile) into `slot`
                auto x_ahead = [&](int c, int slot) {
                    int c2 = c + XS, nrow = row_base, ncol = ncol0;
                    if (c2 >= 4) {
                        const int tn = t + t_step;
                        if (tn >= t_end) return;
                        c2 -= 4;
                        const int nmt = tn / num_nt, nnt = tn % num_nt;
                        nrow = (nmt * CG + static_cast<int>(cta_rank)) * kBlockM + quarter * 32;
                        ncol = nnt * kBlockN + half * 128;
                    }
                    x_issue(slot, ncol + c2 * 32, nrow);
                };
#pragma unroll 1
                for (int c = 0; c < 4; ++c, ++xg) {
                    const int n0 = ncol0 + c * 32;
                    const int s = xg % XS;
                    const int ob = c & 1;   // which half of the bf16 tile
                    uint32_t a[32];
                    tmem_ld32(taddr + c * 32, a);
                    mbar_wait(xbar(s), static_cast<uint32_t>(xg / XS) & 1u);
                    tmem_ld_wait();
                    uint8_t* xrow = stage_row + s * kEpiStageBytes;
                    float v[32];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 xv = *reinterpret_cast<const float4*>(xrow + ((static_cast<uint32_t>(u) ^ r7) << 4));
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias_v + n0) + u);
                        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale_v + n0) + u);
                        v[4 * u] = fmaf(sc.x, __uint_as_float(a[4 * u]) + b4.x, xv.x);
                        v[4 * u + 1] = fmaf(sc.y, __uint_as_float(a[4 * u + 1]) + b4.y, xv.y);
                        v[4 * u + 2] = fmaf(sc.z, __uint_as_float(a[4 * u + 2]) + b4.z, xv.z);
                        v[4 * u + 3] = fmaf(sc.w, __uint_as_float(a[4 * u + 3]) + b4.w, xv.w);
                    }
                    // chunk statistics (two passes over registers), then Chan's merge with the running pair
                    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) { s0 += v[j]; s1 += v[j + 1]; s2 += v[j + 2]; s3 += v[j + 3]; }
                    const float mc = ((s0 + s1) + (s2 + s3)) * (1.0f / 32.0f);
                    float q0 = 0.0f, q1 = 0.0f;
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float d0 = v[j] - mc, d1 = v[j + 1] - mc;
                        q0 = fmaf(d0, d0, q0);
                        q1 = fmaf(d1, d1, q1);
                    }
                    const float delta = mc - mean_r;
                    const float inv = 1.0f / static_cast<float>(c + 1);
                    mean_r = fmaf(delta, inv, mean_r);
                    m2_r += (q0 + q1) + delta * delta * (32.0f * static_cast<float>(c) * inv);
                    // every earlier store of this warp has drained its smem tiles: the bf16 tile may be overwritten
                    if (lane == 0) bulk_wait_read0();
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        *reinterpret_cast<float4*>(xrow + ((static_cast<uint32_t>(u) ^ r7) << 4)) =
                            make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(orow + ((static_cast<uint32_t>(ob * 4 + j) ^ r7) << 4)) =
                            make_uint4(pack_bf16x2(v[8 * j] - pivot, v[8 * j + 1] - pivot), pack_bf16x2(v[8 * j + 2] - pivot, v[8 * j + 3] - pivot),
                                       pack_bf16x2(v[8 * j + 4] - pivot, v[8 * j + 5] - pivot), pack_bf16x2(v[8 * j + 6] - pivot, v[8 * j + 7] - pivot));
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmap_out, stage_u32 + s * kEpiStageBytes, n0, row_base);
                        if (ob == 1) tma_store_2d(&tmap_out2, stage_u32 + XS * kEpiStageBytes, n0 - 32, row_base);
                        bulk_commit();
                        bulk_wait_read0();   // the chunk XS ahead reuses THIS slot once its store has drained
                        x_ahead(c, s);
                    }
                    __syncwarp();
                }
                if (row_ok) {
                    float* sn = p.ln_stats_out + static_cast<size_t>(row) * p.ln_stride;
                    *reinterpret_cast<float2*>(sn + kLnStatsPad + 2 * slot_id) = make_float2(mean_r, m2_r);
                    if (slot_id == 0) sn[0] = pivot;
                }
            } else if constexpr (epi_resid_sp(EPI)) {
                // Split residual stream, 64-column chunks: the chunk's hi / lo tiles arrive by TMA in this warp's slot ring,
                // x = p_old + (hi + lo) (p_old = the pivot the halves were written with, slot 0 of the incoming statistics row),
                // v = x + scale * (acc + bias), and v - p_new (p_new = the row mean those statistics record) is split again IN
                // PLACE: hi' = bf16(v - p_new), lo' = bf16((v - p_new) - hi') -- 16 significant bits relative to the centred value.
                // Both tiles leave by TMA store; the row's (mean, M2) over this warp's 128 columns is merged 32 columns at a time.
                const int slot_id = nt * 2 + half;
                const float p_old = ln_p_old, pivot = ln_pivot;
                const uint64_t p_old2 = pack2(p_old, p_old), npiv2 = pack2(-pivot, -pivot);
                float mean_r = 0.0f, m2_r = 0.0f;
                // lane 0: coordinates of the chunk XS ahead (this tile, or the head of this worker's next tile); false = none left
                auto x_target = [&](int c, int& ncol, int& nrow) {
                    int c2 = c + XS;
                    nrow = row_base;
                    ncol = ncol0;
                    if (c2 >= 2) {
                        const int tn = t + t_step;
                        if (tn >= t_end) return false;
                        c2 -= 2;
                        const int nmt = tn / num_nt, nnt = tn % num_nt;
                        nrow = (nmt * CG + static_cast<int>(cta_rank)) * kBlockM + quarter * 32;
                        ncol = nnt * kBlockN + half * 128;
                    }
                    ncol += c2 * 64;
                    return true;
                };
#pragma unroll 1
                for (int c = 0; c < 2; ++c, ++xg) {
                    const int s = xg % XS;
                    uint8_t* hrow = stage_row + s * 2 * kEpiStageBytes;
                    uint8_t* lrow = hrow + kEpiStageBytes;
                    uint32_t acc[32];
                    tmem_ld32(taddr + c * 64, acc);
                    mbar_wait(xbar(s), static_cast<uint32_t>(xg / XS) & 1u);
                    tmem_ld_wait();
                    if constexpr (XS > 1) {
                        // the refill of the OTHER slot, left pending by the previous chunk: its stores have had this chunk's two waits to
                        // drain, so lane 0 no longer stalls the warp on them (8.5 % of the epilogue warps' samples in the attention-out
                        // projection), and the load still has a whole chunk of arithmetic to land
                        if (lane == 0 && sp_pend_slot >= 0) {
                            bulk_wait_read0();
                            x_issue(sp_pend_slot, sp_pend_col, sp_pend_row);
                            sp_pend_slot = -1;
                        }
                    }
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int n0 = ncol0 + c * 64 + hh * 32;
                        uint64_t v[16];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t off = (static_cast<uint32_t>(hh * 4 + u) ^ r7) << 4;
                            const uint4 hv = *reinterpret_cast<const uint4*>(hrow + off);
                            const uint4 lv = *reinterpret_cast<const uint4*>(lrow + off);
                            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias_v + n0) + 2 * u);
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias_v + n0) + 2 * u + 1);
                            const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale_v + n0) + 2 * u);
                            const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale_v + n0) + 2 * u + 1);
                            const uint64_t bb[4] = {pack2(b0.x, b0.y), pack2(b0.z, b0.w), pack2(b1.x, b1.y), pack2(b1.z, b1.w)};
                            const uint64_t ss[4] = {pack2(s0.x, s0.y), pack2(s0.z, s0.w), pack2(s1.x, s1.y), pack2(s1.z, s1.w)};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint64_t xh = pack2(__uint_as_float(hw[e] << 16), __uint_as_float(hw[e] & 0xffff0000u));
                                const uint64_t xl = pack2(__uint_as_float(lw[e] << 16), __uint_as_float(lw[e] & 0xffff0000u));
                                const uint64_t x2 = add2(p_old2, add2(xh, xl));
                                const uint64_t t2 = add2(pack2(__uint_as_float(acc[8 * u + 2 * e]), __uint_as_float(acc[8 * u + 2 * e + 1])), bb[e]);
                                v[4 * u + e] = fma2(ss[e], t2, x2);
                            }
                        }
                        // the accumulators are consumed: the second half's TMEM read flies under the statistics and the re-split
                        if (hh == 0) tmem_ld32(taddr + c * 64 + 32, acc);
                        // statistics of these 32 columns (two passes over registers), then Chan's merge with the running pair
                        uint64_t sa = v[0], sb = v[1];
#pragma unroll
                        for (int j = 2; j < 16; j += 2) { sa = add2(sa, v[j]); sb = add2(sb, v[j + 1]); }
                        float s_lo, s_hi;
                        unpack2(add2(sa, sb), s_lo, s_hi);
                        const float mc = (s_lo + s_hi) * (1.0f / 32.0f);
                        const uint64_t nmc2 = pack2(-mc, -mc);
                        uint64_t qa = pack2(0.0f, 0.0f), qb = qa;
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const uint64_t d0 = add2(v[j], nmc2), d1 = add2(v[j + 1], nmc2);
                            qa = fma2(d0, d0, qa);
                            qb = fma2(d1, d1, qb);
                        }
                        float q_lo, q_hi;
                        unpack2(add2(qa, qb), q_lo, q_hi);
                        const int kc = 2 * c + hh;               // 32-column groups merged so far
                        const float delta = mc - mean_r;
                        const float inv = 1.0f / static_cast<float>(kc + 1);
                        mean_r = fmaf(delta, inv, mean_r);
                        m2_r += (q_lo + q_hi) + delta * delta * (32.0f * static_cast<float>(kc) * inv);
                        // re-split around the new pivot, in place (a thread only ever touches its own row of the two tiles)
                        const uint64_t neg1 = pack2(-1.0f, -1.0f);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            uint32_t hw[4], lw[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint64_t d2 = add2(v[4 * u + e], npiv2);
                                float d0, d1, r0, r1;
                                unpack2(d2, d0, d1);
                                hw[e] = pack_bf16x2(d0, d1);
                                const uint64_t hf = pack2(__uint_as_float(hw[e] << 16), __uint_as_float(hw[e] & 0xffff0000u));
                                unpack2(fma2(hf, neg1, d2), r0, r1);
                                lw[e] = pack_bf16x2(r0, r1);
                            }
                            const uint32_t off = (static_cast<uint32_t>(hh * 4 + u) ^ r7) << 4;
                            *reinterpret_cast<uint4*>(hrow + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                            *reinterpret_cast<uint4*>(lrow + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                        }
                        if (hh == 0) tmem_ld_wait();
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        const int n0 = ncol0 + c * 64;
                        tma_store_2d(&tmap_out, stage_u32 + s * 2 * kEpiStageBytes, n0, row_base);
                        tma_store_2d(&tmap_out2, stage_u32 + s * 2 * kEpiStageBytes + kEpiStageBytes, n0, row_base);
                        bulk_commit();
                        int ncol, nrow;
                        const bool more = x_target(c, ncol, nrow);   // the chunk XS ahead reuses THIS slot once its stores have drained
                        if constexpr (XS > 1) {
                            if (more) { sp_pend_slot = s; sp_pend_col = ncol; sp_pend_row = nrow; }
                        } else {
                            bulk_wait_read0();
                            if (more) x_issue(s, ncol, nrow);
                        }
                    }
                    __syncwarp();
                }
                if (row_ok) {
                    float* sn = p.ln_stats_out + static_cast<size_t>(row) * p.ln_stride;
                    *reinterpret_cast<float2*>(sn + kLnStatsPad + 2 * slot_id) = make_float2(mean_r, m2_r);
                    if (slot_id == 0) sn[0] = pivot;
                }
            } else if constexpr (epi_tma_store(EPI)) {
                // EPI_F32 / EPI_RESID / EPI_PATCH: 32-column chunks, fp32 out (store / reduce-add / store into the token-row map)
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int n0 = ncol0 + c * 32;
                    if (n0 >= p.N) break;
                    uint32_t a[32];
                    tmem_ld32(taddr + c * 32, a);
                    tmem_ld_wait();
                    float x[32];
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(bias_v + n0) + j4);
                        x[4 * j4] = __uint_as_float(a[4 * j4]) + b4.x;
                        x[4 * j4 + 1] = __uint_as_float(a[4 * j4 + 1]) + b4.y;
                        x[4 * j4 + 2] = __uint_as_float(a[4 * j4 + 2]) + b4.z;
                        x[4 * j4 + 3] = __uint_as_float(a[4 * j4 + 3]) + b4.w;
                        if constexpr (EPI == EPI_RESID) {
                            const float4 sc = __ldg(reinterpret_cast<const float4*>(scale_v + n0) + j4);
                            x[4 * j4] *= sc.x; x[4 * j4 + 1] *= sc.y; x[4 * j4 + 2] *= sc.z; x[4 * j4 + 3] *= sc.w;
                        }
                    }
                    stage_begin();
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        stage_store(u, make_uint4(__float_as_uint(x[4 * u]), __float_as_uint(x[4 * u + 1]),
                                                  __float_as_uint(x[4 * u + 2]), __float_as_uint(x[4 * u + 3])));
                    if constexpr (EPI == EPI_PATCH) {
                        // Patch row m = frame * P + r lands on token row prefix + r of its frame.  tmap_out is [frame][token][col]: the
                        // slab's 32 rows go out as ONE box at (frame, prefix + r0); rows past the frame's last token are clipped by the
                        // map (as are slab rows >= M: their frame index is out of range).  One slab in six runs into the next frame: its
                        // rows >= s1 -- that frame's first patches -- are stored straight from registers.
                        fence_proxy_async_smem();
                        __syncwarp();
                        const int f0 = row_base / p.patches_per_frame, r0 = row_base - f0 * p.patches_per_frame;
                        const int s1 = p.patches_per_frame - r0;               // slab rows that belong to frame f0
                        if (lane == 0) {
                            tma_store_3d(&tmap_out, stage_u32, n0, p.prefix_tokens + r0, f0);
                            bulk_commit();
                        }
                        if (lane >= s1 && row_ok) {
                            float4* o = reinterpret_cast<float4*>(p.out_f32 + (static_cast<size_t>(f0 + 1) * p.tokens_per_frame +
                                                                               p.prefix_tokens + (lane - s1)) * p.ldo + n0);
#pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4) o[j4] = make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
                        }
                    } else {
                        stage_commit(n0, row_base);
                    }
                }
            } else {
                // EPI_TOPK: running per-row top-k, no output tile
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
                        if (p.topk_stacked) {
                            // Stacked form (M <= 64): accumulator rows [0, 64) hold hi(q) . g, rows [64, 128) lo(q) . g -- ONE MMA per
                            // k-step instead of two (and half the query bytes per k-block from L2, which is what bounds this kernel at
                            // small M).  The warps of lane quarters 2 / 3 hand their 32 x 32 partial scores to quarters 0 / 1 through a
                            // swizzled shared-memory tile; named barriers (64 threads) pace the pair, one chunk at a time.
                            const int pair = half * 2 + (quarter & 1);
                            uint8_t* xt = smem_raw + (smem_base + Cfg::kEpiOff + pair * kEpiStageBytes - smem_u32(smem_raw)) + lane * 128;
                            if (quarter >= 2) {
                                if (tk_xc > 0) asm volatile("bar.sync %0, 64;" ::"r"(5 + pair) : "memory");     // tile drained by the partner
#pragma unroll
                                for (int u = 0; u < 8; ++u)
                                    *reinterpret_cast<float4*>(xt + ((static_cast<uint32_t>(u) ^ r7) << 4)) =
                                        make_float4(x[4 * u], x[4 * u + 1], x[4 * u + 2], x[4 * u + 3]);
                                asm volatile("bar.arrive %0, 64;" ::"r"(1 + pair) : "memory");
                                ++tk_xc;
                                continue;
                            }
                            asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const float4 v = *reinterpret_cast<const float4*>(xt + ((static_cast<uint32_t>(u) ^ r7) << 4));
                                x[4 * u] += v.x; x[4 * u + 1] += v.y; x[4 * u + 2] += v.z; x[4 * u + 3] += v.w;
                            }
                            asm volatile("bar.arrive %0, 64;" ::"r"(5 + pair) : "memory");
                        }
                    }

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
                                if (n < p.N && s > tk_s[kTopKMax - 1] && (s < cut_s || (s == cut_s && p.col_base + n > cut_i))) {
                                    // insertion: find the first entry the candidate beats, then SHIFT everything behind it down --
                                    // unconditionally: a displaced entry must also pass entries of EQUAL score (it has the smaller
                                    // index), or a run of ties ends up out of index order and loses the wrong member at the tail
                                    float cs_ = s;
                                    int ci_ = p.col_base + n;
                                    bool shifting = false;
#pragma unroll
                                    for (int q = 0; q < kTopKMax; ++q) {
                                        if (shifting || cs_ > tk_s[q]) {
                                            const float ts = tk_s[q]; const int ti = tk_i[q];
                                            tk_s[q] = cs_; tk_i[q] = ci_;
                                            cs_ = ts; ci_ = ti;
                                            shifting = true;
                                        }
                                    }
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
        if constexpr (epi_tma_store(EPI)) {
            if (lane == 0) bulk_wait0();   // every bulk store of this warp has completed before the CTA may exit
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
