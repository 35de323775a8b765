// Non-causal multi-head attention for the ViT blocks (HF:modeling_dinov3_vit.py:210-235,316-329):
//   O = softmax(Q K^T) V   per (frame, head), head_dim 64, no mask, fp32 softmax.
// Input is the fused projection matrix qkv bf16 [n*t, ld] written by the QKV GEMM epilogue
// (gemm_tcgen05.cuh EPI_QKV): q at column head*64 (pre-scaled by head_dim^-0.5, rotary applied), k at
// k_col0 + head*64 (rotary applied), v at v_col0 + head*64.  All three operands are fetched by TMA straight from
// that matrix; nothing is transposed or padded in memory:
//
//   S[128, KB] = Q[128, 64] * K[KB, 64]^T        tcgen05.mma, both operands K-major, accumulator in TMEM cols [0, KB)
//   P = exp2((S - rowmax) * log2 e)              one thread per query row (tcgen05.ld 32x32b),
//                                                 bf16 P written to smem in the UMMA K-major 128B-swizzle layout
//   O[128, 64] += P[128, KB] * V[KB, 64]         tcgen05.mma, B = the V tile as loaded ([key][dim], dim contiguous)
//                                                 = an MN-major operand (instruction-descriptor bit 16), TMEM cols [0, 64)
//
// One CTA = one 128-row query tile of one (frame, head); keys are visited in blocks of KB <= 256
// (a single block for the 201-token 224x224 case) with the usual running max / running sum.  Keys past the end
// of the frame inside a block belong to the next frame (or are TMA zero fill): their P is forced to 0.
#include "common.cuh"
#include "gemm_tcgen05.cuh"
#include "internal.h"

namespace cre {

constexpr int kAttnThreads = 128;
constexpr int kAttnQBytes = 128 * 128;           // 128 rows x 64 bf16
constexpr int kAttnKPBytes = 4 * 128 * 128;      // K block (<= 256 x 128 B) aliased with P (4 chunks of 128 x 128 B)
constexpr int kAttnVBytes = 256 * 128;           // V block: <= 256 keys x 64 dims
constexpr int kAttnSmemBytes = kAttnQBytes + kAttnKPBytes + kAttnVBytes + 128;   // x2 CTAs (+1 KB reserved each) fits one SM

struct AttnParams {
    int t, heads, kb, nblocks;
    __nv_bfloat16* out;
    int ld_out;
    int k_col0, v_col0;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// UMMA shared-memory descriptor for an MN-major operand tile stored [k][64 elements] with 128-byte rows and the
// 128B swizzle (exactly what a TMA box of 64 bf16 columns x k rows produces): the 64 MN elements of one k are
// contiguous, 8 consecutive k form a 1024-byte swizzle atom, atoms along k are SBO = 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;                      // LBO: stride between 64-element MN blocks (single block: unused)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO: 8 k-rows
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(uint32_t m, uint32_t n) {
    return umma_idesc_bf16(m, n) | (1u << 16);                // B operand is MN-major
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                 const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if ((smem_base & 1023u) != 0) __trap();  // 128B swizzle atoms need 1024-byte alignment
    const uint32_t s_q = smem_base;
    const uint32_t s_kp = s_q + kAttnQBytes;
    const uint32_t s_v = s_kp + kAttnKPBytes;
    const uint32_t bar_q = s_v + kAttnVBytes;
    const uint32_t bar_kv = bar_q + 8;
    const uint32_t bar_s = bar_q + 16;
    const uint32_t bar_o = bar_q + 24;
    const uint32_t tmem_slot = bar_q + 32;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    uint8_t* p_gen = smem_raw + (s_kp - smem_u32(smem_raw));  // generic pointer to the K/P region

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int mtile = blockIdx.x, head = blockIdx.y, frame = blockIdx.z;
    const int T = p.t, KB = p.kb;

    if (tid == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        mbar_init(bar_q, 1);
        mbar_init(bar_kv, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_o, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<1>(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int q_row0 = frame * T + mtile * 128;
    if (tid == 0) {
        mbar_arrive_expect_tx(bar_q, kAttnQBytes);
        tma_load_2d<1>(&tmap_q, bar_q, s_q, head * 64, q_row0, kEvictFirst);
    }

    const int tok = mtile * 128 + tid;  // this thread's query token
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    constexpr float kLog2e = 1.4426950408889634f;

    float o_acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) o_acc[j] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f;

    const uint32_t r7 = tid & 7;
    uint8_t* p_row = p_gen + (tid >> 3) * 1024 + r7 * 128;
    const int nchunk16 = KB >> 4;

    for (int kb = 0; kb < p.nblocks; ++kb) {
        const int key0 = kb * KB;
        const uint32_t ph = kb & 1;
        const int valid = min(KB, T - key0);      // keys of this frame inside the block (>= 1)
        if (tid == 0) {
            mbar_arrive_expect_tx(bar_kv, 2 * KB * 128);
            tma_load_2d<1>(&tmap_kv, bar_kv, s_kp, p.k_col0 + head * 64, frame * T + key0, kEvictNormal);
            tma_load_2d<1>(&tmap_kv, bar_kv, s_v, p.v_col0 + head * 64, frame * T + key0, kEvictNormal);
            if (kb == 0) mbar_wait(bar_q, 0);
            mbar_wait(bar_kv, ph);
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(128, KB);
            const uint64_t dq = umma_desc_k_sw128(s_q);
            const uint64_t dk = umma_desc_k_sw128(s_kp);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16<1>(tmem_base, dq + 2 * k, dk + 2 * k, idesc, k != 0);
            umma_commit<1>(bar_s);
        }
        mbar_wait(bar_s, ph);
        __syncwarp();
        tc_fence_after();

        // ---- pass 1: row maximum over the valid keys of this block ----
        float m_blk = -INFINITY;
        const int nfull = valid >> 4, rem = valid & 15;
        for (int c = 0; c < nfull; ++c) {
            uint32_t v[16];
            tmem_ld16(t_row + c * 16, v);
            tmem_ld_wait();
            float a = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
            float b = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
            for (int j = 4; j < 16; j += 4) {
                a = fmaxf(a, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
                b = fmaxf(b, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
            }
            m_blk = fmaxf(m_blk, fmaxf(a, b));
        }
        if (rem != 0) {
            uint32_t v[16];
            tmem_ld16(t_row + nfull * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < rem) m_blk = fmaxf(m_blk, __uint_as_float(v[j]));
        }
        const float m_new = fmaxf(m_run, m_blk);
        const float alpha = ex2_approx((m_run - m_new) * kLog2e);  // 0 on the first block (m_run = -inf)
        const uint64_t neg_m2 = pack2(-m_new * kLog2e, -m_new * kLog2e);
        const uint64_t l2e2 = pack2(kLog2e, kLog2e);

        // ---- pass 2: P = exp2(S*log2e - m), row sum, bf16 P into the swizzled K-major smem tile ----
        uint64_t l2 = pack2(0.0f, 0.0f);
        for (int c = 0; c < nchunk16; ++c) {
            uint32_t pk[8];
            if (c < nfull || (c == nfull && rem != 0)) {
                uint32_t v[16];
                tmem_ld16(t_row + c * 16, v);
                tmem_ld_wait();
                float e[16];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float x0, x1;
                    unpack2(fma2(pack2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), l2e2, neg_m2), x0, x1);
                    e[j] = ex2_approx(x0);
                    e[j + 1] = ex2_approx(x1);
                }
                if (c == nfull) {   // partial chunk: keys >= valid belong to another frame
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j >= rem) e[j] = 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    l2 = add2(l2, pack2(e[j], e[j + 1]));
                    pk[j >> 1] = pack_bf16x2(e[j], e[j + 1]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) pk[j] = 0u;
            }
            const int chunk = c >> 2;                       // 64-key smem chunk
            const uint32_t u0 = (static_cast<uint32_t>(c) & 3u) << 1;  // first 16-byte unit inside the 128-byte row
            uint8_t* base = p_row + chunk * (128 * 128);
            *reinterpret_cast<uint4*>(base + ((u0 ^ r7) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(base + (((u0 + 1) ^ r7) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        float l_lo, l_hi;
        unpack2(l2, l_lo, l_hi);
        const float l_blk = l_lo + l_hi;
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        tc_fence_before();
        __syncthreads();

        if (tid == 0) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16_bmn(128, 64);
            for (int ks = 0; ks < nchunk16; ++ks) {
                const uint64_t dp = umma_desc_k_sw128(s_kp + (ks >> 2) * (128 * 128)) + 2 * (ks & 3);
                const uint64_t dv = umma_desc_mn_sw128(s_v + ks * 2048);   // 16 keys = two 8-row atoms
                umma_bf16<1>(tmem_base, dp, dv, idesc, ks != 0);
            }
            umma_commit<1>(bar_o);
        }
        mbar_wait(bar_o, ph);
        __syncwarp();
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
            uint32_t v[16];
            tmem_ld16(t_row + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) o_acc[c + j] = fmaf(o_acc[c + j], alpha, __uint_as_float(v[j]));
        }
        l_run = fmaf(l_run, alpha, l_blk);
        m_run = m_new;
        tc_fence_before();
        __syncthreads();  // TMEM and the K/P region are reused by the next key block
    }

    if (tok < T) {
        const float inv = 1.0f / l_run;
        uint4* o = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(frame * T + tok) * p.ld_out + head * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            o[j] = make_uint4(pack_bf16x2(o_acc[8 * j] * inv, o_acc[8 * j + 1] * inv),
                              pack_bf16x2(o_acc[8 * j + 2] * inv, o_acc[8 * j + 3] * inv),
                              pack_bf16x2(o_acc[8 * j + 4] * inv, o_acc[8 * j + 5] * inv),
                              pack_bf16x2(o_acc[8 * j + 6] * inv, o_acc[8 * j + 7] * inv));
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, 256);
    }
}

int launch_attention(const AttnArgs& a, cudaStream_t stream) {
    CRE_REQUIRE(a.n > 0 && a.t > 0 && a.heads > 0, "attention: empty problem");
    CRE_REQUIRE(a.ld % 8 == 0 && a.k_col0 % 8 == 0 && a.v_col0 % 8 == 0, "attention: ld / column offsets must be multiples of 8");
    const int nblocks = (a.t + 255) / 256;
    int kb = (a.t + nblocks - 1) / nblocks;
    kb = (kb + 15) & ~15;
    CRE_REQUIRE(kb >= 16 && kb <= 256, "attention: key block %d out of range", kb);
    const int64_t rows = static_cast<int64_t>(a.n) * a.t;
    CUtensorMap tq, tkv;
    int rc = make_tmap_bf16(&tq, a.qkv, rows, a.ld, a.ld, 128);
    if (rc) return rc;
    rc = make_tmap_bf16(&tkv, a.qkv, rows, a.ld, a.ld, kb);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        CRE_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
        attr_set = true;
    }
    AttnParams p;
    p.t = a.t;
    p.heads = a.heads;
    p.kb = kb;
    p.nblocks = nblocks;
    p.out = static_cast<__nv_bfloat16*>(a.out);
    p.ld_out = a.heads * 64;
    p.k_col0 = a.k_col0;
    p.v_col0 = a.v_col0;
    dim3 grid((a.t + 127) / 128, a.heads, a.n);
    LaunchScope scope(CRE_K_ATTENTION, 4.0 * a.t * static_cast<double>(a.t) * 64.0 * a.heads * a.n, stream);
    attention_kernel<<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(tq, tkv, p);
    CRE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cre
