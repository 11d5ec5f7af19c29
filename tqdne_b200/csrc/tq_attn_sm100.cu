// tq_attn_sm100.cu -- attention core on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), bf16, 32 < T <= 512.
// Reference: QKVAttention.forward (tqdne/blocks.py:156-190): q,k,v = qkv.chunk(3, dim=1), heads split inside each
// third, w = softmax_fp32((q*s)^T (k*s)) with s = d^-1/4, a = w v.  No mask here (a causal attention runs on tq_attn.cu).
// Three kernels: attention_tc_multi_kernel (128 < T <= 512: the pixel-space 2D UNet, T = 256, d = 128, and the 1D UNet,
// T = 508 / 512, d = 64 -- K / V resident over a range of query blocks, P kept in tensor memory, optional log-sum-exp output
// for the training step), attention_tc_kernel (32 < T <= 128 and shapes the multi-block kernel cannot hold) and
// attention_tc_packed_kernel (T in {16, 32}: the latent UNet).  The fp32 parity mode stays on the FFMA kernels (tcgen05 has
// no fp32 MMA).
//
// attention_tc_kernel: one CTA = (sample, head, 128 queries), 256 threads.  All keys of the head are resident, so there is no online
// softmax and no rescaling of the accumulator:
//   TMA      Q [128 x d], K [Tk x d], V [Tk x d] boxes of 128 rows x 64 channels (SWIZZLE_128B) straight out of the
//            channels-last qkv tensor (rows >= T are zero-filled by the TMA unit), Tk = T rounded up to 128
//   MMA 1    S[128 x Tk] = Q K^T   (both operands K-major), fp32 in Tk tensor-memory columns
//   softmax  thread = (query row, half of the keys): row max, p = 2^((s - m) * d^-1/2 * log2 e), bf16 P written into
//            the (now dead) K buffer in the K-major SWIZZLE_128B layout the second MMA reads
//   MMA 2    O[128 x d] = P V      (P K-major; V is read as it lies, keys x channels, through an MN-major
//            descriptor), accumulated over the S columns that the softmax has finished reading
//   epilogue O / row sum -> bf16 -> out[n, t, head*d + c]
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <memory>

#include "tq_common.h"
#include "tq_ptx.cuh"

namespace tq {
namespace {

constexpr int kQB = 128;        // queries per CTA (= TMEM lanes)
constexpr int kThreadsTc = 256;

struct AttnTcParams {
    CUtensorMap map;   // qkv as [N][T][3C] bf16, box {64, 128, 1}
    __nv_bfloat16* out;
    int N, T, heads, Tk;
    float scale_log2;  // d^-1/2 * log2(e)
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int D>
__global__ void __launch_bounds__(kThreadsTc, 1) attention_tc_kernel(const __grid_constant__ AttnTcParams p) {
    constexpr int DC = D / 64;                 // 64-channel slabs per operand
    constexpr uint32_t Q_BYTES = DC * 16384u;  // 128 rows x 128 B per slab
    extern __shared__ uint8_t smem_raw[];
    __shared__ float red_max[2][kQB];
    __shared__ float red_sum[2][kQB];
    __shared__ __align__(8) uint64_t bars[3];
    __shared__ uint32_t tmem_slot;

    const int Tk = p.Tk;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = base;
    const uint32_t kp_smem = q_smem + Q_BYTES;                 // K [DC][Tk x 128 B], later P [Tk/64][128 x 128 B]
    const uint32_t v_smem = kp_smem + (uint32_t)Tk * 256u;     // V [DC][Tk x 128 B]
    const uint32_t slab = (uint32_t)Tk * 128u;                 // one 64-channel slab of K or V
    const uint32_t bar_qk = smem_u32(&bars[0]), bar_v = smem_u32(&bars[1]), bar_mma = smem_u32(&bars[2]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qblocks = (p.T + kQB - 1) / kQB;
    const int qb = blockIdx.x % qblocks;
    const int h = (blockIdx.x / qblocks) % p.heads;
    const int n = blockIdx.x / (qblocks * p.heads);
    const int C = p.heads * D;
    const uint32_t tmem_cols = Tk <= 128 ? 128u : (Tk <= 256 ? 256u : 512u);

    pdl_launch_dependents();
    if (warp == 1 && lane == 0) {
        tma_prefetch_desc(&p.map);
        mbar_init(bar_qk, 1);
        mbar_init(bar_v, 1);
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
    pdl_wait();

    const int rblocks = Tk / 128;
    if (warp == 0) {
        if (elect_one()) {
            // ---- loads: Q + K on one barrier (needed first), V on its own
            mbar_arrive_expect_tx(bar_qk, Q_BYTES + (uint32_t)rblocks * DC * 16384u);
            for (int c = 0; c < DC; ++c) tma_load_3d(q_smem + c * 16384u, &p.map, bar_qk, h * D + 64 * c, qb * kQB, n);
            for (int c = 0; c < DC; ++c)
                for (int b = 0; b < rblocks; ++b)
                    tma_load_3d(kp_smem + c * slab + b * 16384u, &p.map, bar_qk, C + h * D + 64 * c, b * 128, n);
            mbar_arrive_expect_tx(bar_v, (uint32_t)rblocks * DC * 16384u);
            for (int c = 0; c < DC; ++c)
                for (int b = 0; b < rblocks; ++b)
                    tma_load_3d(v_smem + c * slab + b * 16384u, &p.map, bar_v, 2 * C + h * D + 64 * c, b * 128, n);
        }
        __syncwarp();
        // ---- MMA 1: S = Q K^T, at most 256 keys (N) per instruction
        mbar_wait(bar_qk, 0);
        tc_fence_after();
        if (elect_one()) {
            constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B
            const uint32_t q_lo = ((q_smem & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t k_lo = ((kp_smem & 0x3FFFFu) >> 4) | (1u << 16);
            for (int k0 = 0; k0 < Tk; k0 += 256) {
                const int nk = min(256, Tk - k0);
                const uint32_t idesc = umma_idesc_bf16(kQB, nk);
#pragma unroll
                for (int c = 0; c < DC; ++c) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t a_lo = q_lo + ((c * 16384u) >> 4) + 2u * kk;
                        const uint32_t b_lo = k_lo + ((c * slab + (uint32_t)k0 * 128u) >> 4) + 2u * kk;
                        umma_bf16(tmem_base + k0, umma_desc_pack(a_lo, desc_hi), umma_desc_pack(b_lo, desc_hi), idesc,
                                  (c | kk) != 0);
                    }
                }
            }
            umma_commit(bar_mma);
        }
        __syncwarp();
    }

    // ---- softmax over the resident score row; thread = (row, half of the key columns)
    const int row = (warp & 3) * 32 + lane;
    const int half = warp >> 2;
    const uint32_t t_row = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    const int cols = Tk / 2;
    const int col0 = half * cols;
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    float m = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < cols; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + col0 + c, r);
        tmem_ld_wait();
        const int lim = p.T - (col0 + c);  // columns >= T are padding keys
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < lim) m = fmaxf(m, __uint_as_float(r[j]));
    }
    red_max[half][row] = m;
    __syncthreads();
    m = fmaxf(red_max[0][row], red_max[1][row]);
    const float sc = p.scale_log2;
    const float msc = m * sc;
    float l = 0.f;
    const uint32_t p_row = kp_smem + (uint32_t)row * 128u;
    const uint32_t xr = (uint32_t)(row & 7);
#pragma unroll 1
    for (int c = 0; c < cols; c += 32) {
        uint32_t r[32];
        const int col = col0 + c;
        tmem_ld_32x32(t_row + col, r);
        tmem_ld_wait();
        const int lim = p.T - col;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float e0 = (2 * j < lim) ? ex2_approx(fmaf(__uint_as_float(r[2 * j]), sc, -msc)) : 0.f;
            const float e1 = (2 * j + 1 < lim) ? ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), sc, -msc)) : 0.f;
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
            l += __low2float(b2) + __high2float(b2);  // the row sum of the weights the MMA really applies
            pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        const uint32_t sl = p_row + (uint32_t)(col >> 6) * 16384u;  // 64-key slab of P
        const uint32_t ch = (uint32_t)((col & 63) >> 3);             // first 16 B chunk inside the 128 B row
#pragma unroll
        for (int q = 0; q < 4; ++q)
            sts128(sl + (((ch + q) ^ xr) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
    red_sum[half][row] = l;
    fence_proxy_async();  // P was written through the generic proxy; the MMA reads it through the async proxy
    tc_fence_before();    // orders this thread's tcgen05.ld of S before the MMA that overwrites those columns
    __syncthreads();

    // ---- MMA 2: O = P V into columns [0, D)
    if (warp == 0) {
        mbar_wait(bar_v, 0);
        tc_fence_after();
        if (elect_one()) {
            constexpr uint32_t p_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t p_lo = ((kp_smem & 0x3FFFFu) >> 4) | (1u << 16);
            // V [keys][64 channels] per slab: channels (N) contiguous = MN-major; 8-key groups 1024 B apart (SBO),
            // the next 64 channels one slab further (LBO)
            const uint32_t v_lo = ((v_smem & 0x3FFFFu) >> 4) | ((slab >> 4) << 16);
            constexpr uint32_t idesc = umma_idesc_bf16(kQB, D) | (1u << 16);  // B operand MN-major
            const int kslabs = Tk / 64;
#pragma unroll 1
            for (int j = 0; j < kslabs; ++j) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t a_lo = p_lo + ((j * 16384u) >> 4) + 2u * kk;
                    const uint32_t b_lo = v_lo + (((uint32_t)(j * 64 + kk * 16) * 128u) >> 4);
                    umma_bf16(tmem_base, umma_desc_pack(a_lo, p_hi), umma_desc_pack(b_lo, p_hi), idesc, (j | kk) != 0);
                }
            }
            umma_commit(bar_mma);
        }
        __syncwarp();
    }
    mbar_wait(bar_mma, 1);
    tc_fence_after();
    {
        const float inv = 1.f / (red_sum[0][row] + red_sum[1][row]);
        const int t = qb * kQB + row;
        constexpr int OC = D / 2;  // output channels per thread
        __nv_bfloat16* o = p.out + ((long long)n * p.T + t) * C + h * D + half * OC;
#pragma unroll
        for (int c = 0; c < OC; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(t_row + half * OC + c, r);
            tmem_ld_wait();
            if (t < p.T) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(r[8 * q + 2 * j]) * inv,
                                                                        __uint_as_float(r[8 * q + 2 * j + 1]) * inv);
                        w[j] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    *reinterpret_cast<uint4*>(o + c + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-block variant (the default for 128 < T <= 512): one CTA = (sample, head, a RANGE of query blocks).  K and V are
// loaded ONCE and stay resident while the CTA walks its query blocks (the kernel above reloads both -- 128 KB at T = 508 --
// for every block of 128 queries), Q is double buffered, and P never touches shared memory: the softmax writes it back
// into tensor memory as packed bf16 over the score columns it has consumed and the second MMA takes its A operand from
// there (tcgen05.mma with A in TMEM), which removes a 128 KB shared-memory write + read per query block.
//   TMEM columns (Tk = keys rounded up to 128):  S fp32 [0, Tk);  P bf16x2 in place, keys [0, Tk/2) -> [0, Tk/4) and keys
//   [Tk/2, Tk) -> [Tk/2, 3 Tk/4) (each half of the threads stays inside its own half of the row);  O fp32 at [Tk, Tk + D)
//   when Tk <= 256, else at [Tk/4, Tk/4 + D) -- free once the first half of S has been consumed.
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows = lanes, 16 bf16 of K = 8 columns) comes from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

struct AttnMultiParams {
    CUtensorMap map;   // qkv as [N][T][3C] bf16, box {64, 128, 1}
    __nv_bfloat16* out;
    int N, T, heads, Tk;
    int qsplit;        // CTAs per (sample, head): each takes a contiguous range of the query blocks
    float scale_log2;
    float* lse;        // optional [N][heads][T]: log2-domain log-sum-exp of the scaled scores, for the backward pass
};

template <int D>
__global__ void __launch_bounds__(kThreadsTc, 1) attention_tc_multi_kernel(const __grid_constant__ AttnMultiParams p) {
    constexpr int DC = D / 64;
    constexpr uint32_t Q_BYTES = DC * 16384u;
    extern __shared__ uint8_t smem_raw[];
    __shared__ float red_max[2][kQB];
    __shared__ float red_sum[2][kQB];
    __shared__ float red_lx[2][kQB];   // unrounded row sums (only when the log-sum-exp is written out)
    __shared__ __align__(8) uint64_t bars[5];   // q[0], q[1], k, v, mma
    __shared__ uint32_t tmem_slot;

    const int Tk = p.Tk;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = base;                                  // two Q buffers
    const uint32_t k_smem = q_smem + 2u * Q_BYTES;                 // K [DC][Tk x 128 B]
    const uint32_t slab = (uint32_t)Tk * 128u;
    const uint32_t v_smem = k_smem + (uint32_t)DC * slab;          // V [DC][Tk x 128 B]
    const uint32_t bar_q0 = smem_u32(&bars[0]), bar_k = smem_u32(&bars[2]), bar_v = smem_u32(&bars[3]), bar_mma = smem_u32(&bars[4]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qblocks = (p.T + kQB - 1) / kQB;
    const int part = blockIdx.x % p.qsplit;
    const int h = (blockIdx.x / p.qsplit) % p.heads;
    const int n = blockIdx.x / (p.qsplit * p.heads);
    const int per = (qblocks + p.qsplit - 1) / p.qsplit;
    const int qb_begin = part * per, qb_end = min(qblocks, qb_begin + per);
    const int C = p.heads * D;
    const uint32_t o_col = Tk <= 256 ? (uint32_t)Tk : (uint32_t)Tk / 4;
    const uint32_t tmem_cols = (Tk <= 256 && Tk + D <= 256) ? 256u : 512u;

    pdl_launch_dependents();
    if (warp == 1 && lane == 0) {
        tma_prefetch_desc(&p.map);
        for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&bars[i]), 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
    pdl_wait();
    if (qb_begin >= qb_end) {   // (more CTAs than query blocks)
        tc_fence_before();
        __syncthreads();
        if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
        return;
    }

    const int rblocks = Tk / 128;
    auto load_q = [&](int qb) {   // elected thread of warp 0
        const int b = (qb - qb_begin) & 1;
        mbar_arrive_expect_tx(bar_q0 + 8u * b, Q_BYTES);
        for (int c = 0; c < DC; ++c) tma_load_3d(q_smem + b * Q_BYTES + c * 16384u, &p.map, bar_q0 + 8u * b, h * D + 64 * c, qb * kQB, n);
    };
    if (warp == 0 && elect_one()) {
        load_q(qb_begin);
        mbar_arrive_expect_tx(bar_k, (uint32_t)rblocks * DC * 16384u);
        for (int c = 0; c < DC; ++c)
            for (int b = 0; b < rblocks; ++b)
                tma_load_3d(k_smem + c * slab + b * 16384u, &p.map, bar_k, C + h * D + 64 * c, b * 128, n);
        if (qb_begin + 1 < qb_end) load_q(qb_begin + 1);
        mbar_arrive_expect_tx(bar_v, (uint32_t)rblocks * DC * 16384u);
        for (int c = 0; c < DC; ++c)
            for (int b = 0; b < rblocks; ++b)
                tma_load_3d(v_smem + c * slab + b * 16384u, &p.map, bar_v, 2 * C + h * D + 64 * c, b * 128, n);
    }
    __syncwarp();

    const int row = (warp & 3) * 32 + lane;
    const int half = warp >> 2;
    const uint32_t t_row = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    const int cols = Tk / 2;
    const int col0 = half * cols;
    const float sc = p.scale_log2;
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B

    for (int qb = qb_begin; qb < qb_end; ++qb) {
        const int it = qb - qb_begin;
        const int qbuf = it & 1;
        // ---- MMA 1: S = Q K^T
        if (warp == 0) {
            mbar_wait(bar_q0 + 8u * qbuf, (uint32_t)(it >> 1) & 1u);
            if (it == 0) mbar_wait(bar_k, 0);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t q_lo = (((q_smem + qbuf * Q_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
                const uint32_t k_lo = ((k_smem & 0x3FFFFu) >> 4) | (1u << 16);
                for (int k0 = 0; k0 < Tk; k0 += 256) {
                    const int nk = min(256, Tk - k0);
                    const uint32_t idesc = umma_idesc_bf16(kQB, nk);
#pragma unroll
                    for (int c = 0; c < DC; ++c) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint32_t a_lo = q_lo + ((c * 16384u) >> 4) + 2u * kk;
                            const uint32_t b_lo = k_lo + ((c * slab + (uint32_t)k0 * 128u) >> 4) + 2u * kk;
                            umma_bf16(tmem_base + k0, umma_desc_pack(a_lo, desc_hi), umma_desc_pack(b_lo, desc_hi), idesc,
                                      (c | kk) != 0);
                        }
                    }
                }
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        mbar_wait(bar_mma, 0);   // two completions per query block: parities 0, 1
        tc_fence_after();
        if (warp == 0 && qb + 2 < qb_end && elect_one()) load_q(qb + 2);   // this block's Q buffer is free again
        // ---- softmax over the resident score row; thread = (row, half of the key columns)
        // The tensor-memory loads are software pipelined (chunk c + 1 is in flight while chunk c is consumed): with two
        // warps per scheduler nothing else hides their latency.  Only the chunk holding key T needs the padding mask.
        float m = -INFINITY;
        uint32_t ra[32], rb[32];
        auto max_chunk = [&](const uint32_t (&r)[32], int c) {
            const int lim = p.T - (col0 + c);  // columns >= T are padding keys
            if (lim >= 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < lim) m = fmaxf(m, __uint_as_float(r[j]));
            }
        };
        tmem_ld_32x32(t_row + col0, ra);
        tmem_ld_wait();
#pragma unroll 1
        for (int c = 0; c < cols; c += 64) {
            tmem_ld_32x32(t_row + col0 + c + 32, rb);
            max_chunk(ra, c);
            tmem_ld_wait();
            if (c + 64 < cols) tmem_ld_32x32(t_row + col0 + c + 64, ra);
            max_chunk(rb, c + 32);
            tmem_ld_wait();
        }
        red_max[half][row] = m;
        __syncthreads();
        m = fmaxf(red_max[0][row], red_max[1][row]);
        const float msc = m * sc;
        float l = 0.f, lx = 0.f;   // lx: the same sum before the bf16 rounding (what the backward's exact softmax needs)
        const bool want_lse = p.lse != nullptr;
        auto exp_chunk = [&](const uint32_t (&r)[32], int c) {
            const int lim = p.T - (col0 + c);
            uint32_t pk[16];
            if (lim >= 32) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float e0 = ex2_approx(fmaf(__uint_as_float(r[2 * j]), sc, -msc));
                    const float e1 = ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), sc, -msc));
                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
                    l += __low2float(b2) + __high2float(b2);  // the row sum of the weights the MMA really applies
                    if (want_lse) lx += e0 + e1;
                    pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float e0 = (2 * j < lim) ? ex2_approx(fmaf(__uint_as_float(r[2 * j]), sc, -msc)) : 0.f;
                    const float e1 = (2 * j + 1 < lim) ? ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), sc, -msc)) : 0.f;
                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
                    l += __low2float(b2) + __high2float(b2);
                    if (want_lse) lx += e0 + e1;
                    pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
                }
            }
            // packed P over the score columns this thread has already consumed: [col0 + c/2, col0 + c/2 + 16)
            tmem_st_32x16(t_row + col0 + (c >> 1), pk);
        };
        tmem_ld_32x32(t_row + col0, ra);
        tmem_ld_wait();
#pragma unroll 1
        for (int c = 0; c < cols; c += 64) {
            tmem_ld_32x32(t_row + col0 + c + 32, rb);
            exp_chunk(ra, c);
            tmem_ld_wait();
            if (c + 64 < cols) tmem_ld_32x32(t_row + col0 + c + 64, ra);
            exp_chunk(rb, c + 32);
            tmem_ld_wait();
        }
        red_sum[half][row] = l;
        if (want_lse) red_lx[half][row] = lx;
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 2: O = P V, A from tensor memory
        if (warp == 0) {
            if (it == 0) mbar_wait(bar_v, 0);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t v_lo = ((v_smem & 0x3FFFFu) >> 4) | ((slab >> 4) << 16);
                constexpr uint32_t idesc = umma_idesc_bf16(kQB, D) | (1u << 16);  // B operand MN-major
                const int ksteps = Tk / 16;
#pragma unroll 1
                for (int ks = 0; ks < ksteps; ++ks) {
                    const int key0 = ks * 16;
                    const uint32_t pcol = key0 < cols ? (uint32_t)(key0 >> 1) : (uint32_t)(cols + ((key0 - cols) >> 1));
                    const uint32_t b_lo = v_lo + (((uint32_t)key0 * 128u) >> 4);
                    umma_bf16_ts(tmem_base + o_col, tmem_base + pcol, umma_desc_pack(b_lo, desc_hi), idesc, ks != 0);
                }
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        mbar_wait(bar_mma, 1);
        tc_fence_after();
        {
            const float inv = 1.f / (red_sum[0][row] + red_sum[1][row]);
            const int t = qb * kQB + row;
            if (want_lse && half == 0 && t < p.T)   // L_i = m c + log2 sum_j 2^((s_ij - m) c): P_ij = 2^(s_ij c - L_i)
                p.lse[((long long)n * p.heads + h) * p.T + t] = msc + log2f(red_lx[0][row] + red_lx[1][row]);
            constexpr int OC = D / 2;  // output channels per thread
            __nv_bfloat16* o = p.out + ((long long)n * p.T + t) * C + h * D + half * OC;
#pragma unroll
            for (int c = 0; c < OC; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32(t_row + o_col + half * OC + c, r);
                tmem_ld_wait();
                if (t < p.T) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(r[8 * q + 2 * j]) * inv,
                                                                            __uint_as_float(r[8 * q + 2 * j + 1]) * inv);
                            w[j] = *reinterpret_cast<const uint32_t*>(&b2);
                        }
                        *reinterpret_cast<uint4*>(o + c + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();   // every thread has read O (and red_sum) before the next block's S overwrites the columns
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// T = 16 or 32 (latent UNet: 4 x 4 = 16 tokens): 128 / T (sample, head) pairs share one CTA and ONE pair of 128-row
// MMAs.  S[128 x 128] = Q_packed K_packed^T holds the T x T score blocks of the pairs on its diagonal (the
// off-diagonal products are computed and ignored: the tensor pipe is idle anyway and the kernel is bound by the
// 3 x 128 rows it loads); P is written block-diagonal, so O = P V_packed is exact.
struct AttnPackParams {
    CUtensorMap map;   // qkv as [N][T][3C] bf16, box {64, T, 1}
    __nv_bfloat16* out;
    int N, heads, pairs;  // pairs = N * heads
    float scale_log2;
};

template <int D, int T>
__global__ void __launch_bounds__(128) attention_tc_packed_kernel(const __grid_constant__ AttnPackParams p) {
    constexpr int DC = D / 64;
    constexpr int PP = 128 / T;                  // (sample, head) pairs per CTA
    constexpr uint32_t OPB = DC * 16384u;        // one packed operand: DC slabs of 128 rows x 128 B
    constexpr uint32_t BOX = (uint32_t)T * 128u;  // one TMA box: T rows x 64 channels
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[3];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = base;
    const uint32_t kp_smem = q_smem + OPB;       // K, later P [2 slabs][128 x 128 B]
    const uint32_t v_smem = kp_smem + 32768u;
    const uint32_t bar_qk = smem_u32(&bars[0]), bar_v = smem_u32(&bars[1]), bar_mma = smem_u32(&bars[2]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.heads * D;
    constexpr uint32_t tmem_cols = D > 128 ? 256u : 128u;

    pdl_launch_dependents();
    if (warp == 1 && lane == 0) {
        tma_prefetch_desc(&p.map);
        mbar_init(bar_qk, 1);
        mbar_init(bar_v, 1);
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
    pdl_wait();

    const int pair0 = blockIdx.x * PP;
    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_qk, 2u * OPB);
            for (int which = 0; which < 3; ++which) {
                if (which == 2) mbar_arrive_expect_tx(bar_v, OPB);
                const uint32_t dst0 = which == 0 ? q_smem : (which == 1 ? kp_smem : v_smem);
                const uint32_t bar = which == 2 ? bar_v : bar_qk;
                for (int j = 0; j < PP; ++j) {
                    const int pr = min(pair0 + j, p.pairs - 1);  // a ragged last CTA re-loads a valid pair (never stored)
                    const int n = pr / p.heads, h = pr % p.heads;
#pragma unroll
                    for (int c = 0; c < DC; ++c)
                        tma_load_3d(dst0 + c * 16384u + j * BOX, &p.map, bar, which * C + h * D + 64 * c, 0, n);
                }
            }
        }
        __syncwarp();
        mbar_wait(bar_qk, 0);
        tc_fence_after();
        if (elect_one()) {
            constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t q_lo = ((q_smem & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t k_lo = ((kp_smem & 0x3FFFFu) >> 4) | (1u << 16);
            constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
#pragma unroll
            for (int c = 0; c < DC; ++c)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem_base, umma_desc_pack(q_lo + ((c * 16384u) >> 4) + 2u * kk, desc_hi),
                              umma_desc_pack(k_lo + ((c * 16384u) >> 4) + 2u * kk, desc_hi), idesc, (c | kk) != 0);
            umma_commit(bar_mma);
        }
        __syncwarp();
    }

    // ---- softmax of this row's T x T diagonal block
    const int row = warp * 32 + lane;
    const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16);
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    float sv[T];
    {
        uint32_t r[32];
        tmem_ld_32x32(t_row + warp * 32, r);   // the 32 key columns of this warp's 32 / T pairs
        tmem_ld_wait();
        if constexpr (T == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sv[j] = __uint_as_float(r[j]);
        } else {
            const bool hi = lane >= 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) sv[j] = __uint_as_float(hi ? r[16 + j] : r[j]);
        }
    }
    float m = sv[0];
#pragma unroll
    for (int j = 1; j < T; ++j) m = fmaxf(m, sv[j]);
    const float sc = p.scale_log2, msc = m * sc;
    float l = 0.f;
    uint32_t pk[T / 2];
#pragma unroll
    for (int j = 0; j < T / 2; ++j) {
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(ex2_approx(fmaf(sv[2 * j], sc, -msc)), ex2_approx(fmaf(sv[2 * j + 1], sc, -msc)));
        l += __low2float(b2) + __high2float(b2);
        pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
    }
    {
        // block-diagonal P row: 16 chunks of 8 keys, non-zero only for this pair's keys [k0, k0 + T)
        const uint32_t p_row = kp_smem + (uint32_t)row * 128u;
        const uint32_t xr = (uint32_t)(row & 7);
        const int ch0 = (row / T) * (T / 8);
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
            const uint32_t addr = p_row + (uint32_t)(ch >> 3) * 16384u + ((((uint32_t)ch & 7u) ^ xr) << 4);
            const int q = ch - ch0;
            uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
            for (int qq = 0; qq < T / 8; ++qq)
                if (q == qq) { w0 = pk[4 * qq]; w1 = pk[4 * qq + 1]; w2 = pk[4 * qq + 2]; w3 = pk[4 * qq + 3]; }
            sts128(addr, w0, w1, w2, w3);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    if (warp == 0) {
        mbar_wait(bar_v, 0);
        tc_fence_after();
        if (elect_one()) {
            constexpr uint32_t p_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t p_lo = ((kp_smem & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t v_lo = ((v_smem & 0x3FFFFu) >> 4) | ((16384u >> 4) << 16);
            constexpr uint32_t idesc = umma_idesc_bf16(128, D) | (1u << 16);  // B operand (V) MN-major
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem_base, umma_desc_pack(p_lo + ((j * 16384u) >> 4) + 2u * kk, p_hi),
                              umma_desc_pack(v_lo + (((uint32_t)(j * 64 + kk * 16) * 128u) >> 4), p_hi), idesc, (j | kk) != 0);
            umma_commit(bar_mma);
        }
        __syncwarp();
    }
    mbar_wait(bar_mma, 1);
    tc_fence_after();
    {
        const float inv = 1.f / l;
        const int pr = pair0 + row / T, t = row % T;
        const int n = pr / p.heads, h = pr % p.heads;
        __nv_bfloat16* o = p.out + ((long long)n * T + t) * C + h * D;
#pragma unroll
        for (int c = 0; c < D; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(t_row + c, r);
            tmem_ld_wait();
            if (pr < p.pairs) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(r[8 * q + 2 * j]) * inv,
                                                                        __uint_as_float(r[8 * q + 2 * j + 1]) * inv);
                        w[j] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    *reinterpret_cast<uint4*>(o + c + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
    return fn;
}

size_t attn_tc_smem(int Tk, int d) { return 1024 + (size_t)(d / 64) * 16384 + (size_t)Tk * 256 + (size_t)Tk * d * 2; }

template <int D>
int launch_attn_tc(const AttnTcParams& p, cudaStream_t st) {
    const size_t smem = attn_tc_smem(p.Tk, D);
    static PerDeviceMax attr_smem;
    if (attr_smem.raise(smem))
        TQ_CUDA(cudaFuncSetAttribute(attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int qblocks = (p.T + kQB - 1) / kQB;
    TQ_CUDA(launch_pdl(attention_tc_kernel<D>, dim3(p.N * p.heads * qblocks), dim3(kThreadsTc), smem, st, p));
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

size_t attn_multi_smem(int Tk, int d) { return 1024 + 2 * (size_t)(d / 64) * 16384 + 2 * (size_t)Tk * d * 2; }
// the multi-block kernel needs room for O beside S / P in tensor memory: Tk + D <= 512, or D <= Tk / 4 above 256 keys
bool attn_multi_ok(int Tk, int d) {
    const char* e = getenv("TQ_ATTN_MULTI");
    if (e && e[0] == '0') return false;
    if (Tk <= 128) return false;                       // one query block: nothing to share
    if (Tk > 256 && d > Tk / 4) return false;
    return attn_multi_smem(Tk, d) <= 224 * 1024;
}

template <int D>
int launch_attn_multi(const AttnMultiParams& p, cudaStream_t st) {
    const size_t smem = attn_multi_smem(p.Tk, D);
    static PerDeviceMax attr_smem;
    if (attr_smem.raise(smem))
        TQ_CUDA(cudaFuncSetAttribute(attention_tc_multi_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TQ_CUDA(launch_pdl(attention_tc_multi_kernel<D>, dim3(p.N * p.heads * p.qsplit), dim3(kThreadsTc), smem, st, p));
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

template <int D, int T>
int launch_attn_packed(const AttnPackParams& p, cudaStream_t st) {
    constexpr size_t smem = 1024 + (size_t)(D / 64) * 16384 * 2 + 32768;
    static PerDeviceOnce attr;
    if (attr.first()) {
        TQ_CUDA(cudaFuncSetAttribute(attention_tc_packed_kernel<D, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    constexpr int PP = 128 / T;
    TQ_CUDA(launch_pdl(attention_tc_packed_kernel<D, T>, dim3((p.pairs + PP - 1) / PP), dim3(128), smem, st, p));
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace

bool attention_tc_supported(const tq_attn_desc& d) {
    if (d.dtype != TQ_BF16 || (d.d != 64 && d.d != 128) || d.T > 512) return false;
    if ((reinterpret_cast<uintptr_t>(d.qkv) & 15) != 0 || (reinterpret_cast<uintptr_t>(d.out) & 15) != 0) return false;
    if (d.T <= 32) return d.T == 16 || d.T == 32;  // packed kernel: whole (sample, head) pairs per 128-row tile
    const int Tk = (d.T + 127) / 128 * 128;
    return attn_tc_smem(Tk, d.d) <= 224 * 1024;  // + ~2 KB of static shared memory
}

// does the kernel build_attention picks for `d` write tq_attn_desc.lse?  (only the multi-block kernel does)
bool attention_writes_lse(const tq_attn_desc& d) {
    const char* simt = getenv("TQ_ATTN_SIMT");
    if (d.causal || (simt && simt[0] == '1') || !attention_tc_supported(d) || d.T <= 32) return false;
    return attn_multi_ok((d.T + 127) / 128 * 128, d.d);
}

int build_attention_tc(std::vector<Op>& ops, const tq_attn_desc& d) {
    auto enc = encode_fn();
    TQ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    auto p = std::make_shared<AttnTcParams>();
    const int C = d.heads * d.d;
    p->out = static_cast<__nv_bfloat16*>(d.out);
    p->N = d.N; p->T = d.T; p->heads = d.heads;
    p->Tk = (d.T + 127) / 128 * 128;
    p->scale_log2 = 1.4426950408889634f / sqrtf((float)d.d);
    cuuint64_t dims[3] = {(cuuint64_t)(3 * C), (cuuint64_t)d.T, (cuuint64_t)d.N};
    cuuint64_t strides[2] = {(cuuint64_t)(3 * C) * 2, (cuuint64_t)d.T * (3 * C) * 2};
    const bool packed = d.T <= 32;
    cuuint32_t box[3] = {64, packed ? (cuuint32_t)d.T : 128u, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&p->map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d.qkv), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TQ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(qkv) failed with CUresult %d", (int)r);
    const int dd = d.d;
    Op op;
    char nm[64];
    if (packed) {
        auto q = std::make_shared<AttnPackParams>();
        q->map = p->map; q->out = p->out; q->N = d.N; q->heads = d.heads; q->pairs = d.N * d.heads;
        q->scale_log2 = p->scale_log2;
        const int tt = d.T;
        snprintf(nm, sizeof nm, "attention_tc_packed<bf16,d=%d> T=%d", dd, d.T);
        op.name = nm;
        op.small = true;
        op.launch = [q, dd, tt](cudaStream_t st) -> int {
            if (dd == 128) return tt == 16 ? launch_attn_packed<128, 16>(*q, st) : launch_attn_packed<128, 32>(*q, st);
            return tt == 16 ? launch_attn_packed<64, 16>(*q, st) : launch_attn_packed<64, 32>(*q, st);
        };
        ops.push_back(std::move(op));
        return 0;
    }
    if (attn_multi_ok(p->Tk, dd)) {
        auto q = std::make_shared<AttnMultiParams>();
        q->map = p->map; q->out = p->out; q->N = d.N; q->T = d.T; q->heads = d.heads; q->Tk = p->Tk;
        q->scale_log2 = p->scale_log2;
        q->lse = d.lse;
        // CTAs per (sample, head): all query blocks in one CTA (K / V loaded once) unless that leaves SMs idle
        const int qblocks = (d.T + kQB - 1) / kQB;
        int qsplit = 1;
        while (qsplit < qblocks && (long long)d.N * d.heads * qsplit < device_sm_count()) qsplit *= 2;
        if (const char* e = getenv("TQ_ATTN_QSPLIT")) qsplit = std::max(1, std::min(qblocks, atoi(e)));
        q->qsplit = qsplit;
        snprintf(nm, sizeof nm, "attention_tc_multi<bf16,d=%d> T=%d", dd, d.T);
        op.name = nm;
        op.launch = [q, dd](cudaStream_t st) -> int {
            return dd == 128 ? launch_attn_multi<128>(*q, st) : launch_attn_multi<64>(*q, st);
        };
        ops.push_back(std::move(op));
        return 0;
    }
    snprintf(nm, sizeof nm, "attention_tc<bf16,d=%d> T=%d", dd, d.T);
    op.name = nm;
    op.launch = [p, dd](cudaStream_t st) -> int {
        return dd == 128 ? launch_attn_tc<128>(*p, st) : launch_attn_tc<64>(*p, st);
    };
    ops.push_back(std::move(op));
    return 0;
}

}  // namespace tq
