// tq_attn_bwd_sm100.cu -- backward of the attention core on the tensor cores (tcgen05 / TMEM / TMA), bf16, head dim 64,
// 32 < T <= 512.  Training-step row; reference: autograd through QKVAttention.forward
// (tqdne/blocks.py:156-190) inside LightningEDM.step (tqdne/edm.py:115-134).
//
//   forward   S = sigma Q K^T (sigma = d^-1/2),  P = softmax(S),  O = P V
//   backward  D_i = sum_c dO_ic O_ic,  dP = dO V^T,  dS = P o (dP - D),  dQ = sigma dS K,  dK = sigma dS^T Q,  dV = P^T dO
//
// Two kernels, no atomics, each GEMM on tcgen05 with fp32 accumulators in tensor memory:
//   attn_bwd_dq_kernel   CTA = (sample, head, 128 queries).  Pass 1 is the forward's S over ALL keys: row max / sum give
//                        the log-sum-exp L_i, written to `lse` together with D_i for the second kernel.  Pass 2 walks
//                        the keys in blocks of 128: S_j and dP_j side by side in tensor memory, thread = (query row, 64
//                        keys) forms dS_j = exp2(S_j c - L_i) (dP_j - D_i) as bf16 into a K-major SWIZZLE_128B tile, and
//                        dQ += dS_j K_j reads K as it lies through an MN-major descriptor.
//   attn_bwd_dkv_kernel  CTA = (sample, head, 128 keys), walks the queries in blocks of 128 with everything transposed:
//                        S^T = K_j Q_i^T and dP^T = V_j dO_i^T, thread = (key row, 64 queries) writes P^T and dS^T tiles,
//                        dV += P^T dO_i and dK += dS^T Q_i (dO_i / Q_i read through MN-major descriptors).
// Gradients leave as bf16 into dqkv [N, T, 3C] with the forward's channel order (third, head, channel).
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include "tq_common.h"
#include "tq_ptx.cuh"

namespace tq {
namespace {

constexpr int D = 64;             // head dimension (the 1D UNet: 4 heads x 64)
constexpr int BLK = 128;          // rows per tile (queries or keys)
constexpr int BWD_THREADS = 256;
constexpr uint32_t TILE = 16384;  // 128 rows x 128 B

struct AttnBwdParams {
    CUtensorMap qkv_map;  // [N][T][3C] bf16, box {64, 128, 1}
    CUtensorMap do_map;   // [N][T][C]  bf16, box {64, 128, 1}
    const __nv_bfloat16* o;    // forward output [N][T][C]
    const __nv_bfloat16* dout; // [N][T][C]
    __nv_bfloat16* dqkv;       // [N][T][3C]
    float* lse;                // [N][heads][T] log2-domain log-sum-exp of the scaled scores
    float* dsum;               // [N][heads][T] D_i
    int N, T, heads, Tk;
    float scale, scale_log2;   // d^-1/2, d^-1/2 * log2(e)
    int lse_given;             // lse[] was written by the forward kernel: pass 1 of the dQ kernel is skipped
};

__device__ __forceinline__ void tma3(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem, uint32_t lbo_bytes = 16) {
    return ((smem & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
// D[128 x 64 cols at tmem] (+)= A[128 x 128, K-major tile pair a_smem (2 slabs of 64 K)] * B[128 K rows x 64, MN-major]
__device__ __forceinline__ void mma_tile_kmajorA_mnmajorB(uint32_t tmem, uint32_t a_smem, uint32_t b_smem, bool accumulate) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, D) | (1u << 16);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem, umma_desc_pack(desc_lo(a_smem + j * TILE) + 2u * kk, DESC_HI),
                      umma_desc_pack(desc_lo(b_smem + (uint32_t)(j * 64 + kk * 16) * 128u), DESC_HI), idesc,
                      accumulate || (j | kk) != 0);
}
// D[128 x n cols] = A[128 x 64, K-major] * B[n x 64, K-major]^T
__device__ __forceinline__ void mma_kmajor(uint32_t tmem, uint32_t a_smem, uint32_t b_smem, int n) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem, umma_desc_pack(desc_lo(a_smem) + 2u * kk, DESC_HI), umma_desc_pack(desc_lo(b_smem) + 2u * kk, DESC_HI),
                  idesc, kk != 0);
}
// thread's 32 values -> bf16 into row `row` of a K-major SWIZZLE_128B [128 x 128] tile pair, columns [col0, col0 + 32)
__device__ __forceinline__ void store_row_32(uint32_t tile_pair, int row, int col0, const float (&v)[32]) {
    const uint32_t base = tile_pair + (uint32_t)(col0 >> 6) * TILE + (uint32_t)row * 128u;
    const uint32_t xr = (uint32_t)(row & 7), ch0 = (uint32_t)((col0 & 63) >> 3);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
        sts16(base + (((ch0 + ch) ^ xr) << 4), pack2(v[8 * ch], v[8 * ch + 1]), pack2(v[8 * ch + 2], v[8 * ch + 3]),
              pack2(v[8 * ch + 4], v[8 * ch + 5]), pack2(v[8 * ch + 6], v[8 * ch + 7]));
}

// ---------------------------------------------------------------------------------------------------- dQ (+ L, D)
__global__ void __launch_bounds__(BWD_THREADS, 1) attn_bwd_dq_kernel(const __grid_constant__ AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float red[2][BLK];
    __shared__ __align__(8) uint64_t bars[3];
    __shared__ uint32_t tmem_slot;
    const int Tk = p.Tk;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = base, do_smem = base + TILE, ds_smem = base + 2 * TILE;   // dS: two slabs
    const uint32_t k_smem = base + 4 * TILE, v_smem = k_smem + (uint32_t)Tk * 128u;
    const uint32_t bar_ld = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qblocks = (p.T + BLK - 1) / BLK;
    const int qb = blockIdx.x % qblocks;
    const int h = (blockIdx.x / qblocks) % p.heads;
    const int n = blockIdx.x / (qblocks * p.heads);
    const int C = p.heads * D;
    if (warp == 1 && lane == 0) {
        mbar_init(bar_ld, 1);
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(&tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
    const int rblocks = Tk / BLK;
    if (warp == 0 && elect_one()) {
        mbar_arrive_expect_tx(bar_ld, (2u + 2u * rblocks) * TILE);
        tma3(q_smem, &p.qkv_map, bar_ld, h * D, qb * BLK, n);
        tma3(do_smem, &p.do_map, bar_ld, h * D, qb * BLK, n);
        for (int b = 0; b < rblocks; ++b) {
            tma3(k_smem + b * TILE, &p.qkv_map, bar_ld, C + h * D, b * BLK, n);
            tma3(v_smem + b * TILE, &p.qkv_map, bar_ld, 2 * C + h * D, b * BLK, n);
        }
    }
    uint32_t mma_phase = 0;
    // ---- pass 1: S over all keys -> row max / sum (skipped when the forward kernel left the log-sum-exp behind)
    if (warp == 0 && !p.lse_given) {
        mbar_wait(bar_ld, 0);
        tc_fence_after();
        if (elect_one()) {
            for (int k0 = 0; k0 < Tk; k0 += 256) mma_kmajor(tmem + k0, q_smem, k_smem + (uint32_t)k0 * 128u, min(256, Tk - k0));
            umma_commit(bar_mma);
        }
        __syncwarp();
    }
    const int row = (warp & 3) * 32 + lane, half = warp >> 2;
    const uint32_t t_row = tmem + (uint32_t((warp & 3) * 32) << 16);
    const int t = qb * BLK + row;
    const float sc = p.scale_log2;
    float Lreg;
    if (p.lse_given) {
        mbar_wait(bar_ld, 0);   // (every thread: the pass-2 MMAs and the D_i loads below need nothing else from pass 1)
        Lreg = t < p.T ? __ldg(p.lse + ((long long)n * p.heads + h) * p.T + t) : 0.f;
    } else {
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    {
        const int cols = Tk / 2, col0 = half * cols;
        float m = -INFINITY;
        for (int c = 0; c < cols; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(t_row + col0 + c, r);
            tmem_ld_wait();
            const int lim = p.T - (col0 + c);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < lim) m = fmaxf(m, __uint_as_float(r[j]));
        }
        red[half][row] = m;
        __syncthreads();
        m = fmaxf(red[0][row], red[1][row]);
        __syncthreads();
        float l = 0.f;
        for (int c = 0; c < cols; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(t_row + col0 + c, r);
            tmem_ld_wait();
            const int lim = p.T - (col0 + c);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < lim) l += ex2f((__uint_as_float(r[j]) - m) * sc);
        }
        red[half][row] = l;
        __syncthreads();
        l = red[0][row] + red[1][row];
        __syncthreads();
        Lreg = m * sc + log2f(l);   // L_i in the log2 domain: P_ij = 2^(s_ij c - L_i); both threads of a row agree bitwise
    }
    }
    // D_i = sum_c dO_ic O_ic (this thread: half of the 64 channels of its row)
    float Dreg;
    {
        float dsum = 0.f;
        if (t < p.T) {
            const long long off = ((long long)n * p.T + t) * C + h * D + half * 32;
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.dout + off + c));
                const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.o + off + c));
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&aw[k]);
                    const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162*>(&bw[k]);
                    dsum = fmaf(__low2float(x), __low2float(y), dsum);
                    dsum = fmaf(__high2float(x), __high2float(y), dsum);
                }
            }
        }
        red[half][row] = dsum;
        tc_fence_before();   // pass 1's tcgen05.ld of S precede the pass-2 MMAs that overwrite those columns
        __syncthreads();
        Dreg = red[0][row] + red[1][row];
    }
    const float L = Lreg, Dm = Dreg;
    if (half == 0 && t < p.T) {
        const long long o = ((long long)n * p.heads + h) * p.T + t;
        if (!p.lse_given) p.lse[o] = L;
        p.dsum[o] = Dm;
    }
    // ---- pass 2: key blocks of 128
    for (int j = 0; j < rblocks; ++j) {
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                mma_kmajor(tmem, q_smem, k_smem + j * TILE, BLK);          // S_j   -> columns [0, 128)
                mma_kmajor(tmem + BLK, do_smem, v_smem + j * TILE, BLK);   // dP_j  -> columns [128, 256)
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
            uint32_t s[32], dp[32];
            tmem_ld_32x32(t_row + half * 64 + c, s);
            tmem_ld_32x32(t_row + BLK + half * 64 + c, dp);
            tmem_ld_wait();
            const int lim = p.T - (j * BLK + half * 64 + c);
            float ds[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const float pij = q < lim ? ex2f(fmaf(__uint_as_float(s[q]), sc, -L)) : 0.f;
                ds[q] = pij * (__uint_as_float(dp[q]) - Dm);
            }
            store_row_32(ds_smem, row, half * 64 + c, ds);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                mma_tile_kmajorA_mnmajorB(tmem + 2 * BLK, ds_smem, k_smem + j * TILE, j != 0);   // dQ += dS_j K_j
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        // the dS tile and the S / dP columns are reused by the next block: wait until this block's MMA has read them
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
    }
    // ---- epilogue: dQ * sigma -> bf16
    {
        uint32_t r[32];
        tmem_ld_32x32(t_row + 2 * BLK + half * 32, r);
        tmem_ld_wait();
        if (t < p.T) {
            __nv_bfloat16* o = p.dqkv + ((long long)n * p.T + t) * (3 * C) + h * D + half * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(o + 8 * q) =
                    make_uint4(pack2(__uint_as_float(r[8 * q]) * p.scale, __uint_as_float(r[8 * q + 1]) * p.scale),
                               pack2(__uint_as_float(r[8 * q + 2]) * p.scale, __uint_as_float(r[8 * q + 3]) * p.scale),
                               pack2(__uint_as_float(r[8 * q + 4]) * p.scale, __uint_as_float(r[8 * q + 5]) * p.scale),
                               pack2(__uint_as_float(r[8 * q + 6]) * p.scale, __uint_as_float(r[8 * q + 7]) * p.scale));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ---------------------------------------------------------------------------------------------------- dK, dV
__global__ void __launch_bounds__(BWD_THREADS, 1) attn_bwd_dkv_kernel(const __grid_constant__ AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float s_lse[BLK], s_dsum[BLK];
    __shared__ __align__(8) uint64_t bars[3];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t k_smem = base, v_smem = base + TILE, q_smem = base + 2 * TILE, do_smem = base + 3 * TILE;
    const uint32_t pt_smem = base + 4 * TILE, dst_smem = base + 6 * TILE;   // P^T, dS^T: two slabs each
    const uint32_t bar_kv = smem_u32(&bars[0]), bar_q = smem_u32(&bars[1]), bar_mma = smem_u32(&bars[2]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blocks = (p.T + BLK - 1) / BLK;
    const int kb = blockIdx.x % blocks;
    const int h = (blockIdx.x / blocks) % p.heads;
    const int n = blockIdx.x / (blocks * p.heads);
    const int C = p.heads * D;
    if (warp == 1 && lane == 0) {
        mbar_init(bar_kv, 1);
        mbar_init(bar_q, 1);
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(&tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
    if (warp == 0 && elect_one()) {
        mbar_arrive_expect_tx(bar_kv, 2u * TILE);
        tma3(k_smem, &p.qkv_map, bar_kv, C + h * D, kb * BLK, n);
        tma3(v_smem, &p.qkv_map, bar_kv, 2 * C + h * D, kb * BLK, n);
    }
    const int row = (warp & 3) * 32 + lane, half = warp >> 2;   // row = key, half = which 64 queries of the block
    const uint32_t t_row = tmem + (uint32_t((warp & 3) * 32) << 16);
    const int key = kb * BLK + row;
    const float sc = p.scale_log2;
    uint32_t mma_phase = 0, q_phase = 0;
    for (int i = 0; i < blocks; ++i) {
        if (warp == 0) {
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_q, 2u * TILE);
                tma3(q_smem, &p.qkv_map, bar_q, h * D, i * BLK, n);
                tma3(do_smem, &p.do_map, bar_q, h * D, i * BLK, n);
            }
            __syncwarp();
        }
        if (threadIdx.x < BLK) {
            const int tq = i * BLK + threadIdx.x;
            const long long o = ((long long)n * p.heads + h) * p.T + tq;
            s_lse[threadIdx.x] = tq < p.T ? p.lse[o] : INFINITY;   // queries past the end: P = 0
            s_dsum[threadIdx.x] = tq < p.T ? p.dsum[o] : 0.f;
        }
        if (warp == 0) {
            if (i == 0) mbar_wait(bar_kv, 0);
            mbar_wait(bar_q, q_phase);
            tc_fence_after();
            if (elect_one()) {
                mma_kmajor(tmem, k_smem, q_smem, BLK);            // S^T  = K_j Q_i^T  -> columns [0, 128)
                mma_kmajor(tmem + BLK, v_smem, do_smem, BLK);     // dP^T = V_j dO_i^T -> columns [128, 256)
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        q_phase ^= 1;
        __syncthreads();   // s_lse / s_dsum visible
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
            uint32_t s[32], dp[32];
            tmem_ld_32x32(t_row + half * 64 + c, s);
            tmem_ld_32x32(t_row + BLK + half * 64 + c, dp);
            tmem_ld_wait();
            float pt[32], dst[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int qi = half * 64 + c + q;
                const float pij = key < p.T ? ex2f(fmaf(__uint_as_float(s[q]), sc, -s_lse[qi])) : 0.f;
                pt[q] = pij;
                dst[q] = pij * (__uint_as_float(dp[q]) - s_dsum[qi]);
            }
            store_row_32(pt_smem, row, half * 64 + c, pt);
            store_row_32(dst_smem, row, half * 64 + c, dst);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                mma_tile_kmajorA_mnmajorB(tmem + 2 * BLK, pt_smem, do_smem, i != 0);        // dV += P^T dO_i
                mma_tile_kmajorA_mnmajorB(tmem + 2 * BLK + D, dst_smem, q_smem, i != 0);    // dK += dS^T Q_i
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        mbar_wait(bar_mma, mma_phase);   // Q_i / dO_i / the P^T, dS^T tiles and the S^T, dP^T columns are free again
        mma_phase ^= 1;
        tc_fence_after();
    }
    // ---- epilogue: thread = (key row, half): half 0 writes dV, half 1 writes sigma dK (64 channels each, 2 x 32 columns)
    {
        const float mul = half == 0 ? 1.f : p.scale;
        __nv_bfloat16* o = p.dqkv + ((long long)n * p.T + key) * (3 * C) + (half == 0 ? 2 * C : C) + h * D;
#pragma unroll
        for (int c = 0; c < D; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(t_row + 2 * BLK + half * D + c, r);
            tmem_ld_wait();
            if (key < p.T) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(o + c + 8 * q) =
                        make_uint4(pack2(__uint_as_float(r[8 * q]) * mul, __uint_as_float(r[8 * q + 1]) * mul),
                                   pack2(__uint_as_float(r[8 * q + 2]) * mul, __uint_as_float(r[8 * q + 3]) * mul),
                                   pack2(__uint_as_float(r[8 * q + 4]) * mul, __uint_as_float(r[8 * q + 5]) * mul),
                                   pack2(__uint_as_float(r[8 * q + 6]) * mul, __uint_as_float(r[8 * q + 7]) * mul));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int encode_rows(CUtensorMap* m, const void* ptr, int N, int T, int ld, const char* what) {
    static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    if (!enc) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
    }
    TQ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)T, (cuuint64_t)N};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TQ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return 0;
}

}  // namespace
}  // namespace tq

using namespace tq;

extern "C" int tq_attention_backward(const void* qkv, const void* out, const void* dout, void* dqkv, float* ws, int32_t N,
                                     int32_t T, int32_t heads, int32_t d, int32_t lse_given, void* stream) {
    TQ_CHECK(qkv && out && dout && dqkv && ws, "attention_backward: null pointer");
    TQ_CHECK(d == D, "attention_backward: head dim 64 only (the 1D UNet); got %d", d);
    TQ_CHECK(N > 0 && heads > 0 && T > 32 && T <= 512, "attention_backward: 32 < T <= 512");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AttnBwdParams p;
    const int C = heads * D;
    if (encode_rows(&p.qkv_map, qkv, N, T, 3 * C, "qkv")) return 1;
    if (encode_rows(&p.do_map, dout, N, T, C, "dout")) return 1;
    p.o = static_cast<const __nv_bfloat16*>(out);
    p.dout = static_cast<const __nv_bfloat16*>(dout);
    p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
    p.lse = ws;
    p.dsum = ws + (size_t)N * heads * T;
    p.N = N; p.T = T; p.heads = heads; p.Tk = (T + BLK - 1) / BLK * BLK;
    p.lse_given = lse_given ? 1 : 0;
    p.scale = 1.f / sqrtf((float)D);
    p.scale_log2 = p.scale * 1.4426950408889634f;
    const int blocks = (T + BLK - 1) / BLK;
    const size_t smem_q = 1024 + 4 * TILE + 2 * (size_t)p.Tk * 128;
    const size_t smem_kv = 1024 + 8 * TILE;
    static PerDeviceOnce attr;
    if (attr.first()) {
        TQ_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 4 * TILE + 2 * 512 * 128));
        TQ_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv));
    }
    attn_bwd_dq_kernel<<<N * heads * blocks, BWD_THREADS, smem_q, st>>>(p);
    TQ_CUDA(cudaGetLastError());
    attn_bwd_dkv_kernel<<<N * heads * blocks, BWD_THREADS, smem_kv, st>>>(p);
    TQ_CUDA(cudaGetLastError());
    count_launch(2);
    return 0;
}
