// tq_wgrad2d_sm100.cu -- weight gradient of a stride-1 "same" 2-D convolution on the tensor cores (tcgen05 / TMEM / TMA).
// Groundwork for training the 2-D (latent / pixel) EDM UNets and the autoencoder (reference:
// loss.backward() through nn.Conv2d of tqdne/nn.py:16-24).  Channels-last bf16 activations [N, H, W, C]:
//
//     dW[co][ky][kx][ci] = sum over (n, y, x) of dY[n][y][x][co] * X[n][y + ky - kh/2][x + kx - kw/2][ci]   (fp32)
//
// The same GEMM as tq_wgrad_sm100.cu (K = positions, both operands MN-major), with two differences forced by two
// dimensions: a K-step is a box of bw x bh x bn = 64 positions (TMA 4-D boxes; out-of-range rows / columns / samples are
// zero-filled), and every tap loads its own shifted X box -- an x-shift of a flattened row-major tile wraps around the
// image rows, so the 1-D kernel's "one halo buffer, tap = row offset" does not carry over.  kh * kw accumulators of 64
// TMEM columns would not fit 512 columns for 3 x 3, so the taps are split into groups of <= 5 over the grid.
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include "tq_common.h"
#include "tq_ptx.cuh"

namespace tq {
namespace {

constexpr int W2_THREADS = 192;
constexpr int W2_KSTEP = 64;
constexpr int W2_STAGES = 3;
constexpr int W2_A = 2 * W2_KSTEP * 128;   // dY: two 64-channel slabs
constexpr int W2_B = W2_KSTEP * 128;       // X: one 64-channel slab per tap
constexpr int W2_MAXT = 5;                 // taps per CTA (5 x 64 TMEM columns)

struct Wgrad2dParams {
    CUtensorMap ymap;  // dY [N][H][W][Cout], box {64, bw, bh, bn}
    CUtensorMap xmap;  // X  [N][H][W][Cin],  box {64, bw, bh, bn}
    float* dw;         // [Cout][kh*kw][Cin]
    int N, H, W, cin, cout, kh, kw, taps;
    int bw, bh, bn, tiles_x, tiles_y, tiles_n, total_steps, steps_per_cta, kchunks;
    int co_tiles, ci_tiles, tap_groups, taps_per_group;
};

__device__ __forceinline__ void tma4(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__global__ void __launch_bounds__(W2_THREADS, 1) wgrad2d_kernel(const __grid_constant__ Wgrad2dParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * W2_STAGES + 1];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t stage_bytes = W2_A + W2_MAXT * W2_B;
    auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
    auto empty_bar = [&](int s) { return smem_u32(&bars[W2_STAGES + s]); };
    const uint32_t done_bar = smem_u32(&bars[2 * W2_STAGES]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work unit: blockIdx = (kchunk, tap group, ci tile, co tile)
    int u = blockIdx.x;
    const int kc = u % p.kchunks; u /= p.kchunks;
    const int tg = u % p.tap_groups; u /= p.tap_groups;
    const int ci_t = u % p.ci_tiles;
    const int co_t = u / p.ci_tiles;
    const int t_begin = tg * p.taps_per_group;
    const int ntaps = min(p.taps_per_group, p.taps - t_begin);
    const int s_begin = kc * p.steps_per_cta;
    const int s_end = min(p.total_steps, s_begin + p.steps_per_cta);
    const int nsteps = s_end - s_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.ymap);
        tma_prefetch_desc(&p.xmap);
        for (int s = 0; s < W2_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);

    if (nsteps > 0 && ntaps > 0) {
        if (warp == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int s = s_begin; s < s_end; ++s) {
                const int tx = s % p.tiles_x;
                const int r = s / p.tiles_x;
                const int ty = r % p.tiles_y;
                const int tn = r / p.tiles_y;
                const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = tn * p.bn;
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    const uint32_t dst = base + stage * stage_bytes;
                    mbar_arrive_expect_tx(full_bar(stage), W2_A + ntaps * W2_B);
                    tma4(dst, &p.ymap, full_bar(stage), co_t * 128, x0, y0, n0);
                    tma4(dst + W2_KSTEP * 128, &p.ymap, full_bar(stage), co_t * 128 + 64, x0, y0, n0);
                    for (int t = 0; t < ntaps; ++t) {
                        const int tap = t_begin + t;
                        const int ky = tap / p.kw, kx = tap % p.kw;
                        tma4(dst + W2_A + t * W2_B, &p.xmap, full_bar(stage), ci_t * 64, x0 + kx - p.kw / 2, y0 + ky - p.kh / 2, n0);
                    }
                }
                __syncwarp();
                if (++stage == W2_STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        } else if (warp == 1) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);   // A and B MN-major
            constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            int stage = 0;
            uint32_t phase = 0;
            for (int s = 0; s < nsteps; ++s) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_smem = base + stage * stage_bytes;
                    const uint32_t a_lo0 = ((a_smem & 0x3FFFFu) >> 4) | (((uint32_t)(W2_KSTEP * 128) >> 4) << 16);
                    for (int t = 0; t < ntaps; ++t) {
                        const uint32_t b_lo0 = (((a_smem + W2_A + t * W2_B) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
                        for (int kk = 0; kk < W2_KSTEP / 16; ++kk) {
                            const uint32_t off = (uint32_t)(kk * 16 * 128) >> 4;
                            umma_bf16(tmem_base + t * 64, umma_desc_pack(a_lo0 + off, desc_hi), umma_desc_pack(b_lo0 + off, desc_hi),
                                      idesc, (s | kk) != 0);
                        }
                    }
                    umma_commit(empty_bar(stage));
                    if (s == nsteps - 1) umma_commit(done_bar);
                }
                __syncwarp();
                if (++stage == W2_STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        } else {
            const int q = warp & 3;
            mbar_wait(done_bar, 0);
            tc_fence_after();
            const int co = co_t * 128 + q * 32 + lane;
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
            for (int t = 0; t < ntaps; ++t) {
#pragma unroll
                for (int c = 0; c < 64; c += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_row + t * 64 + c, r);
                    tmem_ld_wait();
                    if (co < p.cout) {
                        float* dst = p.dw + ((long long)co * p.taps + t_begin + t) * p.cin + ci_t * 64 + c;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
                                         "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])),
                                         "f"(__uint_as_float(r[j + 3]))
                                         : "memory");
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int encode_nhwc(CUtensorMap* m, const void* ptr, int N, int H, int W, int Cc, int bw, int bh, int bn, const char* what) {
    static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    if (!enc) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
    }
    TQ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    cuuint64_t dims[4] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)H * W * Cc * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TQ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return 0;
}

int pow2_le(int v, int cap) {
    int p = 1;
    while (p * 2 <= v && p * 2 <= cap) p *= 2;
    return p;
}

}  // namespace
}  // namespace tq

using namespace tq;

extern "C" int tq_conv2d_wgrad(const void* x, const void* dy, float* dw, int32_t N, int32_t H, int32_t W, int32_t cin, int32_t cout,
                               int32_t kh, int32_t kw, void* stream) {
    TQ_CHECK(x && dy && dw && N > 0 && H > 0 && W > 0, "conv2d_wgrad: bad arguments");
    TQ_CHECK(kh >= 1 && kw >= 1 && (kh & 1) && (kw & 1) && kh * kw <= 25, "conv2d_wgrad: odd kernel sizes up to 5 x 5");
    TQ_CHECK(cin % 64 == 0 && cout % 64 == 0, "conv2d_wgrad: channel counts must be multiples of 64 (pad the stem / head)");
    TQ_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0, "conv2d_wgrad: 16 B alignment");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Wgrad2dParams p;
    p.dw = dw; p.N = N; p.H = H; p.W = W; p.cin = cin; p.cout = cout; p.kh = kh; p.kw = kw; p.taps = kh * kw;
    // a K-step = bw x bh x bn = 64 positions (powers of two; partial boxes at the edges are zero-filled by TMA)
    p.bw = pow2_le(W, W2_KSTEP);
    p.bh = pow2_le(H, W2_KSTEP / p.bw);
    p.bn = W2_KSTEP / (p.bw * p.bh);
    p.tiles_x = (W + p.bw - 1) / p.bw;
    p.tiles_y = (H + p.bh - 1) / p.bh;
    p.tiles_n = (N + p.bn - 1) / p.bn;
    p.total_steps = p.tiles_x * p.tiles_y * p.tiles_n;
    p.co_tiles = (cout + 127) / 128;
    p.ci_tiles = cin / 64;
    p.tap_groups = (p.taps + W2_MAXT - 1) / W2_MAXT;
    p.taps_per_group = (p.taps + p.tap_groups - 1) / p.tap_groups;
    const int tiles = p.co_tiles * p.ci_tiles * p.tap_groups;
    int kchunks = device_sm_count() / tiles;
    if (kchunks > p.total_steps) kchunks = p.total_steps;
    if (kchunks < 1) kchunks = 1;
    p.steps_per_cta = (p.total_steps + kchunks - 1) / kchunks;
    p.kchunks = (p.total_steps + p.steps_per_cta - 1) / p.steps_per_cta;
    if (encode_nhwc(&p.ymap, dy, N, H, W, cout, p.bw, p.bh, p.bn, "dY")) return 1;
    if (encode_nhwc(&p.xmap, x, N, H, W, cin, p.bw, p.bh, p.bn, "X")) return 1;
    const size_t smem = 1024 + (size_t)W2_STAGES * (W2_A + W2_MAXT * W2_B);
    static PerDeviceOnce attr;
    if (attr.first()) {
        TQ_CUDA(cudaFuncSetAttribute(wgrad2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    wgrad2d_kernel<<<tiles * p.kchunks, W2_THREADS, smem, st>>>(p);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}
