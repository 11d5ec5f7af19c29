// tq_wgrad_sm100.cu -- weight gradient of a stride-1 "same" 1-D convolution on the tensor cores (tcgen05 / TMEM / TMA).
// First kernel of the training-step row (reference: loss.backward() through nn.Conv1d of
// tqdne/nn.py:16-24 in LightningEDM.step, tqdne/edm.py:115-134).  Channels-last bf16 activations as everywhere else:
//
//     dW[co][t][ci] = sum over (n, l) of dY[n][l][co] * X[n][l + t - pad][ci]            (fp32 accumulate, fp32 output)
//     db[co]        = sum over (n, l) of dY[n][l][co]
//
// i.e. per tap t a GEMM with M = Cout, N = Cin and K = N * L POSITIONS.  Both operands are MN-major in this layout
// (channels contiguous, positions = K along the shared-memory rows), the descriptor form tq_attn_sm100.cu uses for V.
// The taps need no extra loads: a tile's K-step is 64 positions of dY [64 x 128 co] plus ONE halo buffer of X
// [64 + k - 1 positions x 64 ci] (TMA zero-fills positions outside [0, L)), and tap t is the same buffer with the
// descriptor start advanced by t rows (a SWIZZLE_128B operand may start at any 128 B row: tools/umma_rowshift_probe.cu).
//
// One CTA = (128 output channels, 64 input channels, a chunk of the K steps): k accumulators of 64 TMEM columns each,
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue (tcgen05.ld -> red.global.add.f32 into dW: the K
// chunks of a tile are summed through fp32 atomics, so dW / db must be zeroed by the caller).
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include <algorithm>

#include <memory>

#include "tq_common.h"
#include "tq_ptx.cuh"

namespace tq {
namespace {

constexpr int WG_THREADS = 192;
constexpr int KSTEP = 64;        // positions per pipeline stage
constexpr int WG_STAGES = 6;
constexpr int A_STAGE = 2 * KSTEP * 128;  // dY: two 64-channel slabs of 64 rows x 128 B
constexpr int MAX_TAPS = 7;

struct WgradParams {
    CUtensorMap ymap;  // dY as [N][L][Cout] bf16, box {64, 64, 1}
    CUtensorMap xmap;  // X  as [N][L][Cin]  bf16, box {64, 64 + taps - 1 rounded up to 8, 1}
    float* dw;         // [Cout][taps][Cin]
    float* db;         // [Cout] or nullptr
    int N, L, cin, cout, taps, pad;
    int dw_ld, ci_off;  // dW row length (input channels of the WHOLE layer) and this source's first channel in it
    int b_rows;        // rows of the X halo box
    int b_stage;       // bytes of the X halo buffer rounded up to 1024
    int steps_per_sample, total_steps, steps_per_cta, kchunks, co_tiles, ci_tiles;
    int probe;         // experiments (TQ_WGRAD_PROBE): 1 = the epilogue reads the accumulators but skips the global reductions
};

__device__ __forceinline__ void tma_load_3d_wg(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad1d_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * WG_STAGES + 1];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = A_STAGE + p.b_stage;
    auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
    auto empty_bar = [&](int s) { return smem_u32(&bars[WG_STAGES + s]); };
    const uint32_t done_bar = smem_u32(&bars[2 * WG_STAGES]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work unit: blockIdx = (kchunk, ci_tile, co_tile)
    int u = blockIdx.x;
    const int kc = u % p.kchunks; u /= p.kchunks;
    const int ci_t = u % p.ci_tiles;
    const int co_t = u / p.ci_tiles;
    const int s_begin = kc * p.steps_per_cta;
    const int s_end = min(p.total_steps, s_begin + p.steps_per_cta);
    const int nsteps = s_end - s_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.ymap);
        tma_prefetch_desc(&p.xmap);
        for (int s = 0; s < WG_STAGES; ++s) {
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

    if (nsteps > 0) {
        if (warp == 0) {
            // ---------------------------------------------------------------- TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int s = s_begin; s < s_end; ++s) {
                const int n = s / p.steps_per_sample;
                const int l0 = (s % p.steps_per_sample) * KSTEP;
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    const uint32_t dst = base + stage * stage_bytes;
                    mbar_arrive_expect_tx(full_bar(stage), A_STAGE + p.b_rows * 128);
                    tma_load_3d_wg(dst, &p.ymap, full_bar(stage), co_t * 128, l0, n);
                    tma_load_3d_wg(dst + KSTEP * 128, &p.ymap, full_bar(stage), co_t * 128 + 64, l0, n);
                    tma_load_3d_wg(dst + A_STAGE, &p.xmap, full_bar(stage), ci_t * 64, l0 - p.pad, n);
                }
                __syncwarp();
                if (++stage == WG_STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        } else if (warp == 1) {
            // ---------------------------------------------------------------- MMA issuer
            // A = dY^T: M = 128 output channels in two 64-channel slabs (LBO), K = positions: 8-row groups 1024 B apart
            // B = X^T : N = 64 input channels (one slab), same K structure; both MN-major
            constexpr uint32_t idesc = umma_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
            constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            int stage = 0;
            uint32_t phase = 0;
            for (int s = 0; s < nsteps; ++s) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_smem = base + stage * stage_bytes;
                    const uint32_t a_lo0 = ((a_smem & 0x3FFFFu) >> 4) | (((uint32_t)(KSTEP * 128) >> 4) << 16);
                    const uint32_t b_lo0 = (((a_smem + A_STAGE) & 0x3FFFFu) >> 4) | (1u << 16);
                    for (int t = 0; t < p.taps; ++t) {
#pragma unroll
                        for (int kk = 0; kk < KSTEP / 16; ++kk) {
                            const uint32_t a_lo = a_lo0 + ((uint32_t)(kk * 16 * 128) >> 4);
                            const uint32_t b_lo = b_lo0 + ((uint32_t)((kk * 16 + t) * 128) >> 4);
                            umma_bf16(tmem_base + t * 64, umma_desc_pack(a_lo, desc_hi), umma_desc_pack(b_lo, desc_hi), idesc,
                                      (s | kk) != 0);
                        }
                    }
                    umma_commit(empty_bar(stage));
                    if (s == nsteps - 1) umma_commit(done_bar);
                }
                __syncwarp();
                if (++stage == WG_STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        } else {
            // ---------------------------------------------------------------- epilogue: warps 2..5 own TMEM lanes 32*(warp%4)
            const int q = warp & 3;
            mbar_wait(done_bar, 0);
            tc_fence_after();
            const int co = co_t * 128 + q * 32 + lane;
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
            for (int t = 0; t < p.taps; ++t) {
#pragma unroll
                for (int c = 0; c < 64; c += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_row + t * 64 + c, r);
                    tmem_ld_wait();
                    if (co < p.cout && !(p.probe & 1)) {
                        // cin is a multiple of 64: the 32 columns are in range and 16 B aligned -> 8 vector reductions
                        float* dst = p.dw + ((long long)co * p.taps + t) * p.dw_ld + p.ci_off + ci_t * 64 + c;
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

// db[co] = sum over positions of dY.  Thread = (8-channel vector of a slab of <= 32 vectors, row lane), 16 B loads, four
// rows in flight; the grid is a fixed ~2 blocks per SM walking the rows with a stride, because what bounded the version
// this replaces (one block per 256 rows) was its ONE fp32 atomic per channel and block: ~2 000 serialised atomics per address.
constexpr int BG_VEC = 32;   // vectors (of 8 channels) per block slab
__global__ void __launch_bounds__(256) bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, long long rows, int cout,
                                                        float* __restrict__ db) {
    __shared__ float red[256 * 9];
    const int cv = cout >> 3;                       // 16 B vectors per row (cout % 8 == 0)
    const int v0 = blockIdx.y * BG_VEC, cvb = min(BG_VEC, cv - v0);
    const int lanes = 256 / cvb;                    // >= 8 row lanes
    const int vi = threadIdx.x % cvb, rl = threadIdx.x / cvb;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto add8 = [&](const uint4& u) {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
            a[2 * k] += __low2float(b2);
            a[2 * k + 1] += __high2float(b2);
        }
    };
    if (rl < lanes) {
        const __nv_bfloat16* base = dy + (long long)(v0 + vi) * 8;
        const long long step = (long long)gridDim.x * lanes;
        long long r = (long long)blockIdx.x * lanes + rl;
        for (; r + 3 * step < rows; r += 4 * step) {
            uint4 u[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) u[q] = __ldg(reinterpret_cast<const uint4*>(base + (r + q * step) * cout));
#pragma unroll
            for (int q = 0; q < 4; ++q) add8(u[q]);
        }
        for (; r < rows; r += step) add8(__ldg(reinterpret_cast<const uint4*>(base + r * cout)));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x * 9 + j] = rl < lanes ? a[j] : 0.f;
    __syncthreads();
    for (int o = threadIdx.x; o < cvb * 8; o += 256) {
        const int v2 = o >> 3, j = o & 7;
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += red[(l * cvb + v2) * 9 + j];
        atomicAdd(db + (v0 + v2) * 8 + j, t);
    }
}

// out[n][c] += sum over the P positions of sample n of dy[n][p][c]: the gradient of a per-sample, per-channel additive
// term (the timestep / conditioning embedding added after a ResBlock's first convolution, tqdne/unet.py:129-141)
__global__ void __launch_bounds__(256) sample_channel_sum_kernel(const __nv_bfloat16* __restrict__ dy, int P, int C,
                                                                 float* __restrict__ out, int out_ld) {
    const int n = blockIdx.y;
    const int c = blockIdx.x * 64 + (threadIdx.x & 63);
    const int rl = threadIdx.x >> 6;
    float a = 0.f;
    if (c < C) {
        const __nv_bfloat16* b = dy + (long long)n * P * C + c;
        for (int r = rl; r < P; r += 4) a += __bfloat162float(b[(long long)r * C]);
    }
    __shared__ float red[4][64];
    red[rl][threadIdx.x & 63] = a;
    __syncthreads();
    if (rl == 0 && c < C)
        out[(long long)n * out_ld + c] += red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
}

PFN_cuTensorMapEncodeTiled_v12000 wg_encode_fn() {
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

int encode_nlc(CUtensorMap* m, const void* ptr, int N, int L, int Cc, int box_rows, const char* what) {
    auto enc = wg_encode_fn();
    TQ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    cuuint64_t dims[3] = {(cuuint64_t)Cc, (cuuint64_t)L, (cuuint64_t)N};
    cuuint64_t strides[2] = {(cuuint64_t)Cc * 2, (cuuint64_t)L * Cc * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TQ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return 0;
}

}  // namespace
}  // namespace tq

using namespace tq;

extern "C" int tq_conv1d_wgrad(const void* x, const void* dy, float* dw, float* db, int32_t N, int64_t L, int32_t cin,
                               int32_t cout, int32_t taps, int32_t dw_ld, int32_t ci_off, void* stream) {
    if (dw_ld <= 0) dw_ld = cin;
    TQ_CHECK(ci_off >= 0 && ci_off % 4 == 0 && ci_off + cin <= dw_ld && dw_ld % 4 == 0, "conv1d_wgrad: bad dw_ld / ci_off");
    TQ_CHECK(x && dy && dw && N > 0 && L > 0, "conv1d_wgrad: bad arguments");
    TQ_CHECK(taps >= 1 && taps <= MAX_TAPS && (taps & 1), "conv1d_wgrad: odd kernel sizes up to %d", MAX_TAPS);
    TQ_CHECK(cin % 64 == 0 && cout % 64 == 0, "conv1d_wgrad: channel counts must be multiples of 64 (pad the stem / head)");
    TQ_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0, "conv1d_wgrad: 16 B alignment");
    TQ_CHECK(L < (1ll << 30), "conv1d_wgrad: sequence too long");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    WgradParams p;
    p.dw = dw; p.db = db; p.N = N; p.L = (int)L; p.cin = cin; p.cout = cout; p.taps = taps; p.pad = taps / 2; p.dw_ld = dw_ld; p.ci_off = ci_off;
    p.b_rows = (KSTEP + taps - 1 + 7) / 8 * 8;
    p.b_stage = (p.b_rows * 128 + 1023) / 1024 * 1024;
    p.steps_per_sample = (int)((L + KSTEP - 1) / KSTEP);
    p.total_steps = p.steps_per_sample * N;
    p.co_tiles = (cout + 127) / 128;
    p.ci_tiles = cin / 64;
    const int tiles = p.co_tiles * p.ci_tiles;
    p.probe = 0;
    if (const char* e = getenv("TQ_WGRAD_PROBE")) p.probe = atoi(e);
    int kchunks = device_sm_count() / tiles;  // one wave of CTAs: every extra K chunk adds a full tile of fp32 reductions
    if (kchunks > p.total_steps) kchunks = p.total_steps;
    if (kchunks < 1) kchunks = 1;
    p.steps_per_cta = (p.total_steps + kchunks - 1) / kchunks;
    p.kchunks = (p.total_steps + p.steps_per_cta - 1) / p.steps_per_cta;
    if (encode_nlc(&p.ymap, dy, N, (int)L, cout, KSTEP, "dY")) return 1;
    if (encode_nlc(&p.xmap, x, N, (int)L, cin, p.b_rows, "X")) return 1;
    const size_t smem = 1024 + (size_t)WG_STAGES * (A_STAGE + p.b_stage);
    static PerDeviceMax attr;
    if (attr.raise(smem)) TQ_CUDA(cudaFuncSetAttribute(wgrad1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad1d_kernel<<<tiles * p.kchunks, WG_THREADS, smem, st>>>(p);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    if (db) {
        const long long rows = (long long)N * L;
        TQ_CHECK(cout % 8 == 0 && cout <= 2048, "conv1d_wgrad: bias gradient needs cout %% 8 == 0 and cout <= 2048");
        const int cv = cout / 8, slabs = (cv + BG_VEC - 1) / BG_VEC;
        const int gx = std::max(1, std::min((int)((rows + 63) / 64), 2 * device_sm_count() / slabs));
        bias_grad_kernel<<<dim3((unsigned)gx, (unsigned)slabs), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(dy), rows, cout, db);
        TQ_CUDA(cudaGetLastError());
        count_launch();
    }
    return 0;
}

extern "C" int tq_sample_channel_sums(const void* dy, float* out, int32_t out_ld, int32_t N, int64_t P, int32_t C, void* stream) {
    if (out_ld <= 0) out_ld = C;
    TQ_CHECK(dy && out && N > 0 && P > 0 && C > 0 && P < (1ll << 31), "sample_channel_sums: bad arguments");
    sample_channel_sum_kernel<<<dim3((C + 63) / 64, N), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dy), (int)P, C, out, out_ld);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}
