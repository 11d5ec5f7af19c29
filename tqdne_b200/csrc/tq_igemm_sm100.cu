// tq_igemm_sm100.cu -- bf16 implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel covers every dense contraction of the denoiser and decoder
// (reference: nn.Conv1d/Conv2d from tqdne/nn.py:16-24 at the call sites listed in include/tqdne_b200.h):
//
//   D[128*CG positions, BN channels] += A_slice[128*CG, 64] * W_slice[BN, 64]^T   for every K-slice
//
// CG = 2 runs a CTA PAIR (cluster of two SMs) on one 256 x BN tile with tcgen05.mma.cta_group::2: each CTA
// stages its own 128 activation rows and only HALF of the weight rows, so the shared-memory fill and the
// operand reads per SM drop by a quarter to a third compared with two independent 128 x BN tiles -- the
// resource that bounds this kernel (the operand tiles are streamed from L2 for every tap).
//
//   warp 0   TMA producer : per slice one 4-D box load of the shifted input window (zero fill outside the
//                           image = "same" padding) + one 2-D box load of the weight block, both landing in
//                           128B-swizzled shared memory; transaction bytes of both CTAs complete on the
//                           leader's mbarrier
//   warp 1   MMA issuer   : (leader CTA) one thread issues tcgen05.mma (M=128*CG, N=BN, K=16) x4 per slice
//                           into a TMEM accumulator; tcgen05.commit (multicast to the pair) releases the
//                           smem stage / publishes the tile
//   warp 2   TMEM allocator (2*BN columns: the accumulator is double buffered so the epilogue of tile i
//                           overlaps the MMAs of tile i+1)
//   warps 4-11 epilogue   : 8 warps = 4 TMEM lane quarters x 2 column halves.  bf16 outputs go through a
//                           per-warp 32-row x 64-channel staging buffer in shared memory: the residual
//                           arrives there by TMA (issued before the accumulator is ready), the warp adds
//                           bias / per-sample embedding / residual to its tcgen05.ld rows in place, and the
//                           result leaves by TMA store (coalesced, clipped at the tensor edge, asynchronous).
//                           The GroupNorm statistics of the consumer (per-(sample, tile, channel) sum and sum of
//                           squares) are column sums over the same staging buffers, written with plain stores into a
//                           slot only this tile owns (no atomics: bit-reproducible); samples that span several
//                           epilogue warps of a tile are summed across the four staging buffers of a column half
//                           behind a 128-thread named barrier.
//                           fp32 outputs (embedding GEMM, the two ragged Cout in {3, 8} convs) use
//                           direct stores.
//
// Activations are channels-last, so the 128x64 A tile of a slice is the TMA box
// {64 ch, bw, bh, bn} of the [N,H,W,C] tensor with bw*bh*bn = 128: rows of 128 B, K-major,
// exactly the canonical SWIZZLE_128B UMMA operand layout.  No im2col buffer exists anywhere.
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <memory>
#include <vector>

#include "tq_common.h"
#include "tq_ptx.cuh"

namespace tq {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;  // 16 KiB
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int EPI_BUF_BYTES = 32 * 64 * 2;  // one epilogue unit: 32 rows x 64 bf16 channels
constexpr int MAX_SMEM = 232448;            // 227 KiB opt-in limit per CTA

struct alignas(64) IgemmParams {
    CUtensorMap amap[4];
    CUtensorMap bmap;
    CUtensorMap omap[4];  // bf16 output per parity class, box = one epilogue unit
    CUtensorMap rmap;     // residual, same geometry as omap[0]
    CUtensorMap amap3[4];  // ROW3 stages: the same sources with a box of bh + 2 rows
    const int4* slices;    // stage records, 3 x int4 each, [num_classes][num_stages]
    int num_slices, num_classes;
    int num_stages;        // pipeline stages per tile (a stage = up to KS plain slices, or one ROW3 group)
    int a3_bytes;          // bytes of a ROW3 activation buffer: (bh + 2) * bw * 128
    int row_bytes;         // bw * 128: one image row of a tile inside a ROW3 buffer
    int N, H, W, bw, bh, bn;
    int tiles_x, tiles_y, m_tiles, m_groups, n_tiles, total_tiles;
    int cout;
    int vec_ok;   // fp32 direct-store path only
    int has_res;  // residual comes in through rmap (bf16 path)
    const float* bias;
    const float* emb;
    int emb_ld;
    const __nv_bfloat16* residual;
    void* out;
    long long out_sn, out_sy, out_sx;
    long long out_class_off[4];
    float* stats;  // [N][parts][cout][2] per-(sample, tile, channel) sum / sum of squares, or nullptr
    int parts;     // num_classes * tiles_x * tiles_y: 128-row tiles per sample
    int wps;       // epilogue warps (32 rows each) that share one sample inside a tile: 1, 2 or 4
    int seg;       // rows of one sample inside an epilogue warp: min(32, bw*bh)
    unsigned long long* prof;  // debug (TQ_IGEMM_PROF=1): [grid][16] cycle counters, else nullptr
    int probe;  // debug (TQ_IGEMM_PROBE bit mask, results are garbage): 1 = no TMA operand loads, 2 = no MMAs,
                // 4 = epilogue only hands the accumulator back
};

__device__ __forceinline__ long long clk() { return clock64(); }
// mbarrier wait that adds the cycles spent waiting to `acc` when profiling is on
__device__ __forceinline__ void mbar_wait_prof(uint32_t bar, uint32_t parity, bool on, long long& acc) {
    if (!on) {
        mbar_wait(bar, parity);
        return;
    }
    const long long t0 = clk();
    mbar_wait(bar, parity);
    acc += clk() - t0;
}

template <int BN, int CG>
struct Cfg {
    static constexpr int BNC = BN / CG;  // weight rows staged per CTA
    static constexpr int B_BYTES = BNC * BK * 2;
    static constexpr int SLICE_BYTES = A_BYTES + B_BYTES;
    // K slices per pipeline stage: with N <= 128 one slice is only <= 256 tensor-pipe cycles, on par with the
    // ~170-cycle mbarrier round trip per stage (tools/pipe_probe.cu), so two slices share one full/empty handshake
    static constexpr int KS = BN >= 256 ? 1 : 2;
    // ROW3 stage (3x3 stride-1 convs whose tile is bh >= 4 full-width rows of ONE sample, N <= 128): the three taps
    // (dy = -1, 0, +1) of one (source, 64-channel block, dx) read the SAME activation buffer of bh + 2 rows -- tap dy
    // is the UMMA descriptor advanced by (dy + 1) image rows -- so the L2 -> shared-memory traffic of the activation
    // operand drops from 3 x 16 KB to 24 KB per three taps.  N = 128 tiles are co-bound by exactly that traffic
    // (TMA-only 54 us, MMA-only 49 us, both 71 us on 128->128 3x3 @32^2).
    static constexpr int STAGE_BYTES = KS * SLICE_BYTES;
    static constexpr int UNITS = BN >= 128 ? BN / 128 : 1;  // 64-channel groups per epilogue warp
    static constexpr int EPI_BYTES = kEpiWarps * UNITS * EPI_BUF_BYTES;
    static constexpr int AUX_BYTES = 512;
    static constexpr int STAGES_FIT = (MAX_SMEM - 1024 - EPI_BYTES - AUX_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
    static constexpr int TMEM_COLS = 2 * BN;  // 128 / 256 / 512: powers of two >= 32
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + AUX_BYTES;
    static_assert(STAGES >= 3, "pipeline too shallow");
    static_assert((2 * STAGES + 4 + kEpiWarps) * 8 + 16 <= AUX_BYTES, "aux region too small");
};

struct TileCoord {
    int cls, n_tile, x0, y0, n0;
    int part;  // statistics slot of this tile inside its samples: cls * tiles_x * tiles_y + ty * tiles_x + tx
};
// pair-level tile index -> coordinates of THIS CTA's 128-row tile (rank selects the half of the 256-row tile)
template <int CG>
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile, int rank) {
    TileCoord t;
    t.n_tile = tile % p.n_tiles;
    int r = tile / p.n_tiles;
    const int m_group = r % p.m_groups;
    t.cls = r / p.m_groups;
    const int m_tile = m_group * CG + rank;  // may be == m_tiles (odd count): all rows out of range
    int tx = m_tile % p.tiles_x;
    int r2 = m_tile / p.tiles_x;
    int ty = r2 % p.tiles_y;
    int tn = r2 / p.tiles_y;
    t.x0 = tx * p.bw;
    t.y0 = ty * p.bh;
    t.n0 = tn * p.bn;
    t.part = (t.cls * p.tiles_y + ty) * p.tiles_x + tx;
    return t;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// named barrier over `count` threads (the four epilogue warps that hold one column half of a tile)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    // immediate barrier ids (a register id makes ptxas reserve all 16 hardware barriers)
    if (id == 1) asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"r"(count) : "memory");
}

template <int BN, int CG, bool OUT_F32>
__global__ void __launch_bounds__(kThreads, 1) igemm_sm100_kernel(const __grid_constant__ IgemmParams p) {
    using C = Cfg<BN, CG>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
    const uint32_t epi_base = base + C::STAGES * C::STAGE_BYTES;
    const uint32_t bar0 = epi_base + C::EPI_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + 2 + a); };
    auto res_bar = [&](int w) { return bar0 + 8u * (2 * C::STAGES + 4 + w); };
    const uint32_t slot_off = (bar0 - raw_addr) + 8u * (2 * C::STAGES + 4 + kEpiWarps);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + slot_off);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
    pdl_launch_dependents();  // the next kernel of the plan may take this SM as soon as this CTA leaves it
    const int cluster_id = blockIdx.x / CG;
    const int num_clusters = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.amap[i]);
        tma_prefetch_desc(&p.bmap);
        if constexpr (!OUT_F32) {
            tma_prefetch_desc(&p.omap[0]);
            if (p.has_res) tma_prefetch_desc(&p.rmap);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            // N = 64 tiles with bf16 output: the two column halves of the epilogue warps take ALTERNATE tiles (one
            // accumulator buffer each), so a buffer is handed back by the four warps of one half only
            mbar_init(tempty_bar(a), (BN == 64 && !OUT_F32) ? CG * (kEpiWarps / 2) : CG * kEpiWarps);
        }
        for (int w = 0; w < kEpiWarps; ++w) mbar_init(res_bar(w), 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_cg<CG>(smem_u32(const_cast<uint32_t*>(tmem_slot)), C::TMEM_COLS);
        tmem_relinquish_cg<CG>();
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers must exist before anything remote targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barriers, tensor memory, descriptor prefetch) overlapped the previous kernel's drain;
    // from here on this kernel reads what its predecessors wrote
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        // The whole warp walks the pipeline (every lane polls the mbarrier); one elected lane issues.  Under
        // elect.sync the operands of the TMA / tcgen05 instructions are provably warp-uniform, so ptxas keeps them
        // in uniform registers instead of wrapping each instruction in a per-lane R2UR "waterfall" loop -- the
        // single-lane version of this loop cost ~400 cycles per slice and bounded every tile shape with N < 256.
        const uint32_t full_leader0 = CG == 2 ? mapa_shared(full_bar(0), 0) : full_bar(0);
        int stage = 0;
        uint32_t phase = 0;
        const bool prof = p.prof != nullptr;
        long long w_empty = 0;
        for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
            const TileCoord t = decode_tile<CG>(p, tile, rank);
            const int4* sl = p.slices + (size_t)t.cls * p.num_stages * 3;
            const int b_row = CG == 2 ? t.n_tile * BN + rank * C::BNC : t.n_tile * BN;
            int4 r0 = __ldg(sl), r1 = __ldg(sl + 1), r2 = __ldg(sl + 2);
            for (int g = 0; g < p.num_stages; ++g) {
                // next stage record in flight while this stage waits for its buffer
                const int gn = g + 1 < p.num_stages ? g + 1 : g;
                const int4 n0 = __ldg(sl + 3 * gn), n1 = __ldg(sl + 3 * gn + 1), n2 = __ldg(sl + 3 * gn + 2);
                mbar_wait_prof(empty_bar(stage), phase ^ 1u, prof, w_empty);
                if (elect_one()) {
                    const uint32_t s_dst = base + stage * C::STAGE_BYTES;
                    const uint32_t fb = CG == 2 ? full_leader0 + 8u * stage : full_bar(stage);
                    if (p.probe & 1) {
                        if (rank == 0) mbar_arrive(full_bar(stage));
                    } else if (r2.x == 1) {
                        // ROW3: one activation buffer of bh + 2 rows, three weight blocks
                        const int src = (short)(r0.x & 0xffff);
                        const int dx = (short)(r0.x >> 16);
                        if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), CG * (p.a3_bytes + 3 * C::B_BYTES));
                        const uint32_t w_dst = s_dst + p.a3_bytes;
                        if constexpr (CG == 1) {
                            tma_load_4d(s_dst, &p.amap3[src], fb, r0.z, t.x0 + dx, t.y0 - 1, t.n0);
                            tma_load_2d(w_dst, &p.bmap, fb, r0.w * BK, b_row);
                            tma_load_2d(w_dst + C::B_BYTES, &p.bmap, fb, r1.x * BK, b_row);
                            tma_load_2d(w_dst + 2 * C::B_BYTES, &p.bmap, fb, r1.y * BK, b_row);
                        } else {
                            tma_load_4d_pair(s_dst, &p.amap3[src], fb, r0.z, t.x0 + dx, t.y0 - 1, t.n0);
                            tma_load_2d_pair(w_dst, &p.bmap, fb, r0.w * BK, b_row);
                            tma_load_2d_pair(w_dst + C::B_BYTES, &p.bmap, fb, r1.x * BK, b_row);
                            tma_load_2d_pair(w_dst + 2 * C::B_BYTES, &p.bmap, fb, r1.y * BK, b_row);
                        }
                    } else {
                        const int n_in = r2.y;
                        if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), n_in * CG * C::SLICE_BYTES);
#pragma unroll
                        for (int j = 0; j < C::KS; ++j) {
                            if (j < n_in) {
                                const int4 v = j == 0 ? r0 : r1;
                                const int src = (short)(v.x & 0xffff);
                                const int dx = (short)(v.x >> 16);
                                const int dy = (short)(v.y & 0xffff);
                                const uint32_t a_dst = s_dst + j * C::SLICE_BYTES;
                                if constexpr (CG == 1) {
                                    tma_load_4d(a_dst, &p.amap[src], fb, v.z, t.x0 + dx, t.y0 + dy, t.n0);
                                    tma_load_2d(a_dst + A_BYTES, &p.bmap, fb, v.w * BK, b_row);
                                } else {
                                    tma_load_4d_pair(a_dst, &p.amap[src], fb, v.z, t.x0 + dx, t.y0 + dy, t.n0);
                                    tma_load_2d_pair(a_dst + A_BYTES, &p.bmap, fb, v.w * BK, b_row);
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                r0 = n0; r1 = n1; r2 = n2;
                if (++stage == C::STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        if (prof && lane == 0) p.prof[blockIdx.x * 16 + 3] = (unsigned long long)w_empty;
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA, whole warp)
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM * CG, BN);
            // descriptor = constant high word | (smem address >> 4): per stage only the low word moves
            const uint32_t desc_lo0 = ((base & 0x3FFFFu) >> 4) | (1u << 16);
            constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const bool prof = p.prof != nullptr;
            const bool no_mma = (p.probe & 2) != 0;
            long long w_full = 0, w_tempty = 0;
            const long long t_begin = clk();
            for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
                mbar_wait_prof(tempty_bar(acc), acc_phase ^ 1u, prof, w_tempty);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                const int4* sl = p.slices + (size_t)(tile / (p.n_tiles * p.m_groups)) * p.num_stages * 3 + 2;
                int4 r2 = __ldg(sl);
                for (int g = 0; g < p.num_stages; ++g) {
                    const int4 n2 = __ldg(sl + 3 * (g + 1 < p.num_stages ? g + 1 : g));
                    mbar_wait_prof(full_bar(stage), phase, prof, w_full);
                    tc_fence_after();
                    if (elect_one()) {
                        if (!no_mma) {
                            const uint32_t s_lo = desc_lo0 + ((stage * C::STAGE_BYTES) >> 4);
                            if (r2.x == 1) {
                                const uint32_t w_lo = s_lo + (p.a3_bytes >> 4);
#pragma unroll
                                for (int j = 0; j < 3; ++j) {
                                    const uint32_t a_lo = s_lo + j * (p.row_bytes >> 4);
                                    const uint32_t b_lo = w_lo + j * (C::B_BYTES >> 4);
#pragma unroll
                                    for (int k = 0; k < BK / 16; ++k)
                                        umma_bf16_cg<CG>(d_tmem, umma_desc_pack(a_lo + 2u * k, desc_hi),
                                                         umma_desc_pack(b_lo + 2u * k, desc_hi), idesc, (g | j | k) != 0);
                                }
                            } else {
                                const int n_in = r2.y;
#pragma unroll
                                for (int j = 0; j < C::KS; ++j) {
                                    if (j < n_in) {
                                        const uint32_t a_lo = s_lo + ((j * C::SLICE_BYTES) >> 4);
                                        const uint32_t b_lo = a_lo + (A_BYTES >> 4);
#pragma unroll
                                        for (int k = 0; k < BK / 16; ++k) {
                                            // +32 B per K=16 step inside the 128 B swizzle row (address field is >> 4)
                                            umma_bf16_cg<CG>(d_tmem, umma_desc_pack(a_lo + 2u * k, desc_hi),
                                                             umma_desc_pack(b_lo + 2u * k, desc_hi), idesc, (g | j | k) != 0);
                                        }
                                    }
                                }
                            }
                        }
                        umma_commit_cg<CG>(empty_bar(stage));
                    }
                    __syncwarp();
                    r2 = n2;
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (elect_one()) umma_commit_cg<CG>(tfull_bar(acc));
                __syncwarp();
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
            if (prof && lane == 0) {
                p.prof[blockIdx.x * 16 + 0] = (unsigned long long)(clk() - t_begin);
                p.prof[blockIdx.x * 16 + 1] = (unsigned long long)w_full;
                p.prof[blockIdx.x * 16 + 2] = (unsigned long long)w_tempty;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (8 warps, both CTAs)
        const int ew = warp - 4;
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        const int half = ew >> 2;
        const int row = q * 32 + lane;
        const int rows_per_sample = p.bw * p.bh;
        const int dxr = row % p.bw;
        const int dyr = (row / p.bw) % p.bh;
        const int dnr = row / rows_per_sample;
        // origin of this warp's 32-row sub-box inside the tile box
        const int sx0 = (q * 32) % p.bw, sy0 = ((q * 32) / p.bw) % p.bh, sn0 = (q * 32) / rows_per_sample;
        const uint32_t ebuf = epi_base + ew * C::UNITS * EPI_BUF_BYTES;
        const uint32_t rbar = res_bar(ew);
        uint32_t rphase = 0;
        const uint32_t tempty0 = CG == 2 ? mapa_shared(tempty_bar(0), 0) : tempty_bar(0);
        // N = 64, bf16 output: a tile has ONE 64-channel unit, which would leave the four warps of column half 1 idle while
        // the epilogue -- not the 5-9 slice mainloop -- bounds the kernel (64->64 k5 @4064: 4 560 cycles per tile against
        // 640 tensor-pipe cycles).  Instead half h takes every second tile of the CTA (always accumulator buffer h): two
        // tiles drain concurrently, each through its own staging buffers and its own named barrier
        // (64->64 k5 @4064 x 64: 44.4 -> 29.7 us; 64->64 3x3 @128^2 x 64: 166 -> 139 us).
        constexpr bool ALT = BN == 64 && !OUT_F32;
        int acc = ALT ? half : 0;
        uint32_t acc_phase = 0;
        int iter = 0;
        const bool prof = p.prof != nullptr && ew == 0;
        long long w_tfull = 0, w_res = 0, w_store = 0;
        uint32_t et[6] = {0, 0, 0, 0, 0, 0}, tlast = 0;  // prof: pre / waits / ld+math+sts / hand-back / store / statistics
#define TQ_EPI_T(i) do { if (prof) { const uint32_t n_ = (uint32_t)clock(); et[i] += n_ - tlast; tlast = n_; } } while (0)
        const long long e_begin = clk();
        for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters, ++iter) {
            if constexpr (ALT) {
                if ((iter & 1) != half) continue;
            }
            if (prof) tlast = (uint32_t)clock();
            const TileCoord t = decode_tile<CG>(p, tile, rank);
            const int n = t.n0 + dnr, y = t.y0 + dyr, x = t.x0 + dxr;
            const bool valid = (n < p.N) && (y < p.H) && (x < p.W);
            const float* emb_row = p.emb ? p.emb + (long long)(n < p.N ? n : 0) * p.emb_ld : nullptr;
            const uint32_t acc_addr = tmem_base + acc * BN + (uint32_t(q * 32) << 16);

            if (p.probe & 4) {
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2) mbar_arrive_cluster(tempty0 + 8u * acc);
                    else mbar_arrive(tempty_bar(acc));
                }
                if constexpr (ALT) acc_phase ^= 1u;
                else {
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1u;
                }
                continue;
            }
            if constexpr (!OUT_F32) {
                // 64-channel units of this warp: column group half + 2u of the BN tile (BN = 64: half 0 only)
                bool unit_on[C::UNITS];
                int unit_col[C::UNITS];
                int n_on = 0;
#pragma unroll
                for (int u = 0; u < C::UNITS; ++u) {
                    unit_col[u] = (BN >= 128 ? half + 2 * u : 0) * 64;
                    unit_on[u] = (t.n_tile * BN + unit_col[u]) < p.cout;   // N = 64: both halves work (alternate tiles)
                    n_on += unit_on[u] ? 1 : 0;
                }
                // statistics over tiles whose samples span several epilogue warps are read ACROSS the staging buffers
                // of the four warps of a column half: nobody may refill a buffer before all four have finished reading
                const bool xstats = p.stats != nullptr && p.wps > 1 && n_on > 0;
                if (xstats) named_bar_sync(1 + half, 128);
                // staging buffers are free once the previous tile's TMA stores have finished reading them
                {
                    const long long t0 = prof ? clk() : 0;
                    if (lane == 0) bulk_wait_read<0>();
                    if (prof) w_store += clk() - t0;
                }
                fence_proxy_async();  // the statistics pass read the buffers through the generic proxy
                __syncwarp();
                const bool res = p.has_res && n_on > 0;
                if (res && lane == 0) {
                    mbar_arrive_expect_tx(rbar, n_on * EPI_BUF_BYTES);
#pragma unroll
                    for (int u = 0; u < C::UNITS; ++u)
                        if (unit_on[u])
                            tma_load_4d(ebuf + u * EPI_BUF_BYTES, &p.rmap, rbar, t.n_tile * BN + unit_col[u], t.x0 + sx0,
                                        t.y0 + sy0, t.n0 + sn0);
                }
                if (emb_row) {
                    // the per-sample embedding row is an L2 round trip: start it before the accumulator wait
#pragma unroll
                    for (int u = 0; u < C::UNITS; ++u)
                        if (unit_on[u]) {
                            prefetch_l1(emb_row + t.n_tile * BN + unit_col[u]);
                            prefetch_l1(emb_row + t.n_tile * BN + unit_col[u] + 32);
                        }
                }
                TQ_EPI_T(0);
                mbar_wait_prof(tfull_bar(acc), acc_phase, prof, w_tfull);
                tc_fence_after();
                if (res) {
                    mbar_wait_prof(rbar, rphase, prof, w_res);
                    rphase ^= 1u;
                }
                TQ_EPI_T(1);
                int done = 0;
#pragma unroll
                for (int u = 0; u < C::UNITS; ++u) {
                    if (!unit_on[u]) continue;  // warp-uniform
                    const int cg0 = t.n_tile * BN + unit_col[u];
                    const uint32_t buf = ebuf + u * EPI_BUF_BYTES;
                    const uint32_t rowp = buf + lane * 128;
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        uint32_t r[32];
                        tmem_ld_32x32(acc_addr + unit_col[u] + h2 * 32, r);
                        float add[32];
                        const int cg = cg0 + h2 * 32;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + cg + j))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                            if (emb_row) {
                                const float4 e = __ldg(reinterpret_cast<const float4*>(emb_row + cg + j));
                                b.x += e.x; b.y += e.y; b.z += e.z; b.w += e.w;
                            }
                            add[j] = b.x; add[j + 1] = b.y; add[j + 2] = b.z; add[j + 3] = b.w;
                        }
                        // the four residual chunks of this row half: all loads issued back to back, ahead of the wait
                        uint4 rq[4];
                        if (res) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) rq[j] = lds128(rowp + (((h2 * 4 + j) ^ (lane & 7)) << 4));
                        }
                        tmem_ld_wait();
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + add[j];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            // 16 B chunk (h2*4 + j) of this row sits at chunk position (.. ^ (row & 7)): SWIZZLE_128B
                            const uint32_t a16 = rowp + (((h2 * 4 + j) ^ (lane & 7)) << 4);
                            if (res) {
                                const uint4 u4 = rq[j];
                                const uint32_t rw[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[k]);
                                    v[j * 8 + 2 * k] += __low2float(b2);
                                    v[j * 8 + 2 * k + 1] += __high2float(b2);
                                }
                            }
                            uint32_t w[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[j * 8 + 2 * k], v[j * 8 + 2 * k + 1]);
                                // rows beyond the tensor edge are clipped by the TMA store; zero them for the statistics
                                w[k] = valid ? *reinterpret_cast<const uint32_t*>(&b2) : 0u;
                            }
                            sts128(a16, make_uint4(w[0], w[1], w[2], w[3]));
                        }
                    }
                    TQ_EPI_T(2);
                    if (++done == n_on) {
                        // accumulator fully read by this warp: hand the TMEM buffer back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (CG == 2) mbar_arrive_cluster(tempty0 + 8u * acc);
                            else mbar_arrive(tempty_bar(acc));
                        }
                    }
                    TQ_EPI_T(3);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(&p.omap[t.cls], buf, cg0, t.x0 + sx0, t.y0 + sy0, t.n0 + sn0);
                        bulk_commit();
                    }
                    TQ_EPI_T(4);
                    if (p.stats != nullptr && p.wps == 1) {
                        // GroupNorm statistics of the consumer, samples of <= 32 rows: a segment of this warp's rows is a
                        // whole (tile of a) sample.  Lane l owns channels cg0 + 2l, 2l+1 (one 32-bit word per row of the
                        // staging buffer); the slot [n][part][c] has exactly one writer: plain stores, no atomics.
                        const uint32_t wsel = lane >> 2, wlo = (lane & 3) << 2;
                        const int ns0 = t.n0 + (q * 32) / rows_per_sample;
                        const int nseg = 32 / p.seg;
                        for (int sg = 0; sg < nseg; ++sg) {
                            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                            const int r0 = sg * p.seg;
                            // seg is a power of two: groups of min(seg, 8) rows, the loads of a group issued together
                            // (one destination register per load, or ptxas serialises them on LDS latency)
                            const int grp = p.seg < 8 ? p.seg : 8;
                            for (int r2 = 0; r2 < p.seg; r2 += grp) {
                                uint32_t wv[8];
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const int rr = r0 + r2 + (k < grp ? k : 0);
                                    wv[k] = lds32(buf + rr * 128 + (((wsel ^ (rr & 7)) << 4) | wlo));
                                }
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    if (k < grp) {
                                        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&wv[k]);
                                        const float f0 = __low2float(b2), f1 = __high2float(b2);
                                        s0 += f0; q0 = fmaf(f0, f0, q0);
                                        s1 += f1; q1 = fmaf(f1, f1, q1);
                                    }
                                }
                            }
                            const int ns = ns0 + (nseg > 1 ? sg : 0);
                            if (ns < p.N)
                                *reinterpret_cast<float4*>(p.stats + (((long long)ns * p.parts + t.part) * p.cout + cg0 + 2 * lane) * 2) =
                                    make_float4(s0, q0, s1, q1);
                        }
                    }
                    TQ_EPI_T(5);
                }
                if (xstats) {
                    // samples of 64 / 128 rows inside the tile: the wps = 2 / 4 warps that hold one sample split its 64
                    // channels between them and each sums ITS channels over all wps x 32 rows, reading the staging buffers
                    // of the whole group (a fixed order: deterministic, and no exchange buffer is needed).  Lane =
                    // (row group rg = source warp, channel pair cp); the row rotation per rg keeps the four row groups on
                    // different shared-memory banks (SWIZZLE_128B: the bank depends on chunk ^ (row & 7)).
                    named_bar_sync(1 + half, 128);  // all four staging buffers of this column half are written
                    const int wps = p.wps, cpn = 32 / wps;
                    const int cp = lane % cpn, rg = lane / cpn;
                    const int qs = q % wps, qg = q - qs;
                    const int idx = qs * cpn + cp;  // channel pair inside the 64-channel unit
                    const uint32_t chunk = idx >> 2, wlo = (idx & 3) << 2;
                    const uint32_t sbuf = epi_base + ((half * 4 + qg + rg) * C::UNITS) * EPI_BUF_BYTES;
                    const int rot = (8 / wps) * rg;
                    const int ns = t.n0 + qg / wps;
#pragma unroll
                    for (int u = 0; u < C::UNITS; ++u) {
                        if (!unit_on[u]) continue;  // uniform over the four warps of the half
                        const uint32_t buf = sbuf + u * EPI_BUF_BYTES;
                        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
                        for (int r2 = 0; r2 < 32; r2 += 8) {
                            uint32_t wv[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const int rr = (r2 + k + rot) & 31;
                                wv[k] = lds32(buf + rr * 128 + (((chunk ^ (rr & 7)) << 4) | wlo));
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&wv[k]);
                                const float f0 = __low2float(b2), f1 = __high2float(b2);
                                s0 += f0; q0 = fmaf(f0, f0, q0);
                                s1 += f1; q1 = fmaf(f1, f1, q1);
                            }
                        }
                        for (int o = cpn; o < 32; o <<= 1) {
                            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                            q0 += __shfl_xor_sync(0xffffffffu, q0, o);
                            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                            q1 += __shfl_xor_sync(0xffffffffu, q1, o);
                        }
                        const int cg0 = t.n_tile * BN + unit_col[u];
                        if (rg == 0 && ns < p.N)
                            *reinterpret_cast<float4*>(p.stats + (((long long)ns * p.parts + t.part) * p.cout + cg0 + 2 * idx) * 2) =
                                make_float4(s0, q0, s1, q1);
                    }
                    TQ_EPI_T(5);
                }
                if (n_on == 0) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CG == 2) mbar_arrive_cluster(tempty0 + 8u * acc);
                        else mbar_arrive(tempty_bar(acc));
                    }
                }
            } else {
                // ---------------- fp32 output: direct stores (embedding GEMM, ragged Cout convs)
                const long long off = p.out_class_off[t.cls] + (long long)n * p.out_sn + (long long)y * p.out_sy +
                                      (long long)x * p.out_sx;
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                constexpr int CH = BN / 2 / 32;  // 32-column chunks per warp
#pragma unroll 1
                for (int ci = 0; ci < CH; ++ci) {
                    const int c = half * (BN / 2) + ci * 32;
                    uint32_t r[32];
                    tmem_ld_32x32(acc_addr + c, r);
                    tmem_ld_wait();
                    const int cg = t.n_tile * BN + c;
                    if (!valid || cg >= p.cout) continue;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + cg + j) : 0.f);
                    if (p.vec_ok) {
                        if (emb_row) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 e = __ldg(reinterpret_cast<const float4*>(emb_row + cg + j));
                                v[j] += e.x; v[j + 1] += e.y; v[j + 2] += e.z; v[j + 3] += e.w;
                            }
                        }
                        if (p.residual) {
                            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off + cg);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint4 u = __ldg(rp + j);
                                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
                                    v[j * 8 + 2 * k] += __low2float(b2);
                                    v[j * 8 + 2 * k + 1] += __high2float(b2);
                                }
                            }
                        }
                        float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + off + cg);
#pragma unroll
                        for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
                        // ragged / unaligned edge (Cout in {3, 6, 8}): scalar path
#pragma unroll 1
                        for (int j = 0; j < 32; ++j) {
                            if (cg + j >= p.cout) break;
                            float o = v[j];
                            if (emb_row) o += __ldg(emb_row + cg + j);
                            if (p.residual) o += __bfloat162float(p.residual[off + cg + j]);
                            static_cast<float*>(p.out)[off + cg + j] = o;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2) mbar_arrive_cluster(tempty0 + 8u * acc);
                    else mbar_arrive(tempty_bar(acc));
                }
            }
            if constexpr (ALT) acc_phase ^= 1u;   // this half owns accumulator buffer `half`: every tile it takes flips the phase
            else {
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
        if constexpr (!OUT_F32) {
            if (lane == 0) bulk_wait_read<0>();  // smem must stay alive until the last TMA stores have read it
        }
        if (prof && lane == 0) {
            p.prof[blockIdx.x * 16 + 4] = (unsigned long long)(clk() - e_begin);
            p.prof[blockIdx.x * 16 + 5] = (unsigned long long)w_tfull;
            p.prof[blockIdx.x * 16 + 6] = (unsigned long long)w_res;
            p.prof[blockIdx.x * 16 + 7] = (unsigned long long)w_store;
            for (int i = 0; i < 6; ++i) p.prof[blockIdx.x * 16 + 8 + i] = (unsigned long long)et[i];
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();  // no CTA of the pair may exit while the other can still signal it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg<CG>(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

// channels-last [N, H, W, C] view with element strides (sn, sy, sx, 1) -> 4-D map with box {64, bw, bh, bn}
int encode_nhwc_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int Cc, long long sn, long long sy, long long sx,
                    int bw, int bh, int bn, const char* what) {
    auto enc = get_encode_fn();
    TQ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    TQ_CHECK(Cc % 64 == 0, "%s: channels must be a multiple of 64 (got %d)", what, Cc);
    TQ_CHECK(sx % 8 == 0 && sn % 8 == 0 && (H == 1 || sy % 8 == 0), "%s: strides must be 16 B multiples", what);
    TQ_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s: pointer must be 16 B aligned", what);
    cuuint64_t dims[4] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const long long sy_eff = (H == 1 && (sy == 0 || sy % 8 != 0)) ? (long long)W * sx : sy;
    cuuint64_t strides[3] = {(cuuint64_t)sx * 2, (cuuint64_t)sy_eff * 2, (cuuint64_t)sn * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TQ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return 0;
}

int encode_weight_map(CUtensorMap* m, const void* w, int cout_pad, int ktot, int rows) {
    auto enc = get_encode_fn();
    TQ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TQ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
    return 0;
}

int pow2_floor(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}

template <int BN, int CG, bool OUT_F32>
int launch_igemm(const IgemmParams& p, int grid, cudaStream_t st) {
    using C = Cfg<BN, CG>;
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        TQ_CUDA(cudaFuncSetAttribute(igemm_sm100_kernel<BN, CG, OUT_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    TQ_CUDA(cudaLaunchKernelEx(&cfg, igemm_sm100_kernel<BN, CG, OUT_F32>, p));
    count_launch();
    return 0;
}

template <int BN, int CG>
int launch_igemm_o(const IgemmParams& p, int grid, bool f32, cudaStream_t st) {
    return f32 ? launch_igemm<BN, CG, true>(p, grid, st) : launch_igemm<BN, CG, false>(p, grid, st);
}

// The automatic (BN, CG = 2) choice: cost = c0 + rounds * (slices * a + e) in SM cycles, fitted to the forced-tile sweep
// of every convolution of a latent denoiser call (tools/tile_sweep.py, profiles/r2_tile_sweep_latent.txt):
//   a   cycles per 64-deep K slice: the tensor pipe at N = 256 / 128 (2 * BN: per FLOP the two are equally efficient since
//       two slices share a stage at N <= 128), the mbarrier handshake floor at N = 64
//   e   per-round cost that does not overlap the next tile's mainloop (accumulator hand-over, epilogue tail)
//   c0  per launch: pipeline fill and the last epilogue, which nothing overlaps
// The model it replaces charged N = 128 slices for their shared-memory traffic (384 cycles) and 700 cycles per round: it kept
// the 16 x 16 level of the latent UNet on 256-wide tiles (256 tiles on 74 clusters = 4 rounds at 86 %) where 128-wide tiles
// (512 tiles, 7 rounds at 99 %) measure 17 % faster.
struct TileCost { long long a, e, c0; };
TileCost tile_cost(int bn) {
    if (bn >= 256) return {516, 2900, 12800};
    if (bn >= 128) return {265, 2450, 10150};
    return {200, 1500, 8750};
}
long long conv_cost(int bn, long long rounds, long long slices) {
    const TileCost c = tile_cost(bn);
    return c.c0 + rounds * (slices * c.a + c.e);
}

}  // namespace

void tile_shape_for(int H, int W, int* bw, int* bh, int* bn) {
    int w = W >= 128 ? 128 : pow2_floor(W);
    int h = 128 / w;
    if (h > 1) h = pow2_floor(H) < h ? pow2_floor(H) : h;
    *bw = w;
    *bh = h;
    *bn = 128 / (w * h);
}

int conv_stats_parts_sm100(const tq_conv_desc& d) {
    int bw, bh, bn;
    tile_shape_for(d.H, d.W, &bw, &bh, &bn);
    // one slot per 128-row tile: a slot per epilogue WARP (no cross-warp pass, no named barriers, 4x the partials for the
    // consuming GroupNorm to add) measured slower end to end -- latent UNet call 4.405 -> 4.481 ms, 1D EDM step 124 -> 141 ms
    return d.num_classes * ((d.W + bw - 1) / bw) * ((d.H + bh - 1) / bh);
}

int build_conv_sm100(std::vector<Op>& ops, const tq_conv_desc& d) {
    TQ_CHECK(d.dtype == TQ_BF16, "sm100 igemm needs bf16 operands");
    TQ_CHECK(d.num_srcs >= 1 && d.num_srcs <= 4, "num_srcs out of range");
    TQ_CHECK(d.num_classes == 1 || d.num_classes == 2 || d.num_classes == 4, "num_classes must be 1, 2 or 4");
    TQ_CHECK(d.num_slices >= 1, "conv needs at least one K slice");
    TQ_CHECK(d.ktot % 64 == 0 && d.cout_pad % 64 == 0, "weight matrix must be padded to 64x64 blocks");
    TQ_CHECK(d.cout >= 1 && d.cout <= d.cout_pad, "cout out of range");
    TQ_CHECK((reinterpret_cast<uintptr_t>(d.weights) & 15) == 0, "weights must be 16 B aligned");
    TQ_CHECK(d.out_dtype == TQ_F32 || d.out_dtype == TQ_BF16, "bad out_dtype");
    const bool f32 = d.out_dtype == TQ_F32;
    TQ_CHECK(f32 || d.cout % 64 == 0, "bf16 outputs need cout %% 64 == 0 on the tensor path (got %d)", d.cout);
    TQ_CHECK(d.stats == nullptr || !f32, "conv statistics are produced for bf16 outputs only");
    TQ_CHECK(d.residual == nullptr || d.num_classes == 1, "residual add needs a single output class");

    auto p = std::make_shared<IgemmParams>();
    memset(p.get(), 0, sizeof(IgemmParams));
    p->N = d.N; p->H = d.H; p->W = d.W;
    tile_shape_for(d.H, d.W, &p->bw, &p->bh, &p->bn);
    p->tiles_x = (d.W + p->bw - 1) / p->bw;
    p->tiles_y = (d.H + p->bh - 1) / p->bh;
    const int tiles_n = (d.N + p->bn - 1) / p->bn;
    p->m_tiles = p->tiles_x * p->tiles_y * tiles_n;
    const int sms = device_sm_count();

    int cg = d.cta_group;
    TQ_CHECK(cg == 0 || cg == 1 || cg == 2, "cta_group must be 0 (auto), 1 or 2");
    if (cg == 0) cg = p->m_tiles >= 2 ? 2 : 1;
    int bn_tile = d.block_n;
    TQ_CHECK(bn_tile == 0 || bn_tile == 64 || bn_tile == 128 || bn_tile == 256, "block_n must be 0, 64, 128 or 256");
    if (bn_tile == 0) {
        long long best = -1;
        for (int cand : {256, 128, 64}) {
            if (cand > 64 && cand / 2 >= d.cout_pad) continue;  // would mostly compute padding
            const long long m_groups = (p->m_tiles + cg - 1) / cg;
            const long long tiles = m_groups * ((d.cout_pad + cand - 1) / cand) * d.num_classes;
            const long long clusters = sms / cg;
            const long long rounds = (tiles + clusters - 1) / clusters;
            const long long cost = conv_cost(cand, rounds, d.num_slices);
            if (best < 0 || cost < best) {
                best = cost;
                bn_tile = cand;
            }
        }
    }
    p->m_groups = (p->m_tiles + cg - 1) / cg;
    p->n_tiles = (d.cout_pad + bn_tile - 1) / bn_tile;
    p->total_tiles = p->m_groups * p->n_tiles * d.num_classes;

    for (int i = 0; i < d.num_srcs; ++i) {
        const tq_src& s = d.srcs[i];
        if (encode_nhwc_map(&p->amap[i], s.ptr, s.N, s.H, s.W, s.C, s.sn, s.sy, s.sx, p->bw, p->bh, p->bn, "conv source"))
            return 1;
    }
    for (int i = d.num_srcs; i < 4; ++i) p->amap[i] = p->amap[0];
    if (encode_weight_map(&p->bmap, d.weights, d.cout_pad, d.ktot, bn_tile / cg)) return 1;

    // epilogue unit box: the 32 consecutive tile rows one epilogue warp owns
    const int sbw = p->bw < 32 ? p->bw : 32;
    const int sbh = p->bh < 32 / sbw ? p->bh : 32 / sbw;
    const int sbn = 32 / (sbw * sbh);
    if (!f32) {
        for (int c = 0; c < d.num_classes; ++c) {
            const __nv_bfloat16* ob = static_cast<const __nv_bfloat16*>(d.out) + d.out_class_off[c];
            if (encode_nhwc_map(&p->omap[c], ob, d.N, d.H, d.W, d.cout, d.out_sn, d.out_sy, d.out_sx, sbw, sbh, sbn,
                                "conv output"))
                return 1;
        }
        for (int c = d.num_classes; c < 4; ++c) p->omap[c] = p->omap[0];
        if (d.residual) {
            if (encode_nhwc_map(&p->rmap, d.residual, d.N, d.H, d.W, d.cout, d.out_sn, d.out_sy, d.out_sx, sbw, sbh, sbn,
                                "conv residual"))
                return 1;
            p->has_res = 1;
        } else {
            p->rmap = p->omap[0];
        }
    }

    const size_t nsl = (size_t)d.num_classes * d.num_slices;
    for (size_t i = 0; i < nsl; ++i) {
        const tq_slice& s = d.slices[i];
        TQ_CHECK(s.src >= 0 && s.src < d.num_srcs, "slice %zu: bad source index", i);
        TQ_CHECK(s.c0 % 64 == 0 && s.c0 + 64 <= d.srcs[s.src].C, "slice %zu: bad channel offset", i);
        TQ_CHECK(s.kb >= 0 && (s.kb + 1) * 64 <= d.ktot, "slice %zu: bad weight block", i);
    }
    // ---- stage records (3 x int4 per stage): plain stages carry up to KS slices, a ROW3 stage the three dy taps of
    // one (source, channel block, dx) over a shared activation buffer of bh + 2 rows
    const int ks = bn_tile >= 256 ? 1 : 2;
    const int b_bytes = (bn_tile / cg) * BK * 2;
    const int stage_bytes = ks * (A_BYTES + b_bytes);
    const int a3_bytes = (p->bh + 2) * p->bw * BK * 2;
    bool row3 = bn_tile <= 128 && p->bn == 1 && p->bw == d.W && p->bh >= 2 && d.num_classes == 1 &&
                a3_bytes + 3 * b_bytes <= stage_bytes && (p->bw * BK * 2) % 1024 == 0;
    if (const char* e = getenv("TQ_ROW3"); e && e[0] == '0') row3 = false;
    std::vector<int4> recs;
    int num_stages = 0, row3_groups = 0;
    auto pack = [](const tq_slice& sl) {
        return make_int4((int)((uint32_t)(uint16_t)sl.src | ((uint32_t)(uint16_t)sl.dx << 16)), (int)(uint16_t)sl.dy, sl.c0, sl.kb);
    };
    for (int c = 0; c < d.num_classes; ++c) {
        const tq_slice* sl = d.slices + (size_t)c * d.num_slices;
        std::vector<int> group(d.num_slices, -1);  // slice -> ROW3 group leader (the dy = -1 slice), or -1
        if (row3) {
            for (int i = 0; i < d.num_slices; ++i) {
                if (sl[i].dy != -1 || group[i] != -1) continue;
                int mid = -1, bot = -1;
                for (int j = 0; j < d.num_slices; ++j) {
                    if (group[j] != -1 || sl[j].src != sl[i].src || sl[j].c0 != sl[i].c0 || sl[j].dx != sl[i].dx) continue;
                    if (sl[j].dy == 0 && mid < 0) mid = j;
                    if (sl[j].dy == 1 && bot < 0) bot = j;
                }
                if (mid >= 0 && bot >= 0) {
                    group[i] = i; group[mid] = i; group[bot] = i;
                    recs.push_back(pack(sl[i]));
                    recs.push_back(make_int4(sl[mid].kb, sl[bot].kb, 0, 0));
                    recs.push_back(make_int4(1, 3, 0, 0));
                    ++row3_groups;
                }
            }
        }
        std::vector<int> plain;
        for (int i = 0; i < d.num_slices; ++i)
            if (group[i] == -1) plain.push_back(i);
        for (size_t i = 0; i < plain.size(); i += ks) {
            const int n_in = (int)std::min<size_t>(ks, plain.size() - i);
            recs.push_back(pack(sl[plain[i]]));
            recs.push_back(n_in > 1 ? pack(sl[plain[i + 1]]) : make_int4(0, 0, 0, -1));
            recs.push_back(make_int4(0, n_in, 0, 0));
        }
        const int st = (int)recs.size() / 3 - num_stages * c;
        if (c == 0) num_stages = st;
        TQ_CHECK(st == num_stages, "conv classes must have the same stage count");
    }
    if (row3_groups > 0) {
        for (int i = 0; i < d.num_srcs; ++i) {
            const tq_src& sr = d.srcs[i];
            if (encode_nhwc_map(&p->amap3[i], sr.ptr, sr.N, sr.H, sr.W, sr.C, sr.sn, sr.sy, sr.sx, p->bw, p->bh + 2, p->bn,
                                "conv source (row3)"))
                return 1;
        }
        for (int i = d.num_srcs; i < 4; ++i) p->amap3[i] = p->amap3[0];
    } else {
        for (int i = 0; i < 4; ++i) p->amap3[i] = p->amap[0];
    }
    p->num_stages = num_stages;
    p->a3_bytes = a3_bytes;
    p->row_bytes = p->bw * BK * 2;
    void* dsl = nullptr;
    TQ_CUDA(cudaMalloc(&dsl, recs.size() * sizeof(int4)));
    std::shared_ptr<void> dsl_owner(dsl, [](void* q) { cudaFree(q); });
    TQ_CUDA(cudaMemcpy(dsl, recs.data(), recs.size() * sizeof(int4), cudaMemcpyHostToDevice));
    static_assert(sizeof(tq_slice) == sizeof(int4), "tq_slice must be 16 bytes");

    p->slices = static_cast<const int4*>(dsl);
    p->num_slices = d.num_slices;
    p->num_classes = d.num_classes;
    p->cout = d.cout;
    p->bias = d.bias;
    p->emb = d.emb;
    p->emb_ld = d.emb_ld;
    p->residual = static_cast<const __nv_bfloat16*>(d.residual);
    p->out = d.out;
    p->out_sn = d.out_sn; p->out_sy = d.out_sy; p->out_sx = d.out_sx;
    for (int i = 0; i < 4; ++i) p->out_class_off[i] = d.out_class_off[i];
    if (f32) {
        bool vec = d.cout % 32 == 0 && d.out_sn % 4 == 0 && d.out_sy % 4 == 0 && d.out_sx % 4 == 0;
        vec = vec && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0;
        for (int i = 0; i < d.num_classes; ++i) vec = vec && d.out_class_off[i] % 4 == 0;
        if (d.emb) vec = vec && d.emb_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(d.emb) & 15) == 0;
        if (d.residual) vec = vec && (reinterpret_cast<uintptr_t>(d.residual) & 15) == 0 && d.out_sn % 8 == 0 &&
                              d.out_sy % 8 == 0 && d.out_sx % 8 == 0;
        p->vec_ok = vec ? 1 : 0;
    } else {
        TQ_CHECK(!d.bias || (reinterpret_cast<uintptr_t>(d.bias) & 15) == 0, "bias must be 16 B aligned");
        TQ_CHECK(!d.emb || (d.emb_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(d.emb) & 15) == 0),
                 "per-sample embedding must be 16 B aligned with emb_ld %% 4 == 0");
    }
    p->stats = d.stats;
    {
        const int rows = p->bw * p->bh;
        p->seg = rows < 32 ? rows : 32;
        p->wps = rows <= 32 ? 1 : rows / 32;
        p->parts = d.num_classes * p->tiles_x * p->tiles_y;
    }
    TQ_CHECK(d.stats == nullptr || (reinterpret_cast<uintptr_t>(d.stats) & 15) == 0,
             "conv statistics buffer must be 16 B aligned");
    TQ_CHECK(d.stats == nullptr || d.stats_parts == p->parts,
             "conv statistics: stats_parts = %d, this geometry writes %d parts per sample (tq_conv_stats_parts)", d.stats_parts,
             p->parts);

    const int clusters = p->total_tiles < sms / cg ? p->total_tiles : sms / cg;
    const int grid = clusters * cg;
    std::shared_ptr<void> prof_owner;
    if (const char* e = getenv("TQ_IGEMM_PROBE")) p->probe = atoi(e);
    if (const char* e = getenv("TQ_IGEMM_PROF"); e && e[0] == '1') {
        void* pb = nullptr;
        TQ_CUDA(cudaMalloc(&pb, (size_t)grid * 16 * sizeof(unsigned long long)));
        TQ_CUDA(cudaMemset(pb, 0, (size_t)grid * 16 * sizeof(unsigned long long)));
        prof_owner.reset(pb, [](void* q) { cudaFree(q); });
        p->prof = static_cast<unsigned long long*>(pb);
    }

    Op op;
    char nm[112];
    snprintf(nm, sizeof nm, "igemm_sm100<BN=%d,CG=%d,%s> tiles=%d slices=%d%s", bn_tile, cg, f32 ? "f32" : "bf16",
             p->total_tiles, d.num_slices, row3_groups > 0 ? " row3" : "");
    // tile FLOPs (padding included) under ~40 GFLOP: a launch of <= ~35 us that is bound by its fill / drain latency
    op.small = 2.0 * p->total_tiles * (128.0 * cg) * bn_tile * d.num_slices * 64.0 < 40e9;
    op.name = nm;
    const std::string opname = nm;
    op.launch = [p, dsl_owner, prof_owner, opname, grid, bn_tile, cg, f32](cudaStream_t st) -> int {
        struct ProfDump {
            const IgemmParams* p; int grid; int cg; cudaStream_t st; const std::string* name;
            ~ProfDump() {
                if (!p->prof) return;
                cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
                cudaStreamIsCapturing(st, &cs);
                if (cs != cudaStreamCaptureStatusNone) return;
                cudaStreamSynchronize(st);
                std::vector<unsigned long long> h((size_t)grid * 16);
                cudaMemcpy(h.data(), p->prof, h.size() * 8, cudaMemcpyDeviceToHost);
                double a[16] = {};
                int nl = 0;
                for (int b = 0; b < grid; ++b) {
                    for (int k = 3; k < 16; ++k) a[k] += (double)h[b * 16 + k] / grid;
                    if (b % cg == 0) {
                        for (int k = 0; k < 3; ++k) a[k] += (double)h[b * 16 + k];
                        ++nl;
                    }
                }
                for (int k = 0; k < 3; ++k) a[k] /= nl;
                const double tiles_per = (double)p->total_tiles / (grid / cg);
                fprintf(stderr, "[igemm prof] %s | per CTA avg cycles: mma total %.0f (wait full %.0f, wait tempty %.0f) | "
                        "producer wait empty %.0f | epi warp total %.0f (wait tfull %.0f, wait res %.0f, wait store %.0f) | "
                        "%.2f tiles/cluster -> %.0f cyc/tile, %.0f cyc/slice\n", name->c_str(), a[0], a[1], a[2], a[3], a[4], a[5],
                        a[6], a[7], tiles_per, a[0] / tiles_per, a[0] / tiles_per / p->num_slices);
                fprintf(stderr, "[igemm prof]   epilogue warp 0, cycles per tile: pre %.0f | waits %.0f | tmem ld + math + sts %.0f | "
                        "hand-back %.0f | fence + TMA store %.0f | statistics %.0f\n", a[8] / tiles_per, a[9] / tiles_per,
                        a[10] / tiles_per, a[11] / tiles_per, a[12] / tiles_per, a[13] / tiles_per);
            }
        } dump{p.get(), grid, cg, st, &opname};
        if (cg == 2) {
            if (bn_tile == 256) return launch_igemm_o<256, 2>(*p, grid, f32, st);
            if (bn_tile == 128) return launch_igemm_o<128, 2>(*p, grid, f32, st);
            return launch_igemm_o<64, 2>(*p, grid, f32, st);
        }
        if (bn_tile == 256) return launch_igemm_o<256, 1>(*p, grid, f32, st);
        if (bn_tile == 128) return launch_igemm_o<128, 1>(*p, grid, f32, st);
        return launch_igemm_o<64, 1>(*p, grid, f32, st);
    };
    ops.push_back(std::move(op));
    return 0;
}

}  // namespace tq
