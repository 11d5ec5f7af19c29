// tq_igemm_sm100.cu -- bf16 implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel covers every dense contraction of the denoiser and decoder
// (reference: nn.Conv1d/Conv2d from tqdne/nn.py:16-24 at the call sites listed in include/tqdne_b200.h):
//
//   D[128 positions, BN channels] += A_slice[128, 64] * W_slice[BN, 64]^T   for every K-slice
//
//   warp 0   TMA producer : per slice one 4-D box load of the shifted input window (zero fill outside
//                           the image = "same" padding) + one 2-D box load of the weight block, both
//                           landing in 128B-swizzled shared memory, completion on an mbarrier
//   warp 1   MMA issuer   : one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per slice into a TMEM
//                           accumulator; tcgen05.commit releases the smem stage / publishes the tile
//   warp 2   TMEM allocator (2*BN columns: the accumulator is double buffered so the epilogue of tile
//                           i overlaps the MMAs of tile i+1)
//   warps 4-7 epilogue     : tcgen05.ld the accumulator (lane = output position), add bias, per-sample
//                           embedding and residual, convert and store channels-last
//
// Activations are channels-last, so the 128x64 A tile of a slice is the TMA box
// {64 ch, bw, bh, bn} of the [N,H,W,C] tensor with bw*bh*bn = 128: rows of 128 B, K-major,
// exactly the canonical SWIZZLE_128B UMMA operand layout.  No im2col buffer exists anywhere.
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <memory>

#include "tq_common.h"
#include "tq_ptx.cuh"

namespace tq {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;  // 16 KiB
constexpr int kThreads = 256;

struct alignas(64) IgemmParams {
    CUtensorMap amap[4];
    CUtensorMap bmap;
    const int4* slices;
    int num_slices, num_classes;
    int N, H, W, bw, bh, bn;
    int tiles_x, tiles_y, m_tiles, n_tiles, total_tiles;
    int cout;
    int vec_ok;
    const float* bias;
    const float* emb;
    int emb_ld;
    const __nv_bfloat16* residual;
    void* out;
    long long out_sn, out_sy, out_sx;
    long long out_class_off[4];
    float* stats;  // [N][cout][2] per-(sample, channel) sum / sum of squares, or nullptr
    int seg;       // rows of one sample inside an epilogue warp: min(32, bw*bh)
};

template <int BN>
struct Cfg {
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = 2 * BN;  // 128 / 256 / 512: powers of two >= 32
    static constexpr int AUX_BYTES = (2 * STAGES + 4) * 8 + 16 + BN * 4;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + AUX_BYTES;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct TileCoord {
    int cls, n_tile, x0, y0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile) {
    TileCoord t;
    t.n_tile = tile % p.n_tiles;
    int r = tile / p.n_tiles;
    int m_tile = r % p.m_tiles;
    t.cls = r / p.m_tiles;
    int tx = m_tile % p.tiles_x;
    int r2 = m_tile / p.tiles_x;
    int ty = r2 % p.tiles_y;
    int tn = r2 / p.tiles_y;
    t.x0 = tx * p.bw;
    t.y0 = ty * p.bh;
    t.n0 = tn * p.bn;
    return t;
}

template <int BN, bool OUT_F32>
__global__ void __launch_bounds__(kThreads, 1) igemm_sm100_kernel(const __grid_constant__ IgemmParams p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t* aux = smem_raw + (base - raw_addr) + C::STAGES * C::STAGE_BYTES;
    const uint32_t bar0 = base + C::STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(aux + (2 * C::STAGES + 4) * 8);
    float* bias_s = reinterpret_cast<float*>(aux + (2 * C::STAGES + 4) * 8 + 16);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.amap[i]);
        tma_prefetch_desc(&p.bmap);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 128);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord t = decode_tile(p, tile);
                const int4* sl = p.slices + (size_t)t.cls * p.num_slices;
                for (int s = 0; s < p.num_slices; ++s) {
                    const int4 v = __ldg(sl + s);
                    const int src = (short)(v.x & 0xffff);
                    const int dx = (short)(v.x >> 16);
                    const int dy = (short)(v.y & 0xffff);
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    const uint32_t a_dst = base + stage * C::STAGE_BYTES;
                    tma_load_4d(a_dst, &p.amap[src], full_bar(stage), v.z, t.x0 + dx, t.y0 + dy, t.n0);
                    tma_load_2d(a_dst + A_BYTES, &p.bmap, full_bar(stage), v.w * BK, t.n_tile * BN);
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int s = 0; s < p.num_slices; ++s) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_addr = base + stage * C::STAGE_BYTES;
                    const uint64_t da = umma_desc_sw128(a_addr);
                    const uint64_t db = umma_desc_sw128(a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // +32 B per K=16 step inside the 128 B swizzle row (address field is >> 4)
                        umma_bf16(d_tmem, da + 2u * k, db + 2u * k, idesc, (s | k) != 0);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tfull_bar(acc));
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (128 threads)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 128;
        const int dxr = row % p.bw;
        const int dyr = (row / p.bw) % p.bh;
        const int dnr = row / (p.bw * p.bh);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile);
            named_bar_sync(1, 128);
            for (int i = et; i < BN; i += 128) {
                const int c = t.n_tile * BN + i;
                bias_s[i] = (p.bias != nullptr && c < p.cout) ? __ldg(p.bias + c) : 0.f;
            }
            named_bar_sync(1, 128);
            const int n = t.n0 + dnr, y = t.y0 + dyr, x = t.x0 + dxr;
            const bool valid = (n < p.N) && (y < p.H) && (x < p.W);
            const long long off = p.out_class_off[t.cls] + (long long)n * p.out_sn + (long long)y * p.out_sy +
                                  (long long)x * p.out_sx;
            const float* emb_row = p.emb ? p.emb + (long long)n * p.emb_ld : nullptr;

            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + acc * BN + c + (uint32_t(q * 32) << 16), r);
                tmem_ld_wait();
                if (c + 32 >= BN) {  // accumulator fully read: hand the TMEM buffer back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(tempty_bar(acc));
                }
                const int cg = t.n_tile * BN + c;
                if (cg >= p.cout) continue;  // warp-uniform: zero-padded weight rows
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias_s[c + j];
                if (p.vec_ok) {
                    if (valid && emb_row) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 e = __ldg(reinterpret_cast<const float4*>(emb_row + cg + j));
                            v[j] += e.x; v[j + 1] += e.y; v[j + 2] += e.z; v[j + 3] += e.w;
                        }
                    }
                    if (valid && p.residual) {
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
                    if constexpr (OUT_F32) {
                        if (valid) {
                            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + off + cg);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                    } else {
                        uint32_t w[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
                            w[k] = *reinterpret_cast<const uint32_t*>(&b2);
                            // the statistics describe the tensor as stored (bf16-rounded)
                            v[2 * k] = __low2float(b2);
                            v[2 * k + 1] = __high2float(b2);
                        }
                        if (valid) {
                            uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + off + cg);
#pragma unroll
                            for (int j = 0; j < 4; ++j) op[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                        }
                    }
                    if (p.stats != nullptr) {
                        // GroupNorm statistics of the consumer, fused here: per-(sample, channel) sum and
                        // sum of squares over this warp's rows.  Recursive halving across lanes: every step
                        // trades half of the channels with the partner lane, so after log2(seg) steps a lane
                        // owns 32/seg channels summed over the seg rows of its sample.
                        float q[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            v[j] = valid ? v[j] : 0.f;
                            q[j] = v[j] * v[j];
                        }
                        int cb = 0;
#pragma unroll
                        for (int s = 0; s < 5; ++s) {
                            const int nh = 16 >> s;
                            const int o = p.seg >> (s + 1);
                            if (o == 0) break;
                            const bool upper = (lane & o) != 0;
                            cb += upper ? nh : 0;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if (j < nh) {
                                    const float sv = upper ? v[j] : v[j + nh];
                                    const float kv = upper ? v[j + nh] : v[j];
                                    const float sq = upper ? q[j] : q[j + nh];
                                    const float kq = upper ? q[j + nh] : q[j];
                                    v[j] = kv + __shfl_xor_sync(0xffffffffu, sv, o);
                                    q[j] = kq + __shfl_xor_sync(0xffffffffu, sq, o);
                                }
                            }
                        }
                        const int cnt = 32 / p.seg;
                        if (n < p.N) {
                            float2* sp = reinterpret_cast<float2*>(p.stats) + (long long)n * p.cout + cg + cb;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < cnt) atomicAdd(sp + j, make_float2(v[j], q[j]));
                        }
                    }
                } else if (valid) {
                    // ragged / unaligned edge (Cout in {3, 6, 8}): scalar path
#pragma unroll 1
                    for (int j = 0; j < 32; ++j) {
                        if (cg + j >= p.cout) break;
                        float o = v[j];
                        if (emb_row) o += __ldg(emb_row + cg + j);
                        if (p.residual) o += __bfloat162float(p.residual[off + cg + j]);
                        if constexpr (OUT_F32) static_cast<float*>(p.out)[off + cg + j] = o;
                        else static_cast<__nv_bfloat16*>(p.out)[off + cg + j] = __float2bfloat16_rn(o);
                    }
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
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

int encode_src_map(CUtensorMap* m, const tq_src& s, int bw, int bh, int bn) {
    auto enc = get_encode_fn();
    TQ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    TQ_CHECK(s.C % 64 == 0, "conv source channels must be a multiple of 64 (got %d)", s.C);
    TQ_CHECK(s.sx % 8 == 0 && s.sn % 8 == 0 && (s.H == 1 || s.sy % 8 == 0), "conv source strides must be 16 B multiples");
    TQ_CHECK((reinterpret_cast<uintptr_t>(s.ptr) & 15) == 0, "conv source pointer must be 16 B aligned");
    cuuint64_t dims[4] = {(cuuint64_t)s.C, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.N};
    long long sy = (s.H == 1 && s.sy == 0) ? (long long)s.W * s.sx : s.sy;
    cuuint64_t strides[3] = {(cuuint64_t)s.sx * 2, (cuuint64_t)sy * 2, (cuuint64_t)s.sn * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(s.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TQ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed with CUresult %d", (int)r);
    return 0;
}

int encode_weight_map(CUtensorMap* m, const void* w, int cout_pad, int ktot, int bn_tile) {
    auto enc = get_encode_fn();
    TQ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn_tile};
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

template <int BN, bool OUT_F32>
int launch_igemm(const IgemmParams& p, int grid, cudaStream_t st) {
    using C = Cfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        TQ_CUDA(cudaFuncSetAttribute(igemm_sm100_kernel<BN, OUT_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES));
        attr_set = true;
    }
    igemm_sm100_kernel<BN, OUT_F32><<<grid, kThreads, C::SMEM_BYTES, st>>>(p);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
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

int build_conv_sm100(std::vector<Op>& ops, const tq_conv_desc& d) {
    TQ_CHECK(d.dtype == TQ_BF16, "sm100 igemm needs bf16 operands");
    TQ_CHECK(d.num_srcs >= 1 && d.num_srcs <= 4, "num_srcs out of range");
    TQ_CHECK(d.num_classes == 1 || d.num_classes == 2 || d.num_classes == 4, "num_classes must be 1, 2 or 4");
    TQ_CHECK(d.num_slices >= 1, "conv needs at least one K slice");
    TQ_CHECK(d.ktot % 64 == 0 && d.cout_pad % 64 == 0, "weight matrix must be padded to 64x64 blocks");
    TQ_CHECK(d.cout >= 1 && d.cout <= d.cout_pad, "cout out of range");
    TQ_CHECK((reinterpret_cast<uintptr_t>(d.weights) & 15) == 0, "weights must be 16 B aligned");

    auto p = std::make_shared<IgemmParams>();
    memset(p.get(), 0, sizeof(IgemmParams));
    p->N = d.N; p->H = d.H; p->W = d.W;
    tile_shape_for(d.H, d.W, &p->bw, &p->bh, &p->bn);
    int bn_tile = d.block_n;
    if (bn_tile == 0) {
        bn_tile = d.cout_pad % 256 == 0 ? 256 : (d.cout_pad % 128 == 0 ? 128 : 64);
        // prefer enough tiles to fill the machine
        long long m_tiles = (long long)((d.W + p->bw - 1) / p->bw) * ((d.H + p->bh - 1) / p->bh) *
                            ((d.N + p->bn - 1) / p->bn) * d.num_classes;
        const int sms = device_sm_count();
        while (bn_tile > 64 && m_tiles * (d.cout_pad / bn_tile) < 2LL * sms && d.cout_pad % (bn_tile / 2) == 0)
            bn_tile /= 2;
    }
    TQ_CHECK(bn_tile == 64 || bn_tile == 128 || bn_tile == 256, "block_n must be 64, 128 or 256");
    TQ_CHECK(d.cout_pad % bn_tile == 0, "cout_pad must be a multiple of block_n");

    for (int i = 0; i < d.num_srcs; ++i)
        if (encode_src_map(&p->amap[i], d.srcs[i], p->bw, p->bh, p->bn)) return 1;
    for (int i = d.num_srcs; i < 4; ++i) p->amap[i] = p->amap[0];
    if (encode_weight_map(&p->bmap, d.weights, d.cout_pad, d.ktot, bn_tile)) return 1;

    const size_t nsl = (size_t)d.num_classes * d.num_slices;
    for (size_t i = 0; i < nsl; ++i) {
        const tq_slice& s = d.slices[i];
        TQ_CHECK(s.src >= 0 && s.src < d.num_srcs, "slice %zu: bad source index", i);
        TQ_CHECK(s.c0 % 64 == 0 && s.c0 + 64 <= d.srcs[s.src].C, "slice %zu: bad channel offset", i);
        TQ_CHECK(s.kb >= 0 && (s.kb + 1) * 64 <= d.ktot, "slice %zu: bad weight block", i);
    }
    void* dsl = nullptr;
    TQ_CUDA(cudaMalloc(&dsl, nsl * sizeof(tq_slice)));
    std::shared_ptr<void> dsl_owner(dsl, [](void* q) { cudaFree(q); });
    TQ_CUDA(cudaMemcpy(dsl, d.slices, nsl * sizeof(tq_slice), cudaMemcpyHostToDevice));
    static_assert(sizeof(tq_slice) == sizeof(int4), "tq_slice must be 16 bytes");

    p->slices = static_cast<const int4*>(dsl);
    p->num_slices = d.num_slices;
    p->num_classes = d.num_classes;
    p->tiles_x = (d.W + p->bw - 1) / p->bw;
    p->tiles_y = (d.H + p->bh - 1) / p->bh;
    const int tiles_n = (d.N + p->bn - 1) / p->bn;
    p->m_tiles = p->tiles_x * p->tiles_y * tiles_n;
    p->n_tiles = d.cout_pad / bn_tile;
    p->total_tiles = p->m_tiles * p->n_tiles * d.num_classes;
    p->cout = d.cout;
    p->bias = d.bias;
    p->emb = d.emb;
    p->emb_ld = d.emb_ld;
    p->residual = static_cast<const __nv_bfloat16*>(d.residual);
    p->out = d.out;
    p->out_sn = d.out_sn; p->out_sy = d.out_sy; p->out_sx = d.out_sx;
    for (int i = 0; i < 4; ++i) p->out_class_off[i] = d.out_class_off[i];
    const int oalign = d.out_dtype == TQ_F32 ? 4 : 8;
    bool vec = d.cout % 32 == 0 && d.out_sn % oalign == 0 && d.out_sy % oalign == 0 && d.out_sx % oalign == 0;
    vec = vec && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0;
    for (int i = 0; i < d.num_classes; ++i) vec = vec && d.out_class_off[i] % oalign == 0;
    if (d.emb) vec = vec && d.emb_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(d.emb) & 15) == 0;
    if (d.residual) vec = vec && (reinterpret_cast<uintptr_t>(d.residual) & 15) == 0 && d.out_sn % 8 == 0 &&
                          d.out_sy % 8 == 0 && d.out_sx % 8 == 0;
    p->vec_ok = vec ? 1 : 0;
    p->stats = d.stats;
    {
        const int rows = p->bw * p->bh;
        p->seg = rows < 32 ? rows : 32;
    }
    TQ_CHECK(d.stats == nullptr || (vec && (reinterpret_cast<uintptr_t>(d.stats) & 7) == 0),
             "conv statistics need a vectorisable epilogue (cout %% 32 == 0, aligned strides) and an 8 B aligned buffer");

    const int grid = p->total_tiles < device_sm_count() ? p->total_tiles : device_sm_count();
    const bool f32 = d.out_dtype == TQ_F32;
    TQ_CHECK(d.out_dtype == TQ_F32 || d.out_dtype == TQ_BF16, "bad out_dtype");

    Op op;
    char nm[96];
    snprintf(nm, sizeof nm, "igemm_sm100<BN=%d,%s> tiles=%d slices=%d", bn_tile, f32 ? "f32" : "bf16", p->total_tiles,
             d.num_slices);
    op.name = nm;
    op.launch = [p, dsl_owner, grid, bn_tile, f32](cudaStream_t st) -> int {
        if (bn_tile == 256) return f32 ? launch_igemm<256, true>(*p, grid, st) : launch_igemm<256, false>(*p, grid, st);
        if (bn_tile == 128) return f32 ? launch_igemm<128, true>(*p, grid, st) : launch_igemm<128, false>(*p, grid, st);
        return f32 ? launch_igemm<64, true>(*p, grid, st) : launch_igemm<64, false>(*p, grid, st);
    };
    ops.push_back(std::move(op));
    return 0;
}

}  // namespace tq
