// tq_igemm_simt.cu -- FFMA implicit-GEMM convolution: the fp32 parity mode of the engine
// (north_star: "fp32 mode rel-L2 <= 1e-5"; tcgen05 has no fp32 MMA) and an on-device cross-check of
// the tensor path (instantiated for bf16 operands too).  Same tq_conv_desc semantics as
// tq_igemm_sm100.cu: K-slices of 64 channels, shifted source windows with zero fill, parity classes,
// bias / per-sample embedding / residual epilogue.  64x64 output tile per 256-thread block.
#include <cuda_bf16.h>

#include <memory>

#include "tq_common.h"

namespace tq {
namespace {

struct SimtParams {
    const void* src_ptr[4];
    int sN[4], sH[4], sW[4];
    long long s_sn[4], s_sy[4], s_sx[4];
    const int4* slices;
    int num_slices, num_classes;
    int N, H, W;
    long long M;  // N*H*W rows per class
    int cout, ktot;
    const void* weights;
    const float* bias;
    const float* emb;
    int emb_ld;
    const void* residual;
    void* out;
    long long out_sn, out_sy, out_sx;
    long long out_class_off[4];
    float* stats;  // [N][parts][cout][2] or nullptr (same meaning as in the tensor-core kernel; a part = one 64-row tile)
    int parts, parts_per_class;
    long long P;   // rows of one sample inside a class: H * W
};

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T, bool OUT_F32>
__global__ void __launch_bounds__(256) igemm_simt_kernel(const SimtParams p) {
    __shared__ float As[64][68];  // [k][m]
    __shared__ float Bs[64][68];  // [k][n]
    const int cls = blockIdx.z;
    const long long m0 = (long long)blockIdx.x * 64;
    const int n0 = blockIdx.y * 64;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;

    // loader mapping: 4 threads per row, 16 consecutive channels each
    const int lrow = tid >> 2, lpart = (tid & 3) * 16;
    const long long lm = m0 + lrow;
    const bool lvalid = lm < p.M;
    int ln = 0, ly = 0, lx = 0;
    if (lvalid) {
        lx = (int)(lm % p.W);
        long long r = lm / p.W;
        ly = (int)(r % p.H);
        ln = (int)(r / p.H);
    }
    const T* wrow = static_cast<const T*>(p.weights) + (long long)(n0 + lrow) * p.ktot;

    float acc[4][4] = {};
    const int4* sl = p.slices + (size_t)cls * p.num_slices;
    for (int s = 0; s < p.num_slices; ++s) {
        const int4 v = sl[s];
        const int src = (short)(v.x & 0xffff);
        const int dx = (short)(v.x >> 16);
        const int dy = (short)(v.y & 0xffff);
        const int c0 = v.z, kb = v.w;
        const int sx = lx + dx, sy = ly + dy;
        const bool in = lvalid && sx >= 0 && sx < p.sW[src] && sy >= 0 && sy < p.sH[src] && ln < p.sN[src];
        const T* ap = static_cast<const T*>(p.src_ptr[src]) + (long long)ln * p.s_sn[src] + (long long)sy * p.s_sy[src] +
                      (long long)sx * p.s_sx[src] + c0 + lpart;
#pragma unroll
        for (int j = 0; j < 16; ++j) As[lpart + j][lrow] = in ? to_f(ap[j]) : 0.f;
        const T* bp = wrow + (long long)kb * 64 + lpart;
#pragma unroll
        for (int j = 0; j < 16; ++j) Bs[lpart + j][lrow] = to_f(bp[j]);
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // stored values of the tile, kept for the statistics pass (As is free after the last k loop)
    float (*tile)[68] = As;  // [row][channel]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        const bool mv = m < p.M;
        const int x = mv ? (int)(m % p.W) : 0;
        const long long r = mv ? m / p.W : 0;
        const int y = (int)(r % p.H);
        const int n = (int)(r / p.H);
        const long long off = p.out_class_off[cls] + (long long)n * p.out_sn + (long long)y * p.out_sy + (long long)x * p.out_sx;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            float o = 0.f;
            if (mv && c < p.cout) {
                o = acc[i][j];
                if (p.bias) o += p.bias[c];
                if (p.emb) o += p.emb[(long long)n * p.emb_ld + c];
                if (p.residual) o += to_f(static_cast<const T*>(p.residual)[off + c]);
                if constexpr (OUT_F32) static_cast<float*>(p.out)[off + c] = o;
                else {
                    const __nv_bfloat16 ob = __float2bfloat16_rn(o);
                    static_cast<__nv_bfloat16*>(p.out)[off + c] = ob;
                    o = __bfloat162float(ob);
                }
            }
            if (p.stats) tile[ty * 4 + i][tx * 4 + j] = o;
        }
    }
    if (p.stats == nullptr) return;
    __syncthreads();
    // GroupNorm statistics of the consumer: thread c walks the 64 rows of its channel in order and leaves one
    // (sum, sum of squares) per sample the tile touches in the slot [n][part][c], part = tile index relative to the
    // sample's first tile -- one writer per slot, a fixed summation order, no atomics (bit-reproducible)
    if (tid < 64) {
        const int c = n0 + tid;
        if (c < p.cout) {
            float sm = 0.f, sq = 0.f;
            long long cur_n = m0 / p.P;
            for (int r = 0; r < 64; ++r) {
                const long long m = m0 + r;
                if (m >= p.M) break;
                const long long n = m / p.P;
                if (n != cur_n) {
                    const long long part = cls * p.parts_per_class + (blockIdx.x - (cur_n * p.P) / 64);
                    *reinterpret_cast<float2*>(p.stats + ((cur_n * p.parts + part) * p.cout + c) * 2) = make_float2(sm, sq);
                    sm = sq = 0.f;
                    cur_n = n;
                }
                const float o = tile[r][tid];
                sm += o;
                sq = fmaf(o, o, sq);
            }
            if (cur_n * p.P < p.M) {
                const long long part = cls * p.parts_per_class + (blockIdx.x - (cur_n * p.P) / 64);
                *reinterpret_cast<float2*>(p.stats + ((cur_n * p.parts + part) * p.cout + c) * 2) = make_float2(sm, sq);
            }
        }
    }
}

}  // namespace

// 64-row tiles a sample of P rows can touch: P / 64 when the samples are tile aligned, else one more than the span
static int simt_parts_per_class(long long P) { return P % 64 == 0 ? (int)(P / 64) : (int)((P - 1) / 64) + 2; }
int conv_stats_parts_simt(const tq_conv_desc& d) { return d.num_classes * simt_parts_per_class((long long)d.H * d.W); }

int build_conv_simt(std::vector<Op>& ops, const tq_conv_desc& d) {
    TQ_CHECK(d.dtype == TQ_BF16 || d.dtype == TQ_F32, "simt igemm: bad dtype");
    TQ_CHECK(d.num_srcs >= 1 && d.num_srcs <= 4, "num_srcs out of range");
    TQ_CHECK(d.num_classes == 1 || d.num_classes == 2 || d.num_classes == 4, "num_classes must be 1, 2 or 4");
    TQ_CHECK(d.ktot % 64 == 0 && d.cout_pad % 64 == 0, "weight matrix must be padded to 64x64 blocks");
    TQ_CHECK(d.out_dtype == TQ_F32 || d.out_dtype == TQ_BF16, "bad out_dtype");
    auto p = std::make_shared<SimtParams>();
    memset(p.get(), 0, sizeof(SimtParams));
    for (int i = 0; i < d.num_srcs; ++i) {
        const tq_src& s = d.srcs[i];
        TQ_CHECK(s.C % 64 == 0, "conv source channels must be a multiple of 64 (got %d)", s.C);
        p->src_ptr[i] = s.ptr;
        p->sN[i] = s.N; p->sH[i] = s.H; p->sW[i] = s.W;
        p->s_sn[i] = s.sn; p->s_sy[i] = s.sy; p->s_sx[i] = s.sx;
    }
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
    p->slices = static_cast<const int4*>(dsl);
    p->num_slices = d.num_slices; p->num_classes = d.num_classes;
    p->N = d.N; p->H = d.H; p->W = d.W;
    p->M = (long long)d.N * d.H * d.W;
    p->cout = d.cout; p->ktot = d.ktot;
    p->weights = d.weights; p->bias = d.bias; p->emb = d.emb; p->emb_ld = d.emb_ld;
    p->residual = d.residual; p->out = d.out;
    p->out_sn = d.out_sn; p->out_sy = d.out_sy; p->out_sx = d.out_sx;
    for (int i = 0; i < 4; ++i) p->out_class_off[i] = d.out_class_off[i];
    p->stats = d.stats;
    p->P = (long long)d.H * d.W;
    p->parts_per_class = simt_parts_per_class(p->P);
    p->parts = d.num_classes * p->parts_per_class;
    TQ_CHECK(d.stats == nullptr || (reinterpret_cast<uintptr_t>(d.stats) & 7) == 0, "conv statistics buffer must be 8 B aligned");
    TQ_CHECK(d.stats == nullptr || d.stats_parts == p->parts,
             "conv statistics: stats_parts = %d, this geometry writes %d parts per sample (tq_conv_stats_parts)", d.stats_parts,
             p->parts);

    dim3 grid((unsigned)((p->M + 63) / 64), (unsigned)(d.cout_pad / 64), (unsigned)d.num_classes);
    TQ_CHECK(grid.y <= 65535, "too many output channels for the simt grid");
    const bool f32in = d.dtype == TQ_F32, f32out = d.out_dtype == TQ_F32;
    Op op;
    char nm[96];
    snprintf(nm, sizeof nm, "igemm_simt<%s,%s> M=%lld slices=%d", f32in ? "f32" : "bf16", f32out ? "f32" : "bf16", p->M,
             d.num_slices);
    op.name = nm;
    op.launch = [p, dsl_owner, grid, f32in, f32out](cudaStream_t st) -> int {
        if (f32in) {
            if (f32out) igemm_simt_kernel<float, true><<<grid, 256, 0, st>>>(*p);
            else igemm_simt_kernel<float, false><<<grid, 256, 0, st>>>(*p);
        } else {
            if (f32out) igemm_simt_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>(*p);
            else igemm_simt_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>(*p);
        }
        TQ_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    };
    ops.push_back(std::move(op));
    return 0;
}

}  // namespace tq
