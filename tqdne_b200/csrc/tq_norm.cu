// tq_norm.cu -- GroupNorm(32 groups) [+ SiLU] over a virtual channel concat, channels-last.
// Reference: GroupNorm32 / normalization (tqdne/nn.py:11-13,90-105: nn.GroupNorm(32, C), eps 1e-5,
// computed in fp32) followed by nn.SiLU (tqdne/unet.py:85-88,100-103), input optionally
// th.cat([h, skip], dim=1) (tqdne/unet.py:396) -- the concat is never materialised before the norm.
//
// HBM-bound.  The per-(sample, channel) sum / sum-of-squares normally arrive from the epilogue of the
// convolution that produced the tensor (tq_conv_desc.stats: per-tile partial sums, plain stores), so the
// norm is ONE streaming pass: affine + SiLU, 16 B vector loads, four independent loads in flight per
// thread.  A stand-alone statistics pass (block reduction, one partial slot per block) remains for inputs
// that did not come out of a conv.  Nothing is accumulated with atomics: the partials are added in index
// order (tq_gnstats.cuh), so two runs of the same inputs are bit-identical.  Per-channel partials make
// groups that straddle the concat boundary (C = 768, 384, 192) free.
#include <cuda_bf16.h>

#include <memory>

#include "tq_common.h"
#include "tq_gnstats.cuh"

namespace tq {
namespace {

struct GnParams {
    const void* x0;
    const void* x1;
    int N, P, C0, C1;
    const float* gamma;
    const float* beta;
    float eps;
    int silu;
    void* y;
    float* ws;  // statistics pass: [N][chunks][C0][2] then [N][chunks][C1][2]; finalize pass: [N][32][2] behind it
    const float* st0;  // [N][parts0][C0][2]  (== ws when the statistics pass ran)
    const float* st1;  // [N][parts1][C1][2]
    int parts0, parts1;
    const float* gfinal;  // [N][32][2] group mean / rstd from gn_finalize_kernel (many parts), or nullptr
    int chunks;
    const unsigned long long* drop_seed;  // training only: dropout after the activation (nullptr = none)
    float drop_p;
    int drop_site;
    const float* film;  // FiLM (use_scale_shift_norm): row n = [scale[Ct] | shift[Ct]] of sample n, or nullptr
    int film_ld;
};

// y = GroupNorm(x) * (1 + scale[n][c]) + shift[n][c] (ResBlock with use_scale_shift_norm, tqdne/unet.py:135-139): folded into
// the per-channel affine of this block's sample before the activation
__device__ __forceinline__ void film8(const GnParams& p, int n, int c0, int Ct, float (&a)[8], float (&b)[8]) {
    const float* row = p.film + (long long)n * p.film_ld + c0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float sc = 1.f + __ldcg(row + j), sh = __ldcg(row + Ct + j);
        a[j] *= sc;
        b[j] = fmaf(b[j], sc, sh);
    }
}

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        v[2 * k] = __low2float(b2);
        v[2 * k + 1] = __high2float(b2);
    }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        w[k] = *reinterpret_cast<const uint32_t*>(&b2);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

template <typename T>
__device__ void stats_one_source(const T* x, int C, int P, int n, int chunk, int chunks, float* ws_nc, float* red) {
    const int cv = C >> 3;            // 8-channel vectors per position
    const int lanes = 256 / cv;       // position lanes
    const int tid = threadIdx.x;
    const int vi = tid % cv, pl = tid / cv;
    const int per = (P + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(P, p0 + per);
    float s[8] = {}, ss[8] = {};
    if (pl < lanes) {
        const T* base = x + ((long long)n * P) * C + vi * 8;
        for (int pix = p0 + pl; pix < p1; pix += lanes) {
            float v[8];
            load8<T>(base + (long long)pix * C, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j] += v[j];
                ss[j] = fmaf(v[j], v[j], ss[j]);
            }
        }
    }
    // red[tid][16]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[tid * 17 + j] = s[j];
        red[tid * 17 + 8 + j] = ss[j];
    }
    __syncthreads();
    // thread (vi2, j) sums over the position lanes
    for (int o = tid; o < cv * 16; o += 256) {
        const int vi2 = o >> 4, j = o & 15;
        float a = 0.f;
        for (int l = 0; l < lanes; ++l) a += red[(l * cv + vi2) * 17 + j];
        const int c = vi2 * 8 + (j & 7);
        ws_nc[2 * c + (j >> 3)] = a;   // this block's partial slot: one writer, no atomics
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const GnParams p) {
    __shared__ float red[256 * 17];
    const int n = blockIdx.y, chunk = blockIdx.x;
    // scratch layout: [N][chunks][C0][2] followed by [N][chunks][C1][2] (the layout a producing conv writes per tensor,
    // parts = chunks)
    float* ws0 = p.ws + ((long long)n * p.chunks + chunk) * p.C0 * 2;
    float* ws1 = p.ws + (long long)p.N * p.chunks * p.C0 * 2 + ((long long)n * p.chunks + chunk) * p.C1 * 2;
    stats_one_source<T>(static_cast<const T*>(p.x0), p.C0, p.P, n, chunk, p.chunks, ws0, red);
    if (p.C1 > 0) stats_one_source<T>(static_cast<const T*>(p.x1), p.C1, p.P, n, chunk, p.chunks, ws1, red);
}

template <typename T>
__device__ __forceinline__ float silu_f(float v) {
    if constexpr (sizeof(T) == 4) return v / (1.f + expf(-v));
    else {
        // x * sigmoid(x) = 0.5 x (1 + tanh(x / 2)): one MUFU op; its 2^-11 error is below bf16 rounding
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
        return v * fmaf(0.5f, t, 0.5f);
    }
}

constexpr int GN_UNROLL = 4;

// 8 consecutive channels of one position, as loaded (converted to fp32 only when consumed: half the registers
// for bf16, which lets four CTAs share an SM with four 16 B loads in flight per thread)
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
    uint4 u;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void get(float (&v)[8]) const {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
            v[2 * k] = __low2float(b2);
            v[2 * k + 1] = __high2float(b2);
        }
    }
};
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = __ldg(reinterpret_cast<const float4*>(p));
        b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    }
    __device__ __forceinline__ void get(float (&v)[8]) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};

template <typename T>
struct SrcView {
    const T* xb;   // first element of this thread's channel vector at position 0 of the sample
    T* yb;
    int C, lanes, pl, cvec0;  // cvec0: first channel (index in the concatenated tensor) of the vector
    int pix, p1;
    bool on;
};

template <typename T>
__device__ __forceinline__ SrcView<T> make_view(const T* x, int C, int cbase, int Ct, int P, int n, int chunk, int chunks, T* y) {
    SrcView<T> v;
    const int cv = C >> 3;
    v.C = C;
    v.lanes = 256 / cv;
    const int vi = threadIdx.x % cv;
    v.pl = threadIdx.x / cv;
    v.on = v.pl < v.lanes;
    const int per = (P + chunks - 1) / chunks;
    const int p0 = chunk * per;
    v.p1 = min(P, p0 + per);
    v.pix = p0 + v.pl;
    v.cvec0 = cbase + vi * 8;
    v.xb = x + ((long long)n * P) * C + vi * 8;
    v.yb = y + ((long long)n * P) * Ct + v.cvec0;
    return v;
}

// `silu` in {0, 1}; for bf16 + SiLU the caller passes a / 2 and b / 2 (exact), so that h = x a' + b' = v / 2 and
// SiLU(v) = v sigmoid(v) = h (1 + tanh(h)) = fma(h, tanh(h), h): FMA, MUFU, FMA per element instead of five operations
struct Drop {
    unsigned long long seed;
    float p, keep;
};
template <typename T, bool DROP>
__device__ __forceinline__ void emit8(const Raw8<T>& r, const float (&a)[8], const float (&b)[8], int silu, T* dst, const Drop& dr,
                                      long long idx0) {
    float v[8];
    r.get(v);
    if constexpr (sizeof(T) == 2) {
        if (silu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float h = fmaf(v[j], a[j], b[j]);
                float t;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
                v[j] = fmaf(h, t, h);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], a[j], b[j]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] = fmaf(v[j], a[j], b[j]);
            if (silu) v[j] = silu_f<T>(v[j]);
        }
    }
    if constexpr (DROP) {
        float mk[8];
        dropout_scale8(dr.seed, idx0, dr.p, dr.keep, mk);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= mk[j];
    }
    store8(dst, v);
}

template <typename T, bool DROP>
__device__ __forceinline__ void stream_source(SrcView<T>& s, const float (&a)[8], const float (&b)[8], int silu, int Ct,
                                              Raw8<T> (&pre)[GN_UNROLL], bool have_pre, const Drop& dr, long long nbase) {
    // element index of (position pix, this thread's first channel) in the [N][P][Ct] output: the dropout counter
    auto eidx = [&](int pix) { return (nbase + pix) * Ct + s.cvec0; };
    if (!s.on) return;
    const int step = GN_UNROLL * s.lanes;
    if (have_pre) {
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u)
            emit8<T, DROP>(pre[u], a, b, silu, s.yb + (long long)(s.pix + u * s.lanes) * Ct, dr, eidx(s.pix + u * s.lanes));
        s.pix += step;
    }
    for (; s.pix + (GN_UNROLL - 1) * s.lanes < s.p1; s.pix += step) {
        Raw8<T> r[GN_UNROLL];
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) r[u].load(s.xb + (long long)(s.pix + u * s.lanes) * s.C);
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u)
            emit8<T, DROP>(r[u], a, b, silu, s.yb + (long long)(s.pix + u * s.lanes) * Ct, dr, eidx(s.pix + u * s.lanes));
    }
    for (; s.pix < s.p1; s.pix += s.lanes) {
        Raw8<T> r;
        r.load(s.xb + (long long)s.pix * s.C);
        emit8<T, DROP>(r, a, b, silu, s.yb + (long long)s.pix * Ct, dr, eidx(s.pix));
    }
}

// scale / shift of 8 consecutive channels from the group statistics in shared memory
__device__ __forceinline__ void affine8(const float* gstat, int cpg, int c0, const float4 (&g4)[2], const float4 (&b4)[2],
                                        float (&a)[8], float (&b)[8]) {
    const float gm[8] = {g4[0].x, g4[0].y, g4[0].z, g4[0].w, g4[1].x, g4[1].y, g4[1].z, g4[1].w};
    const float bt[8] = {b4[0].x, b4[0].y, b4[0].z, b4[0].w, b4[1].x, b4[1].y, b4[1].z, b4[1].w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int g = (c0 + j) / cpg;
        a[j] = gstat[2 * g + 1] * gm[j];
        b[j] = bt[j] - gstat[2 * g] * a[j];
    }
}

// Tensors cut into many tiles per sample (pixel-space 128 x 128 images: 128 parts): the group reduction over
// [parts][channels of the group] is done ONCE per sample here instead of once per position chunk in the apply kernel.
// One warp per (group, sample): 32 x N warps in flight, each adding its items in a fixed order.
__global__ void __launch_bounds__(32) gn_finalize_kernel(const GnParams p, float* gfinal) {
    const int g = blockIdx.x, n = blockIdx.y;
    pdl_launch_dependents();
    pdl_wait();
    const int cpg = (p.C0 + p.C1) / 32;
    const GnStatSrc ss{p.st0, p.st1, p.parts0, p.parts1, p.C0, p.C1};
    gn_group_stats_warp(ss, n, g, cpg, 1.f / ((float)cpg * (float)p.P), p.eps, gfinal + ((long long)n * 32 + g) * 2);
}

template <typename T, bool DROP>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 4 : 2) gn_apply_kernel(const GnParams p) {
    __shared__ float gstat[64];  // [32][2] group mean, rstd
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int Ct = p.C0 + p.C1;
    const int cpg = Ct / 32;
    T* y = static_cast<T*>(p.y);
    pdl_launch_dependents();
    SrcView<T> v0 = make_view<T>(static_cast<const T*>(p.x0), p.C0, 0, Ct, p.P, n, chunk, p.chunks, y);
    // (0) gamma / beta are weights: they do not wait for the producing kernel
    float4 g4[2], b4[2];
    if (v0.on) {
        g4[0] = __ldg(reinterpret_cast<const float4*>(p.gamma + v0.cvec0));
        g4[1] = __ldg(reinterpret_cast<const float4*>(p.gamma + v0.cvec0) + 1);
        b4[0] = __ldg(reinterpret_cast<const float4*>(p.beta + v0.cvec0));
        b4[1] = __ldg(reinterpret_cast<const float4*>(p.beta + v0.cvec0) + 1);
    }
    pdl_wait();
    // (1) everything that does not depend on the statistics goes out first: the first batch of activations
    Raw8<T> pre[GN_UNROLL];
    const bool have_pre = v0.on && v0.pix + (GN_UNROLL - 1) * v0.lanes < v0.p1;
    if (have_pre) {
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) pre[u].load(v0.xb + (long long)(v0.pix + u * v0.lanes) * v0.C);
    }
    // (2) group statistics from the per-tile partial sums, added in a fixed order (tq_gnstats.cuh)
    if (p.gfinal) {
        if (threadIdx.x < 64) gstat[threadIdx.x] = __ldcg(p.gfinal + (long long)n * 64 + threadIdx.x);
    } else {
        const GnStatSrc ss{p.st0, p.st1, p.parts0, p.parts1, p.C0, p.C1};
        gn_group_stats(ss, n, cpg, 1.f / ((float)cpg * (float)p.P), p.eps, gstat);
    }
    __syncthreads();
    float a[8], b[8];
    const bool halve = sizeof(T) == 2 && p.silu;   // see emit8
    if (v0.on) {
        affine8(gstat, cpg, v0.cvec0, g4, b4, a, b);
        if (p.film) film8(p, n, v0.cvec0, Ct, a, b);
        if (halve) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] *= 0.5f; b[j] *= 0.5f; }
        }
    }
    Drop dr{0ull, 0.f, 1.f};
    if constexpr (DROP) {
        dr.seed = *p.drop_seed + 0x632BE59BD9B4E019ull * (unsigned long long)(p.drop_site + 1);
        dr.p = p.drop_p;
        dr.keep = 1.f / (1.f - p.drop_p);
    }
    const long long nbase = (long long)n * p.P;
    stream_source<T, DROP>(v0, a, b, p.silu, Ct, pre, have_pre, dr, nbase);
    if (p.C1 > 0) {
        SrcView<T> v1 = make_view<T>(static_cast<const T*>(p.x1), p.C1, p.C0, Ct, p.P, n, chunk, p.chunks, y);
        if (v1.on) {
            g4[0] = __ldg(reinterpret_cast<const float4*>(p.gamma + v1.cvec0));
            g4[1] = __ldg(reinterpret_cast<const float4*>(p.gamma + v1.cvec0) + 1);
            b4[0] = __ldg(reinterpret_cast<const float4*>(p.beta + v1.cvec0));
            b4[1] = __ldg(reinterpret_cast<const float4*>(p.beta + v1.cvec0) + 1);
            affine8(gstat, cpg, v1.cvec0, g4, b4, a, b);
            if (p.film) film8(p, n, v1.cvec0, Ct, a, b);
            if (halve) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { a[j] *= 0.5f; b[j] *= 0.5f; }
            }
        }
        stream_source<T, DROP>(v1, a, b, p.silu, Ct, pre, false, dr, nbase);
    }
}

}  // namespace

// position chunks per sample, whether the group reduction gets its own launch, and the scratch floats the op needs
struct GnLayout {
    int chunks;
    bool finalize;
    size_t final_floats, stats_floats;
};
static GnLayout gn_layout(const tq_gn_desc& d) {
    GnLayout L{};
    const int Ct = d.C0 + d.C1;
    // position chunks per sample: the grid (chunks x N CTAs) should fill whole waves of the resident-CTA slots
    // (4 CTAs/SM for bf16, 2 for fp32) -- N = 256 with 5 chunks was 2.16 waves, i.e. a third wave at 16 % occupancy
    const bool f32 = d.dtype == TQ_F32;
    const int slots = device_sm_count() * (f32 ? 2 : 4);
    const int max_chunks = (d.P + 31) / 32;
    // one wave of large chunks: every extra chunk repeats the statistics prologue (an L2 round trip + barrier)
    // before it streams; measured over the 51 norms of a UNet call at N = 256: 2 chunks 0.93 ms, 5 chunks 1.03 ms
    int chunks = slots / (d.N > 0 ? d.N : 1);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    if (const char* e = getenv("TQ_GN_CHUNKS")) {  // experiments only
        const int c = atoi(e);
        if (c >= 1 && c <= max_chunks) chunks = c;
    }
    L.chunks = chunks;
    const bool have_stats = d.stats0 != nullptr;
    const int parts = have_stats ? (d.parts0 > d.parts1 ? d.parts0 : d.parts1) : chunks;
    // more than 16 dependent-order loads per thread in the apply prologue (8 threads per group): reduce once per sample
    L.finalize = chunks > 1 && (long long)parts * (Ct / 32) > 16 * 8;
    L.final_floats = L.finalize ? (size_t)d.N * 64 : 0;
    L.stats_floats = have_stats ? 0 : (size_t)d.N * chunks * Ct * 2;
    return L;
}

int build_groupnorm(std::vector<Op>& ops, const tq_gn_desc& d) {
    TQ_CHECK(d.dtype == TQ_BF16 || d.dtype == TQ_F32, "groupnorm: bad dtype");
    const int Ct = d.C0 + d.C1;
    TQ_CHECK(d.C0 > 0 && d.C0 % 8 == 0 && d.C1 % 8 == 0 && d.C1 >= 0, "groupnorm: channel counts must be multiples of 8");
    TQ_CHECK(Ct % 32 == 0, "groupnorm: 32 groups need C %% 32 == 0 (C=%d)", Ct);
    TQ_CHECK(d.C0 <= 2048 && d.C1 <= 2048, "groupnorm: at most 2048 channels per source");
    TQ_CHECK(d.x0 && d.y && d.gamma && d.beta, "groupnorm: null pointer");
    TQ_CHECK(d.C1 == 0 || d.x1, "groupnorm: second source missing");
    const bool have_stats = d.stats0 != nullptr;
    TQ_CHECK(!have_stats || d.C1 == 0 || d.stats1, "groupnorm: statistics of the second source missing");
    TQ_CHECK(have_stats || !d.stats1, "groupnorm: statistics of the first source missing");
    TQ_CHECK(!have_stats || (d.parts0 >= 1 && (d.C1 == 0 || d.parts1 >= 1)), "groupnorm: parts0 / parts1 of the producer statistics missing");
    const GnLayout L = gn_layout(d);
    TQ_CHECK(L.final_floats + L.stats_floats == 0 || d.ws, "groupnorm: this op needs a scratch buffer (tq_groupnorm_ws_floats)");
    auto p = std::make_shared<GnParams>();
    p->x0 = d.x0; p->x1 = d.x1; p->N = d.N; p->P = d.P; p->C0 = d.C0; p->C1 = d.C1;
    p->gamma = d.gamma; p->beta = d.beta; p->eps = d.eps; p->silu = d.silu; p->y = d.y;
    const int chunks = L.chunks;
    p->chunks = chunks;
    float* gfinal = L.finalize ? d.ws : nullptr;
    p->ws = d.ws ? d.ws + L.final_floats : nullptr;
    p->st0 = have_stats ? d.stats0 : p->ws;
    p->st1 = have_stats ? d.stats1 : (d.C1 > 0 ? p->ws + (size_t)d.N * chunks * d.C0 * 2 : nullptr);
    p->parts0 = have_stats ? d.parts0 : chunks;
    p->parts1 = have_stats ? d.parts1 : chunks;
    p->gfinal = nullptr;
    const bool f32 = d.dtype == TQ_F32;
    const size_t smem = 0;
    dim3 grid(chunks, d.N);

    if (!have_stats) {
        Op st;
        st.name = f32 ? "gn_stats<f32>" : "gn_stats<bf16>";
        st.launch = [p, grid, f32](cudaStream_t s) -> int {
            if (f32) gn_stats_kernel<float><<<grid, 256, 0, s>>>(*p);
            else gn_stats_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(*p);
            TQ_CUDA(cudaGetLastError());
            count_launch();
            return 0;
        };
        ops.push_back(std::move(st));
    }
    if (L.finalize) {
        Op fn;
        fn.name = "gn_finalize";
        fn.small = true;
        const int N = d.N;
        auto pf = std::make_shared<GnParams>(*p);   // reads the partial sums (gfinal unset)
        fn.launch = [pf, gfinal, N](cudaStream_t s) -> int {
            TQ_CUDA(launch_pdl(gn_finalize_kernel, dim3(32, N), dim3(32), 0, s, *pf, gfinal));
            TQ_CUDA(cudaGetLastError());
            count_launch();
            return 0;
        };
        ops.push_back(std::move(fn));
        p->gfinal = gfinal;
    }
    Op ap;
    ap.name = f32 ? "gn_apply<f32>" : "gn_apply<bf16>";
    ap.small = (double)d.N * d.P * Ct * (f32 ? 8 : 4) < 40e6;  // < 40 MB moved: latency-bound
    const bool drop = d.drop_seed != nullptr && d.drop_p > 0.f;
    TQ_CHECK(!drop || (!f32 && d.drop_p < 1.f), "groupnorm: fused dropout is built for bf16 and p < 1");
    p->drop_seed = reinterpret_cast<const unsigned long long*>(d.drop_seed); p->drop_p = d.drop_p; p->drop_site = d.drop_site;
    p->film = d.film; p->film_ld = d.film_ld;
    ap.launch = [p, grid, f32, smem, drop](cudaStream_t s) -> int {
        if (f32) TQ_CUDA(launch_pdl(gn_apply_kernel<float, false>, grid, dim3(256), smem, s, *p));
        else if (drop) TQ_CUDA(launch_pdl(gn_apply_kernel<__nv_bfloat16, true>, grid, dim3(256), smem, s, *p));
        else TQ_CUDA(launch_pdl(gn_apply_kernel<__nv_bfloat16, false>, grid, dim3(256), smem, s, *p));
        TQ_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    };
    ops.push_back(std::move(ap));
    return 0;
}

}  // namespace tq

extern "C" int64_t tq_groupnorm_ws_floats(const tq_gn_desc* d) {
    if (!d || d->N <= 0 || d->P <= 0 || d->C0 <= 0) return -1;
    const tq::GnLayout L = tq::gn_layout(*d);
    return (int64_t)(L.final_floats + L.stats_floats);
}
