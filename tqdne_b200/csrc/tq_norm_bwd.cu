// tq_norm_bwd.cu -- backward of GroupNorm(32 groups) [+ SiLU] over a virtual channel concat, channels-last.
// Training-step row.  Reference: autograd through GroupNorm32 / nn.SiLU (tqdne/nn.py:11-13,90-105,
// tqdne/unet.py:85-88,100-103) inside LightningEDM.step (tqdne/edm.py:115-134).
//
//   forward   xh = (x - mu_g) rstd_g,  v = xh gamma_c + beta_c,  y = v sigmoid(v)   (or y = v without SiLU)
//   backward  dv = dy * s (1 + v (1 - s)), s = sigmoid(v)
//             dbeta_c  = sum_{n,p} dv            dgamma_c = sum_{n,p} dv xh
//             dx = rstd_g ( gamma_c dv - M1_{n,g} - xh M2_{n,g} ),
//                  M1 = mean over the group of gamma_c dv,  M2 = mean over the group of gamma_c dv xh
//
// Product path (`gn_bwd_cluster_kernel`): ONE launch, one thread-block cluster (<= 8 CTAs) per sample.  Every CTA streams
// its position chunk once for the per-channel sums A, B, the cluster exchanges them through distributed shared memory
// (fixed rank order: deterministic, no atomics, no scratch memset), and every CTA streams the SAME chunk again for dx -- a
// re-read that comes out of L2 (a chunk is ~130 KB) instead of HBM: 6 B per element of HBM traffic instead of 10, one launch
// instead of two + a memset.  The kernel is bound by its ARITHMETIC (dropout hash + SiLU derivative, ~57 instructions per
// element before this was done), so phase 1 parks dv = dy * mask * act'(v) in the dx buffer and phase 2 reads it back instead
// of evaluating either again.  The two-kernel version below stays as the fallback / A-B switch (TQ_GN_BWD_2PASS=1).
//
// HBM-bound, two streaming passes over (x, dy): pass 1 leaves A[n][c] = sum_p dv and B[n][c] = sum_p dv xh in a
// scratch buffer laid out like the forward statistics; pass 2 forms M1 / M2 from them and writes dx (10 B per element in
// bf16).  mu / rstd come from the per-(sample, channel) sums the FORWARD conv epilogue left behind (tq_conv_desc.stats).
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "tq_common.h"
#include "tq_gnstats.cuh"
#include "tq_ptx.cuh"

namespace tq {
namespace {

struct GnBwdParams {
    const void* x0; const void* x1; const void* dy;
    void* dx0; void* dx1;
    const void* add0; const void* add1;   // optional second gradient path of x0 / x1, added to dx
    int N, P, C0, C1;
    const float* gamma; const float* beta;
    float eps; int silu;
    const float* st0; const float* st1;   // forward partial sums [N][parts0][C0][2], [N][parts1][C1][2]
    int parts0, parts1;
    float* ws;                            // [N][C0 + C1][2]: A, B
    float* dgamma; float* dbeta;          // [C0 + C1], accumulated into
    const unsigned long long* drop_seed; float drop_p; int drop_site;   // the forward's fused dropout (nullptr = none)
    float* dx_sum; int dx_sum_ld;         // optional: dx_sum[n * ld + c] += sum over positions of dx0 (embedding gradient)
    float* dbias0; float* dbias1;         // optional: dbias[c] += sum over samples and positions of dx0 / dx1 (the bias gradient
                                          // of the convolution that produced x0 / x1, when dx is that tensor's whole gradient)
    int chunks;
};

template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
    using Raw = uint4;
    static __device__ __forceinline__ Raw ldraw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
    // coherent (L2) load: for data this kernel has written itself -- the read-only path may hold a stale line
    static __device__ __forceinline__ Raw ldraw_cg(const __nv_bfloat16* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
    static __device__ __forceinline__ void cvt(const Raw& u, float (&v)[8]) {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
            v[2 * k] = __low2float(b2);
            v[2 * k + 1] = __high2float(b2);
        }
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
            v[2 * k] = __low2float(b2);
            v[2 * k + 1] = __high2float(b2);
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <> struct Vec8<float> {
    struct Raw { float4 a, b; };
    static __device__ __forceinline__ Raw ldraw(const float* p) {
        return Raw{__ldg(reinterpret_cast<const float4*>(p)), __ldg(reinterpret_cast<const float4*>(p) + 1)};
    }
    static __device__ __forceinline__ Raw ldraw_cg(const float* p) {
        return Raw{__ldcg(reinterpret_cast<const float4*>(p)), __ldcg(reinterpret_cast<const float4*>(p) + 1)};
    }
    static __device__ __forceinline__ void cvt(const Raw& r, float (&v)[8]) {
        v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
    }
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};

template <typename T>
__device__ __forceinline__ float sigmoid_f(float v) {
    if constexpr (sizeof(T) == 4) return 1.f / (1.f + expf(-v));
    else {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
        return fmaf(0.5f, t, 0.5f);
    }
}

// group mean / rstd of sample n from the forward partial sums (fixed summation order: tq_gnstats.cuh)
__device__ __forceinline__ void group_stats(const GnBwdParams& p, int n, int cpg, float* gstat) {
    const GnStatSrc ss{p.st0, p.st1, p.parts0, p.parts1, p.C0, p.C1};
    gn_group_stats(ss, n, cpg, 1.f / ((float)cpg * (float)p.P), p.eps, gstat);
}

// the thread's 8-channel vector of the concatenated tensor: pointers into the right source
template <typename T>
struct VecSrc {
    const T* x; T* dx; const T* add; int C; int c_local;
};
template <typename T>
__device__ __forceinline__ VecSrc<T> pick_src(const GnBwdParams& p, int n, int c0) {
    VecSrc<T> s;
    if (c0 < p.C0) {
        s.C = p.C0; s.c_local = c0;
        s.x = static_cast<const T*>(p.x0) + (long long)n * p.P * p.C0 + c0;
        s.dx = p.dx0 ? static_cast<T*>(p.dx0) + (long long)n * p.P * p.C0 + c0 : nullptr;
        s.add = p.add0 ? static_cast<const T*>(p.add0) + (long long)n * p.P * p.C0 + c0 : nullptr;
    } else {
        s.C = p.C1; s.c_local = c0 - p.C0;
        s.x = static_cast<const T*>(p.x1) + (long long)n * p.P * p.C1 + s.c_local;
        s.dx = p.dx1 ? static_cast<T*>(p.dx1) + (long long)n * p.P * p.C1 + s.c_local : nullptr;
        s.add = p.add1 ? static_cast<const T*>(p.add1) + (long long)n * p.P * p.C1 + s.c_local : nullptr;
    }
    return s;
}

template <typename T, int PASS>
__global__ void __launch_bounds__(256) gn_bwd_kernel(const GnBwdParams p) {
    __shared__ float gstat[64];   // mean, rstd per group
    __shared__ float gm[64];      // pass 2: M1, M2 per group
    __shared__ float red[256 * 17];
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int Ct = p.C0 + p.C1, cpg = Ct / 32;
    const int cv = Ct >> 3, lanes = 256 / cv;
    const int vi = threadIdx.x % cv, pl = threadIdx.x / cv;
    const bool on = pl < lanes;
    const int c0 = vi * 8;
    group_stats(p, n, cpg, gstat);
    if constexpr (PASS == 2) {
        // M1_g = sum_{c in g} gamma_c A_c / m, M2_g = sum_{c in g} gamma_c B_c / m
        const float* ws = p.ws + (long long)n * Ct * 2;
        const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
        float m1 = 0.f, m2 = 0.f;
        for (int c = g * cpg + sub; c < (g + 1) * cpg; c += 8) {
            const float2 q = __ldcg(reinterpret_cast<const float2*>(ws + 2 * c));
            const float gmm = __ldg(p.gamma + c);
            m1 = fmaf(gmm, q.x, m1);
            m2 = fmaf(gmm, q.y, m2);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            m1 += __shfl_xor_sync(0xffffffffu, m1, o);
            m2 += __shfl_xor_sync(0xffffffffu, m2, o);
        }
        if (sub == 0) {
            const float inv = 1.f / ((float)cpg * (float)p.P);
            gm[2 * g] = m1 * inv;
            gm[2 * g + 1] = m2 * inv;
        }
    }
    __syncthreads();
    float ga[8], be[8], mu[8], rs[8], M1[8] = {}, M2[8] = {};
    if (on) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = (c0 + j) / cpg;
            ga[j] = __ldg(p.gamma + c0 + j);
            be[j] = __ldg(p.beta + c0 + j);
            mu[j] = gstat[2 * g];
            rs[j] = gstat[2 * g + 1];
            if constexpr (PASS == 2) {
                M1[j] = gm[2 * g];
                M2[j] = gm[2 * g + 1];
            }
        }
    }
    const int per = (p.P + p.chunks - 1) / p.chunks;
    const int p0 = chunk * per, p1 = min(p.P, p0 + per);
    float A[8] = {}, B[8] = {}, S[8] = {};
    if (on) {
        const VecSrc<T> s = pick_src<T>(p, n, c0);
        const T* dyb = static_cast<const T*>(p.dy) + (long long)n * p.P * Ct + c0;
        // one position: everything from (x, dy[, add]) already converted to fp32
        const bool drop = p.drop_seed != nullptr;
        const unsigned long long dseed = drop ? *p.drop_seed + 0x632BE59BD9B4E019ull * (unsigned long long)(p.drop_site + 1) : 0ull;
        const float dkeep = drop ? 1.f / (1.f - p.drop_p) : 1.f;
        auto body = [&](const float (&xv)[8], const float (&dv)[8], const float (&ad)[8], int pix) {
            float out[8];
            const long long idx0 = ((long long)n * p.P + pix) * Ct + c0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = (xv[j] - mu[j]) * rs[j];
                float d = dv[j];
                if (drop) d *= dropout_scale(dseed, idx0 + j, p.drop_p, dkeep);   // y = dropout(act(v)): dy first meets the mask
                if (p.silu) {
                    const float v = fmaf(xh, ga[j], be[j]);
                    const float sg = sigmoid_f<T>(v);
                    d *= sg * fmaf(v, 1.f - sg, 1.f);
                }
                if constexpr (PASS == 1) {
                    A[j] += d;
                    B[j] = fmaf(d, xh, B[j]);
                } else {
                    out[j] = rs[j] * (fmaf(ga[j], d, -M1[j]) - xh * M2[j]) + ad[j];
                    S[j] += out[j];
                }
            }
            if constexpr (PASS == 2) Vec8<T>::store(s.dx + (long long)pix * s.C, out);
        };
        constexpr int UN = sizeof(T) == 2 ? 4 : 2;   // positions in flight per thread, kept RAW (16 B) until consumed
        using Raw = typename Vec8<T>::Raw;
        const bool has_add = PASS == 2 && s.add != nullptr;
        int pix = p0 + pl;
        for (; pix + (UN - 1) * lanes < p1; pix += UN * lanes) {
            Raw rx[UN], rd[UN], ra[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                rx[u] = Vec8<T>::ldraw(s.x + (long long)(pix + u * lanes) * s.C);
                rd[u] = Vec8<T>::ldraw(dyb + (long long)(pix + u * lanes) * Ct);
                if (has_add) ra[u] = Vec8<T>::ldraw(s.add + (long long)(pix + u * lanes) * s.C);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                float xv[8], dv[8], ad[8] = {};
                Vec8<T>::cvt(rx[u], xv);
                Vec8<T>::cvt(rd[u], dv);
                if (has_add) Vec8<T>::cvt(ra[u], ad);
                body(xv, dv, ad, pix + u * lanes);
            }
        }
        for (; pix < p1; pix += lanes) {
            float xv[8], dv[8], ad[8] = {};
            Vec8<T>::load(s.x + (long long)pix * s.C, xv);
            Vec8<T>::load(dyb + (long long)pix * Ct, dv);
            if (has_add) Vec8<T>::load(s.add + (long long)pix * s.C, ad);
            body(xv, dv, ad, pix);
        }
    }
    if constexpr (PASS == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            red[threadIdx.x * 17 + j] = A[j];
            red[threadIdx.x * 17 + 8 + j] = B[j];
        }
        __syncthreads();
        float* ws = p.ws + (long long)n * Ct * 2;
        for (int o = threadIdx.x; o < cv * 16; o += 256) {
            const int vi2 = o >> 4, j = o & 15;
            float a = 0.f;
            for (int l = 0; l < lanes; ++l) a += red[(l * cv + vi2) * 17 + j];
            const int c = vi2 * 8 + (j & 7);
            atomicAdd(ws + 2 * c + (j >> 3), a);
            float* par = (j >> 3) ? p.dgamma : p.dbeta;   // dbeta_c = sum_n A, dgamma_c = sum_n B
            if (par) atomicAdd(par + c, a);
        }
    } else if (p.dx_sum != nullptr || p.dbias0 != nullptr || p.dbias1 != nullptr) {
        // channel sums of dx over this block's positions: same smem tree as pass 1
#pragma unroll
        for (int j = 0; j < 8; ++j) red[threadIdx.x * 17 + j] = on ? S[j] : 0.f;
        __syncthreads();
        for (int o = threadIdx.x; o < cv * 8; o += 256) {
            const int vi2 = o >> 3, j = o & 7;
            const int c = vi2 * 8 + j;
            float a = 0.f;
            for (int l = 0; l < lanes; ++l) a += red[(l * cv + vi2) * 17 + j];
            if (c < p.C0) {
                if (p.dx_sum) atomicAdd(p.dx_sum + (long long)n * p.dx_sum_ld + c, a);
                if (p.dbias0) atomicAdd(p.dbias0 + c, a);
            } else if (p.dbias1) {
                atomicAdd(p.dbias1 + (c - p.C0), a);
            }
        }
    }
}


__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr));
    return v;
}

// One cluster per sample (gridDim.x = cluster size = position chunks, blockIdx.y = sample): see the file header.
template <typename T>
__global__ void __launch_bounds__(256, 2) gn_bwd_cluster_kernel(const GnBwdParams p) {
    __shared__ float gstat[64];   // mean, rstd per group
    __shared__ float gm[64];      // M1, M2 per group
    __shared__ float red[256 * 17];
    extern __shared__ float part[];   // [Ct][2]: this CTA's per-channel (A, B) over its chunk -- read by the whole cluster
    const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int Ct = p.C0 + p.C1, cpg = Ct / 32;
    const int cv = Ct >> 3, lanes = 256 / cv;
    const int vi = threadIdx.x % cv, pl = threadIdx.x / cv;
    const bool on = pl < lanes;
    const int c0 = vi * 8;
    group_stats(p, n, cpg, gstat);
    __syncthreads();
    // per channel: xh = x rs + nmr,  v = xh gamma + beta = x gs + bs   (nmr = -mu rs, gs = gamma rs, bs = beta + gamma nmr)
    float rs[8], nmr[8], gs[8], bs[8];
    if (on) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = (c0 + j) / cpg;
            const float ga = __ldg(p.gamma + c0 + j), be = __ldg(p.beta + c0 + j);
            rs[j] = gstat[2 * g + 1];
            nmr[j] = -gstat[2 * g] * rs[j];
            gs[j] = ga * rs[j];
            bs[j] = fmaf(ga, nmr[j], be);
        }
    }
    const int per = (p.P + nchunks - 1) / nchunks;
    const int p0 = chunk * per, p1 = min(p.P, p0 + per);
    const VecSrc<T> s = pick_src<T>(p, n, on ? c0 : 0);
    const T* dyb = static_cast<const T*>(p.dy) + (long long)n * p.P * Ct + (on ? c0 : 0);
    const bool drop = p.drop_seed != nullptr;
    const unsigned long long dseed = drop ? *p.drop_seed + 0x632BE59BD9B4E019ull * (unsigned long long)(p.drop_site + 1) : 0ull;
    const float dkeep = drop ? 1.f / (1.f - p.drop_p) : 1.f;
    // dv = dL/dv of one vector (dropout mask and the SiLU derivative applied to dy), xh = normalised input
    auto grad8 = [&](const float (&xv)[8], const float (&dv)[8], int pix, float (&xh)[8], float (&d)[8]) {
        float mk[8];
        if (drop) dropout_scale8(dseed, ((long long)n * p.P + pix) * Ct + c0, p.drop_p, dkeep, mk);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            xh[j] = fmaf(xv[j], rs[j], nmr[j]);
            float dd = dv[j];
            if (drop) dd *= mk[j];
            if (p.silu) {
                const float v = fmaf(xv[j], gs[j], bs[j]);
                const float sg = sigmoid_f<T>(v);
                dd *= sg * fmaf(v, 1.f - sg, 1.f);
            }
            d[j] = dd;
        }
    };
    constexpr int UN = 2;   // positions in flight per thread (two CTAs per SM: 128 registers per thread)
    using Raw = typename Vec8<T>::Raw;
    // ---------------------------------------------------------------- phase 1: A_c = sum dv, B_c = sum dv xh over the chunk
    float A[8] = {}, B[8] = {};
    if (on) {
        int pix = p0 + pl;
        for (; pix + (UN - 1) * lanes < p1; pix += UN * lanes) {
            Raw rx[UN], rd[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                rx[u] = Vec8<T>::ldraw(s.x + (long long)(pix + u * lanes) * s.C);
                rd[u] = Vec8<T>::ldraw(dyb + (long long)(pix + u * lanes) * Ct);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                float xv[8], dv[8], xh[8], d[8];
                Vec8<T>::cvt(rx[u], xv);
                Vec8<T>::cvt(rd[u], dv);
                grad8(xv, dv, pix + u * lanes, xh, d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    A[j] += d[j];
                    B[j] = fmaf(d[j], xh[j], B[j]);
                }
                Vec8<T>::store(s.dx + (long long)(pix + u * lanes) * s.C, d);   // parked for phase 2 (see there)
            }
        }
        for (; pix < p1; pix += lanes) {
            float xv[8], dv[8], xh[8], d[8];
            Vec8<T>::load(s.x + (long long)pix * s.C, xv);
            Vec8<T>::load(dyb + (long long)pix * Ct, dv);
            grad8(xv, dv, pix, xh, d);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                A[j] += d[j];
                B[j] = fmaf(d[j], xh[j], B[j]);
            }
            Vec8<T>::store(s.dx + (long long)pix * s.C, d);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[threadIdx.x * 17 + j] = on ? A[j] : 0.f;
        red[threadIdx.x * 17 + 8 + j] = on ? B[j] : 0.f;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < cv * 16; o += 256) {
        const int vi2 = o >> 4, j = o & 15;
        float a = 0.f;
        for (int l = 0; l < lanes; ++l) a += red[(l * cv + vi2) * 17 + j];
        const int c = vi2 * 8 + (j & 7);
        part[2 * c + (j >> 3)] = a;
        float* par = (j >> 3) ? p.dgamma : p.dbeta;   // dbeta_c = sum_n A, dgamma_c = sum_n B
        if (par) atomicAdd(par + c, a);
    }
    cluster_sync_all();   // every CTA's partials are in its shared memory and visible to the cluster
    // ---------------------------------------------------------------- M1_g, M2_g from the partials of all ranks (fixed order)
    {
        const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
        const uint32_t part_addr = smem_u32(part);
        float m1 = 0.f, m2 = 0.f;
        for (int c = g * cpg + sub; c < (g + 1) * cpg; c += 8) {
            float a = 0.f, b = 0.f;
            for (int r = 0; r < nchunks; ++r) {
                const uint32_t ra = mapa_shared(part_addr + 8u * c, (uint32_t)r);
                a += ld_dsmem_f32(ra);
                b += ld_dsmem_f32(ra + 4u);
            }
            const float gmm = __ldg(p.gamma + c);
            m1 = fmaf(gmm, a, m1);
            m2 = fmaf(gmm, b, m2);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            m1 += __shfl_xor_sync(0xffffffffu, m1, o);
            m2 += __shfl_xor_sync(0xffffffffu, m2, o);
        }
        if (sub == 0) {
            const float inv = 1.f / ((float)cpg * (float)p.P);
            gm[2 * g] = m1 * inv;
            gm[2 * g + 1] = m2 * inv;
        }
    }
    cluster_sync_all();   // nobody leaves (or reuses `part`) while a peer may still read its shared memory; also orders gm
    // ---------------------------------------------------------------- phase 2: dx over the same chunk (L2-resident re-read)
    // The kernel is bound by its arithmetic, not by HBM (dropout hash + SiLU derivative: ~30 instructions per element),
    // so phase 1 parks dv -- dy with the mask and the activation derivative applied -- in the dx buffer, in the tensor's
    // own precision, and phase 2 reads it back (the same thread reads what it wrote; coherent load) instead of
    // evaluating the hash and the sigmoid a second time.
    // dx = rs (gamma d - M1 - xh M2) [+ add] = gs d - rs M1 - xh (rs M2)
    float rM1[8], rM2[8], S[8] = {};
    const bool want_sum = p.dx_sum != nullptr || p.dbias0 != nullptr || p.dbias1 != nullptr;
    if (on) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = (c0 + j) / cpg;
            rM1[j] = rs[j] * gm[2 * g];
            rM2[j] = rs[j] * gm[2 * g + 1];
        }
        const bool has_add = s.add != nullptr;
        auto emit = [&](const float (&xv)[8], const float (&d)[8], const float (&ad)[8], int pix) {
            float out[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = fmaf(xv[j], rs[j], nmr[j]);
                out[j] = fmaf(-xh, rM2[j], fmaf(gs[j], d[j], ad[j] - rM1[j]));
            }
            if (want_sum) {
#pragma unroll
                for (int j = 0; j < 8; ++j) S[j] += out[j];
            }
            Vec8<T>::store(s.dx + (long long)pix * s.C, out);
        };
        int pix = p0 + pl;
        for (; pix + (UN - 1) * lanes < p1; pix += UN * lanes) {
            Raw rx[UN], rd[UN], ra[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                rx[u] = Vec8<T>::ldraw(s.x + (long long)(pix + u * lanes) * s.C);
                rd[u] = Vec8<T>::ldraw_cg(s.dx + (long long)(pix + u * lanes) * s.C);
                if (has_add) ra[u] = Vec8<T>::ldraw(s.add + (long long)(pix + u * lanes) * s.C);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                float xv[8], dv[8], ad[8] = {};
                Vec8<T>::cvt(rx[u], xv);
                Vec8<T>::cvt(rd[u], dv);
                if (has_add) Vec8<T>::cvt(ra[u], ad);
                emit(xv, dv, ad, pix + u * lanes);
            }
        }
        for (; pix < p1; pix += lanes) {
            float xv[8], dv[8], ad[8] = {};
            Vec8<T>::load(s.x + (long long)pix * s.C, xv);
            Vec8<T>::cvt(Vec8<T>::ldraw_cg(s.dx + (long long)pix * s.C), dv);
            if (has_add) Vec8<T>::load(s.add + (long long)pix * s.C, ad);
            emit(xv, dv, ad, pix);
        }
    }
    if (want_sum) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) red[threadIdx.x * 17 + j] = on ? S[j] : 0.f;
        __syncthreads();
        for (int o = threadIdx.x; o < cv * 8; o += 256) {
            const int vi2 = o >> 3, j = o & 7;
            const int c = vi2 * 8 + j;
            float a = 0.f;
            for (int l = 0; l < lanes; ++l) a += red[(l * cv + vi2) * 17 + j];
            if (c < p.C0) {
                if (p.dx_sum) atomicAdd(p.dx_sum + (long long)n * p.dx_sum_ld + c, a);
                if (p.dbias0) atomicAdd(p.dbias0 + c, a);
            } else if (p.dbias1) {
                atomicAdd(p.dbias1 + (c - p.C0), a);
            }
        }
    }
}

template <typename T>
int launch_gn_bwd_cluster(const GnBwdParams& p, int chunks, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)chunks, (unsigned)p.N);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = (size_t)(p.C0 + p.C1) * 2 * sizeof(float);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)chunks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (chunks > 8) {
        static PerDeviceOnce np;
        if (np.first()) TQ_CUDA(cudaFuncSetAttribute(gn_bwd_cluster_kernel<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    TQ_CUDA(cudaLaunchKernelEx(&cfg, gn_bwd_cluster_kernel<T>, p));
    return 0;
}

}  // namespace
}  // namespace tq

using namespace tq;

extern "C" int tq_gn_silu_backward(const tq_gn_bwd_desc* d, void* stream) {
    TQ_CHECK(d != nullptr, "gn_silu_backward: null descriptor");
    TQ_CHECK(d->dtype == TQ_BF16 || d->dtype == TQ_F32, "gn_silu_backward: bad dtype");
    const int Ct = d->C0 + d->C1;
    TQ_CHECK(d->C0 > 0 && d->C0 % 8 == 0 && d->C1 >= 0 && d->C1 % 8 == 0 && Ct % 32 == 0 && Ct <= 2048,
             "gn_silu_backward: channel counts must be multiples of 8, their sum a multiple of 32 and <= 2048");
    TQ_CHECK(d->x0 && d->dy && d->gamma && d->beta && d->stats0 && d->ws && d->dx0, "gn_silu_backward: null pointer");
    TQ_CHECK(d->C1 == 0 || (d->x1 && d->stats1 && d->dx1), "gn_silu_backward: second source incomplete");
    TQ_CHECK(d->N > 0 && d->P > 0, "gn_silu_backward: empty tensor");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GnBwdParams p;
    p.x0 = d->x0; p.x1 = d->x1; p.dy = d->dy; p.dx0 = d->dx0; p.dx1 = d->dx1; p.add0 = d->dx_add0; p.add1 = d->dx_add1;
    p.N = d->N; p.P = d->P; p.C0 = d->C0; p.C1 = d->C1; p.gamma = d->gamma; p.beta = d->beta; p.eps = d->eps;
    p.silu = d->silu; p.st0 = d->stats0; p.st1 = d->C1 > 0 ? d->stats1 : nullptr;
    p.parts0 = d->parts0; p.parts1 = d->C1 > 0 ? d->parts1 : 1;
    TQ_CHECK(p.parts0 >= 1 && p.parts1 >= 1, "gn_silu_backward: parts0 / parts1 of the forward statistics missing"); p.ws = d->ws;
    p.drop_seed = reinterpret_cast<const unsigned long long*>(d->drop_seed); p.drop_p = d->drop_p; p.drop_site = d->drop_site;
    if (!(d->drop_p > 0.f)) p.drop_seed = nullptr;
    p.dbias0 = d->dbias0; p.dbias1 = d->C1 > 0 ? d->dbias1 : nullptr;
    p.dgamma = d->dgamma; p.dbeta = d->dbeta; p.dx_sum = d->dx_sum; p.dx_sum_ld = d->dx_sum_ld > 0 ? d->dx_sum_ld : d->C0;
    const int slots = device_sm_count() * 4;
    const int max_chunks = (d->P + 31) / 32;
    int chunks = slots / d->N;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    const char* two = getenv("TQ_GN_BWD_2PASS");
    if (!(two && two[0] == '1')) {
        // one cluster of <= 8 CTAs (the portable cluster size) per sample, both phases in one launch; two CTAs per SM
        // (128 registers per thread), one wave
        int cl = (device_sm_count() * 2) / d->N;
        if (cl > max_chunks) cl = max_chunks;
        if (cl > 8) cl = 8;
        if (cl < 1) cl = 1;
        while (cl & (cl - 1)) cl &= cl - 1;   // power of two
        if (const char* e = getenv("TQ_GN_BWD_CL")) cl = std::max(1, std::min(std::min(16, max_chunks), atoi(e)));   // experiments
        p.chunks = cl;
        if (d->dtype == TQ_F32) {
            if (launch_gn_bwd_cluster<float>(p, cl, st)) return 1;
        } else {
            if (launch_gn_bwd_cluster<__nv_bfloat16>(p, cl, st)) return 1;
        }
        TQ_CUDA(cudaGetLastError());
        count_launch(1);
        return 0;
    }
    p.chunks = chunks;
    TQ_CUDA(cudaMemsetAsync(d->ws, 0, (size_t)d->N * Ct * 2 * sizeof(float), st));
    const dim3 grid(chunks, d->N);
    if (d->dtype == TQ_F32) {
        gn_bwd_kernel<float, 1><<<grid, 256, 0, st>>>(p);
        gn_bwd_kernel<float, 2><<<grid, 256, 0, st>>>(p);
    } else {
        gn_bwd_kernel<__nv_bfloat16, 1><<<grid, 256, 0, st>>>(p);
        gn_bwd_kernel<__nv_bfloat16, 2><<<grid, 256, 0, st>>>(p);
    }
    TQ_CUDA(cudaGetLastError());
    count_launch(2);
    return 0;
}
