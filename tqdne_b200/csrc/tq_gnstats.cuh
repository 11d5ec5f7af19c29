// tq_gnstats.cuh -- GroupNorm statistics from the per-tile partial sums a convolution epilogue leaves behind.
//
// Layout written by the producing conv (tq_conv_desc.stats): [N][parts][C][2] fp32, (sum, sum of squares) of the STORED
// output over the positions of sample n that fall into part k.  A part is one 128-row tile of the tensor-core kernel
// (one 64-row tile of the FFMA kernel); every slot is written by exactly one warp with a plain store -- no atomics --
// and the consumer adds the parts in index order, so the statistics (and with them every activation of the network)
// are bit-reproducible from run to run.  Reference semantics: nn.GroupNorm(32, C), eps 1e-5, fp32 (tqdne/nn.py:11-13).
#pragma once
#include <cuda_runtime.h>

namespace tq {

struct GnStatSrc {
    const float* st0;   // [N][parts0][C0][2]
    const float* st1;   // [N][parts1][C1][2] or nullptr
    int parts0, parts1;
    int C0, C1;
};

// Group mean / rstd of sample n into gstat[32][2].  256 threads: 8 threads per group walk the (channel, part) items of
// the group in a fixed order (thread `sub` takes items sub, sub + 8, ...), then a fixed xor-shuffle tree.
// The caller synchronises the block before reading gstat.
__device__ __forceinline__ void gn_group_stats(const GnStatSrc& s, int n, int cpg, float inv_count, float eps, float* gstat) {
    const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
    float sm = 0.f, sq = 0.f;
    const int ca = g * cpg, cb = ca + cpg;
#pragma unroll 1
    for (int src = 0; src < 2; ++src) {
        // channels of the group that live in this source
        const int lo = src == 0 ? ca : (ca > s.C0 ? ca : s.C0);
        const int hi = src == 0 ? (cb < s.C0 ? cb : s.C0) : cb;
        if (hi <= lo) continue;
        const int C = src == 0 ? s.C0 : s.C1;
        const int parts = src == 0 ? s.parts0 : s.parts1;
        const int cl = lo - (src == 0 ? 0 : s.C0);   // first channel, local to the source
        const int w = hi - lo;
        const float* base = (src == 0 ? s.st0 : s.st1) + (long long)n * parts * C * 2;
        const int items = w * parts;
        int it = sub;
        // four loads in flight per thread; the additions keep the item order
        for (; it + 24 < items; it += 32) {
            float2 q[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i2 = it + 8 * j;
                const int k = i2 / w, c = cl + (i2 - k * w);
                q[j] = __ldcg(reinterpret_cast<const float2*>(base + ((long long)k * C + c) * 2));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sm += q[j].x;
                sq += q[j].y;
            }
        }
        for (; it < items; it += 8) {
            const int k = it / w, c = cl + (it - k * w);
            const float2 q = __ldcg(reinterpret_cast<const float2*>(base + ((long long)k * C + c) * 2));
            sm += q.x;
            sq += q.y;
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (sub == 0) {
        const float mean = sm * inv_count;
        float var = sq * inv_count - mean * mean;
        var = var < 0.f ? 0.f : var;
        gstat[2 * g] = mean;
        gstat[2 * g + 1] = 1.f / sqrtf(var + eps);
    }
}

// The same reduction by ONE WARP for one (sample, group): lane l takes items l, l + 32, ... in order, then a fixed
// xor-shuffle tree.  Used by the stand-alone finalize kernel (grid = 32 groups x N samples) for tensors cut into many parts.
__device__ __forceinline__ void gn_group_stats_warp(const GnStatSrc& s, int n, int g, int cpg, float inv_count, float eps,
                                                    float* mean_rstd) {
    const int lane = threadIdx.x & 31;
    float sm = 0.f, sq = 0.f;
    const int ca = g * cpg, cb = ca + cpg;
#pragma unroll 1
    for (int src = 0; src < 2; ++src) {
        const int lo = src == 0 ? ca : (ca > s.C0 ? ca : s.C0);
        const int hi = src == 0 ? (cb < s.C0 ? cb : s.C0) : cb;
        if (hi <= lo) continue;
        const int C = src == 0 ? s.C0 : s.C1;
        const int parts = src == 0 ? s.parts0 : s.parts1;
        const int cl = lo - (src == 0 ? 0 : s.C0);
        const int w = hi - lo;
        const float* base = (src == 0 ? s.st0 : s.st1) + (long long)n * parts * C * 2;
        const int items = w * parts;
        int it = lane;
        for (; it + 96 < items; it += 128) {
            float2 q[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i2 = it + 32 * j;
                const int k = i2 / w, c = cl + (i2 - k * w);
                q[j] = __ldcg(reinterpret_cast<const float2*>(base + ((long long)k * C + c) * 2));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sm += q[j].x;
                sq += q[j].y;
            }
        }
        for (; it < items; it += 32) {
            const int k = it / w, c = cl + (it - k * w);
            const float2 q = __ldcg(reinterpret_cast<const float2*>(base + ((long long)k * C + c) * 2));
            sm += q.x;
            sq += q.y;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
        const float mean = sm * inv_count;
        float var = sq * inv_count - mean * mean;
        var = var < 0.f ? 0.f : var;
        mean_rstd[0] = mean;
        mean_rstd[1] = 1.f / sqrtf(var + eps);
    }
}

}  // namespace tq
