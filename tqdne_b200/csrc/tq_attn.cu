// tq_attn.cu -- attention core, channels-last, fp32 math with online softmax.
// Reference: QKVAttention.forward (tqdne/blocks.py:156-190): q,k,v = qkv.chunk(3, dim=1), heads split
// inside each third, w = softmax_fp32((q*s)^T (k*s)) with s = d^-1/4, a = w v.  The optional causal mask
// (use_causal_mask, blocks.py:181-186: key s > query t -> -inf; off in every shipped config) is handled by these FFMA
// kernels only: a causal attention never takes the tensor-core kernels of tq_attn_sm100.cu.
//
// One CTA = (sample, head, 16 queries); 4 warps x 4 queries.  K/V are staged in shared memory in
// 64-key blocks (row pitch d+1 floats: conflict-free both for "lane = key" score dots and
// "lane = channel" PV accumulation).  The latent config has T=16, d=128 (0.02 % of FLOPs); T=256/508
// (pixel / 1D configs) run through the same kernel.
#include <cuda_bf16.h>

#include <cstdlib>
#include <memory>

#include "tq_common.h"

namespace tq {
namespace {

constexpr int QB = 16;   // queries per CTA
constexpr int KBLK = 64;  // keys per smem block

struct AttnParams {
    const void* qkv;
    void* out;
    int N, T, heads, d;
    int causal;
};

__device__ __forceinline__ float ld_f(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_f(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_f(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_f(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T, int D>
__global__ void __launch_bounds__(128) attention_kernel(const AttnParams p) {
    extern __shared__ float sm[];
    pdl_launch_dependents();
    pdl_wait();
    constexpr int PITCH = D + 1;
    float* Ks = sm;                       // [KBLK][PITCH]
    float* Vs = Ks + KBLK * PITCH;        // [KBLK][PITCH]
    float* Qs = Vs + KBLK * PITCH;        // [QB][D]
    float* Ps = Qs + QB * D;              // [4 warps][KBLK]
    const int C = p.heads * D;
    const int ld = 3 * C;
    const int qblocks = (p.T + QB - 1) / QB;
    const int qb = blockIdx.x % qblocks;
    const int h = (blockIdx.x / qblocks) % p.heads;
    const int n = blockIdx.x / (qblocks * p.heads);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float scale = 1.f / sqrtf(sqrtf((float)D));
    const T* base = static_cast<const T*>(p.qkv) + (long long)n * p.T * ld + h * D;

    for (int i = tid; i < QB * D; i += 128) {
        const int qi = i / D, c = i % D;
        const int t = qb * QB + qi;
        Qs[i] = t < p.T ? ld_f(base + (long long)t * ld + c) * scale : 0.f;
    }
    constexpr int DPL = D / 32;  // channels per lane
    float m_run[4], l_run[4], acc[4][DPL];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_run[i] = -INFINITY;
        l_run[i] = 0.f;
#pragma unroll
        for (int j = 0; j < DPL; ++j) acc[i][j] = 0.f;
    }

    // causal: keys beyond the last query of this CTA contribute nothing
    const int k_end = p.causal ? min(p.T, qb * QB + QB) : p.T;
    for (int k0 = 0; k0 < k_end; k0 += KBLK) {
        __syncthreads();
        for (int i = tid; i < KBLK * D; i += 128) {
            const int s = i / D, c = i % D;
            const int t = k0 + s;
            float kv = 0.f, vv = 0.f;
            if (t < p.T) {
                kv = ld_f(base + (long long)t * ld + C + c) * scale;
                vv = ld_f(base + (long long)t * ld + 2 * C + c);
            }
            Ks[s * PITCH + c] = kv;
            Vs[s * PITCH + c] = vv;
        }
        __syncthreads();
#pragma unroll
        for (int qi = 0; qi < 4; ++qi) {
            const float* q = Qs + (warp * 4 + qi) * D;
            const int k_lim = p.causal ? min(p.T, qb * QB + warp * 4 + qi + 1) : p.T;   // keys [0, k_lim) are visible
            float sc[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int s = lane + 32 * j;
                float a = 0.f;
                const float* kr = Ks + s * PITCH;
#pragma unroll 8
                for (int c = 0; c < D; ++c) a = fmaf(q[c], kr[c], a);
                sc[j] = (k0 + s < k_lim) ? a : -INFINITY;
            }
            float bm = fmaxf(sc[0], sc[1]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
            const float m_new = fmaxf(m_run[qi], bm);
            const float corr = expf(m_run[qi] - m_new);  // exp(-inf) = 0 on the first block
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float e = expf(sc[j] - m_new);
                Ps[warp * KBLK + lane + 32 * j] = e;
                ps += e;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
            l_run[qi] = l_run[qi] * corr + ps;
            m_run[qi] = m_new;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < DPL; ++j) acc[qi][j] *= corr;
            const float* pw = Ps + warp * KBLK;
#pragma unroll 4
            for (int s = 0; s < KBLK; ++s) {
                const float pv = pw[s];
#pragma unroll
                for (int j = 0; j < DPL; ++j) acc[qi][j] = fmaf(pv, Vs[s * PITCH + lane + 32 * j], acc[qi][j]);
            }
            __syncwarp();
        }
    }
    T* ob = static_cast<T*>(p.out) + (long long)n * p.T * C + h * D;
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
        const int t = qb * QB + warp * 4 + qi;
        if (t >= p.T) continue;
        const float inv = 1.f / l_run[qi];
#pragma unroll
        for (int j = 0; j < DPL; ++j) st_f(ob + (long long)t * C + lane + 32 * j, acc[qi][j] * inv);
    }
}

// ---------------------------------------------------------------------------------------------------
// T <= 32 (latent config: 4x4 = 16 tokens): one CTA per (sample, head) holds q, k, v and the T x T
// probabilities in shared memory; no key blocking, no online softmax.
template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        v[2 * k] = __low2float(b2);
        v[2 * k + 1] = __high2float(b2);
    }
}

// T <= 32 (latent UNet: T = 16): one CTA per (sample, head), everything in shared memory.  The kernel is bound by
// shared-memory instruction issue, so both contractions are register-blocked over 16 B loads: a thread owns a
// (query, key pair) for QK^T and a (query, 4-channel group) for PV.
template <typename T, int D>
__global__ void __launch_bounds__(128) attention_small_kernel(const AttnParams p) {
    extern __shared__ __align__(16) float sm[];
    pdl_launch_dependents();
    pdl_wait();
    constexpr int PITCH = D + 4;     // rows stay 16 B aligned; 8 consecutive rows cover all 32 banks
    const int Tn = p.T;
    const int Tp = (Tn + 1) & ~1;    // keys padded to a pair
    float* Qs = sm;                  // [Tp][PITCH]
    float* Ks = Qs + Tp * PITCH;     // [Tp][PITCH]
    float* Vs = Ks + Tp * PITCH;     // [Tp][PITCH]
    float* Ps = Vs + Tp * PITCH;     // [Tn][Tp + 1]
    const int C = p.heads * D;
    const int ld = 3 * C;
    const int h = blockIdx.x % p.heads;
    const int n = blockIdx.x / p.heads;
    const int tid = threadIdx.x;
    const float scale = 1.f / sqrtf(sqrtf((float)D));
    const T* base = static_cast<const T*>(p.qkv) + (long long)n * Tn * ld + h * D;
    constexpr int VPR = D / 8;  // 16 B vectors per row
    for (int i = tid; i < 3 * Tp * VPR; i += 128) {
        const int which = i / (Tp * VPR);
        const int r = (i / VPR) % Tp, vc = i % VPR;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < Tn) ld8<T>(base + (long long)r * ld + which * C + vc * 8, v);
        float* dst = (which == 0 ? Qs : (which == 1 ? Ks : Vs)) + r * PITCH + vc * 8;
        const float sc = which == 2 ? 1.f : scale;
        reinterpret_cast<float4*>(dst)[0] = make_float4(v[0] * sc, v[1] * sc, v[2] * sc, v[3] * sc);
        reinterpret_cast<float4*>(dst)[1] = make_float4(v[4] * sc, v[5] * sc, v[6] * sc, v[7] * sc);
    }
    __syncthreads();
    // scores: item = (query, key pair)
    const int kpairs = Tp / 2;
    for (int i = tid; i < Tn * kpairs; i += 128) {
        const int qi = i / kpairs, k0 = (i % kpairs) * 2;
        const float4* q = reinterpret_cast<const float4*>(Qs + qi * PITCH);
        const float4* ka = reinterpret_cast<const float4*>(Ks + k0 * PITCH);
        const float4* kb = reinterpret_cast<const float4*>(Ks + (k0 + 1) * PITCH);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
        for (int c = 0; c < D / 4; ++c) {
            const float4 qv = q[c], x = ka[c], y = kb[c];
            a0 = fmaf(qv.x, x.x, a0); a0 = fmaf(qv.y, x.y, a0); a0 = fmaf(qv.z, x.z, a0); a0 = fmaf(qv.w, x.w, a0);
            a1 = fmaf(qv.x, y.x, a1); a1 = fmaf(qv.y, y.y, a1); a1 = fmaf(qv.z, y.z, a1); a1 = fmaf(qv.w, y.w, a1);
        }
        Ps[qi * (Tp + 1) + k0] = a0;
        Ps[qi * (Tp + 1) + k0 + 1] = a1;
    }
    __syncthreads();
    if (tid < Tn) {
        float* pr = Ps + tid * (Tp + 1);
        const int k_lim = p.causal ? tid + 1 : Tn;   // causal: query tid sees keys 0 .. tid
        float m = -INFINITY;
        for (int k = 0; k < k_lim; ++k) m = fmaxf(m, pr[k]);
        float l = 0.f;
        for (int k = 0; k < Tn; ++k) {
            const float e = k < k_lim ? expf(pr[k] - m) : 0.f;
            pr[k] = e;
            l += e;
        }
        const float inv = 1.f / l;
        for (int k = 0; k < Tn; ++k) pr[k] *= inv;
    }
    __syncthreads();
    // output: item = (query, 4-channel group)
    T* ob = static_cast<T*>(p.out) + (long long)n * Tn * C + h * D;
    for (int i = tid; i < Tn * (D / 4); i += 128) {
        const int qi = i / (D / 4), c4 = i % (D / 4);
        const float* pr = Ps + qi * (Tp + 1);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < Tn; ++k) {
            const float w = pr[k];
            const float4 v = *reinterpret_cast<const float4*>(Vs + k * PITCH + c4 * 4);
            a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
        }
        T* o = ob + (long long)qi * C + c4 * 4;
        st_f(o, a.x); st_f(o + 1, a.y); st_f(o + 2, a.z); st_f(o + 3, a.w);
    }
}

template <typename T, int D>
int launch_attn(const AttnParams& p, cudaStream_t st) {
    if (p.T <= 32) {
        const int tp = (p.T + 1) & ~1;
        const size_t smem = (size_t)(3 * tp * (D + 4) + p.T * (tp + 1)) * sizeof(float);
        static PerDeviceOnce attr_small;
        if (attr_small.first()) {
            TQ_CUDA(cudaFuncSetAttribute(attention_small_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        }
        TQ_CUDA(launch_pdl(attention_small_kernel<T, D>, dim3(p.N * p.heads), dim3(128), smem, st, p));
        TQ_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    }
    const size_t smem = (size_t)(2 * KBLK * (D + 1) + QB * D + 4 * KBLK) * sizeof(float);
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        TQ_CUDA(cudaFuncSetAttribute(attention_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int qblocks = (p.T + QB - 1) / QB;
    TQ_CUDA(launch_pdl(attention_kernel<T, D>, dim3(p.N * p.heads * qblocks), dim3(128), smem, st, p));
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace

int build_attention(std::vector<Op>& ops, const tq_attn_desc& d) {
    TQ_CHECK(d.dtype == TQ_BF16 || d.dtype == TQ_F32, "attention: bad dtype");
    TQ_CHECK(d.d == 64 || d.d == 128 || d.d == 32, "attention: head dim must be 32, 64 or 128 (got %d)", d.d);
    TQ_CHECK(d.N > 0 && d.T > 0 && d.heads > 0 && d.qkv && d.out, "attention: bad arguments");
    {
        // T > 32 in bf16 runs on the tensor cores; TQ_ATTN_SIMT=1 keeps the FFMA kernel (A/B and cross-check)
        const char* simt = getenv("TQ_ATTN_SIMT");
        if (!d.causal && attention_tc_supported(d) && !(simt && simt[0] == '1')) return build_attention_tc(ops, d);
    }
    auto p = std::make_shared<AttnParams>();
    p->qkv = d.qkv; p->out = d.out; p->N = d.N; p->T = d.T; p->heads = d.heads; p->d = d.d;
    p->causal = d.causal ? 1 : 0;
    const bool f32 = d.dtype == TQ_F32;
    const int dd = d.d;
    Op op;
    char nm[64];
    snprintf(nm, sizeof nm, "attention<%s,d=%d> T=%d%s", f32 ? "f32" : "bf16", dd, d.T, d.causal ? " causal" : "");
    op.name = nm;
    op.launch = [p, f32, dd](cudaStream_t st) -> int {
        if (f32) {
            if (dd == 128) return launch_attn<float, 128>(*p, st);
            if (dd == 64) return launch_attn<float, 64>(*p, st);
            return launch_attn<float, 32>(*p, st);
        }
        if (dd == 128) return launch_attn<__nv_bfloat16, 128>(*p, st);
        if (dd == 64) return launch_attn<__nv_bfloat16, 64>(*p, st);
        return launch_attn<__nv_bfloat16, 32>(*p, st);
    };
    ops.push_back(std::move(op));
    return 0;
}

}  // namespace tq
