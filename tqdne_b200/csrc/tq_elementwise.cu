// tq_elementwise.cu -- the small fp32 pieces around the denoiser: embedding MLPs, Fourier features,
// EDM preconditioning + Heun/Euler state update (fp64 state like the reference), layout changes and the
// moving-average-envelope inverse.  All HBM / latency bound, vectorised where the shapes allow.
#include <cuda_bf16.h>

#include <memory>

#include "tq_common.h"

namespace tq {
namespace {

__device__ __forceinline__ float silu_acc(float v) { return v / (1.f + expf(-v)); }

// ---------------------------------------------------------------------------------------------------
// y[m][j] = sum_k act_in(x[m][k]) W[j][k] + b[j] (+ add[m or 0][j]);  y_act = SiLU(y) in fp32 / bf16
// Reference: nn.Linear + nn.SiLU of time_mlp / cond_mlp (tqdne/unet.py:210-227), applied as
// emb = time_mlp(time_embed(t)); emb += cond_mlp(cond) (unet.py:383-388); SiLU of emb_layers (unet.py:92).
struct LinParams {
    int M, K, Nout, x_rows;
    const float* x;
    const float* W;
    const float* b;
    int act_in;
    const float* add;
    int add_rows;
    float* y;
    void* y_act;
    int y_act_dtype;
};

__device__ __forceinline__ void linear_epilogue(const LinParams& p, float a, int m, int j) {
    if (p.b) a += p.b[j];
    if (p.add) a += p.add[(long long)(p.add_rows == 1 ? 0 : m) * p.Nout + j];
    const long long o = (long long)m * p.Nout + j;
    if (p.y) p.y[o] = a;
    if (p.y_act) {
        const float s = silu_acc(a);
        if (p.y_act_dtype == TQ_F32) static_cast<float*>(p.y_act)[o] = s;
        else static_cast<__nv_bfloat16*>(p.y_act)[o] = __float2bfloat16_rn(s);
    }
}

// one warp per dot product; with a single shared input row (x_rows == 1: every sample of a sampler step has the
// same noise level) the dot product is computed once per output column and the epilogue fans out over the rows
__global__ void __launch_bounds__(256) linear_kernel(const LinParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp_global = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const bool shared_row = p.x_rows == 1;
    const long long total = shared_row ? p.Nout : (long long)p.M * p.Nout;
    if (warp_global >= total) return;
    const int m = shared_row ? 0 : warp_global / p.Nout, j = shared_row ? warp_global : warp_global % p.Nout;
    const float* xr = p.x + (long long)m * p.K;
    const float* wr = p.W + (long long)j * p.K;
    float a = 0.f;
    for (int k = lane; k < p.K; k += 32) {
        float xv = xr[k];
        if (p.act_in) xv = silu_acc(xv);
        a = fmaf(xv, __ldg(wr + k), a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (shared_row) {
        for (int mm = lane; mm < p.M; mm += 32) linear_epilogue(p, a, mm, j);
    } else if (lane == 0) {
        linear_epilogue(p, a, m, j);
    }
}

// Many input rows (the training step: M = batch rows, Nout up to ~4000 for the concatenated emb_layers): one warp per
// output COLUMN keeps its weight row in registers and walks all M rows, whose activated values the block stages in shared
// memory in chunks of LIN_MB rows -- the weight matrix is read once instead of once per row (110 -> ~10 us at 64 x 256 x 4032).
constexpr int LIN_MB = 16;
constexpr int LIN_KMAX = 512;   // K / 32 weight registers per lane
__global__ void __launch_bounds__(256) linear_rows_kernel(const LinParams p) {
    extern __shared__ float xs[];   // [LIN_MB][K]
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    const int K = p.K, kr = (K + 31) / 32;
    float w[LIN_KMAX / 32];
#pragma unroll
    for (int q = 0; q < LIN_KMAX / 32; ++q) {
        const int k = lane + 32 * q;
        w[q] = (q < kr && k < K && j < p.Nout) ? __ldg(p.W + (long long)j * K + k) : 0.f;
    }
    for (int m0 = 0; m0 < p.M; m0 += LIN_MB) {
        const int mb = min(LIN_MB, p.M - m0);
        __syncthreads();
        for (int i = threadIdx.x; i < mb * K; i += 256) {
            float xv = p.x[(long long)m0 * K + i];
            if (p.act_in) xv = silu_acc(xv);
            xs[i] = xv;
        }
        __syncthreads();
        if (j < p.Nout) {
            float res = 0.f;   // lane mm keeps the result of row m0 + mm: the epilogues of a chunk then run side by side
            for (int mm = 0; mm < mb; ++mm) {
                const float* xr = xs + mm * K;
                float a = 0.f;
#pragma unroll
                for (int q = 0; q < LIN_KMAX / 32; ++q) {
                    const int k = lane + 32 * q;
                    if (q < kr && k < K) a = fmaf(xr[k], w[q], a);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == mm) res = a;
            }
            if (lane < mb) linear_epilogue(p, res, m0 + lane, j);
        }
    }
}

// Reference: GaussianFourierProjection.forward (tqdne/blocks.py:22-26): h = x[:,None]*W[None,:]*2*pi;
// cat([sin h, cos h]).
__global__ void fourier_kernel(const float* t, const float* W, int M, int half, float* feat) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * half) return;
    const int m = i / half, k = i % half;
    const float h = t[m] * W[k] * 2.f * 3.14159265358979323846f;
    feat[(long long)m * 2 * half + k] = sinf(h);
    feat[(long long)m * 2 * half + half + k] = cosf(h);
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void st_as(T* p, float v);
template <>
__device__ __forceinline__ void st_as<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_as<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// Reference: LightningEDM.forward (tqdne/edm.py:105-113): sample_in = sample.to(fp32) * c_in(sigma)
template <typename T>
__global__ void precondition_kernel(const double* x, T* xin, long long NP, int C, int Cpad, float c_in, float* t_next,
                                    float t_value) {
    if (t_next != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *t_next = t_value;   // the NEXT denoiser call's c_noise
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP * Cpad) return;
    const long long r = i / Cpad;
    const int c = (int)(i % Cpad);
    const float v = c < C ? (float)x[r * C + c] * c_in : 0.f;
    st_as<T>(xin + i, v);
}

// Reference: Heun step 1 (tqdne/edm.py:176-183) with the denoiser post-scaling folded in
// (edm.py:111-113): D = F*c_out + c_skip*x32;  d = (x - D)/sigma;  x1 = x + d*dt  (fp64 like the reference)
template <typename T>
__global__ void euler_kernel(const double* x, const float* F, int Cf, double* d, double* x1, T* xin, long long NP, int C,
                             int Cpad, float c_out, float c_skip, float sigma, float dt, float c_in_next, int write_xin,
                             float* t_next, float t_value) {
    if (t_next != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *t_next = t_value;   // the NEXT denoiser call's c_noise
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP * Cpad) return;
    const long long r = i / Cpad;
    const int c = (int)(i % Cpad);
    if (c >= C) {
        if (write_xin) st_as<T>(xin + i, 0.f);
        return;
    }
    const double xv = x[r * C + c];
    const float x32 = (float)xv;
    const float D = __fadd_rn(__fmul_rn(F[r * Cf + c], c_out), __fmul_rn(c_skip, x32));  // no FMA: torch rounds each op
    const double dv = (xv - (double)D) / (double)sigma;
    const double xn = xv + dv * (double)dt;
    if (d) d[r * C + c] = dv;
    x1[r * C + c] = xn;
    if (write_xin) st_as<T>(xin + i, (float)xn * c_in_next);
}

// Reference: Heun 2nd-order correction (tqdne/edm.py:186-194):
// d' = (x1 - D')/sigma';  x = x + dt*(0.5 d + 0.5 d')
template <typename T>
__global__ void heun_kernel(double* x, const double* x1, const double* d, const float* F, int Cf, T* xin, long long NP,
                            int C, int Cpad, float c_out, float c_skip, float sigma_next, float dt, float c_in_next,
                            int write_xin, float* t_next, float t_value) {
    if (t_next != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *t_next = t_value;   // the NEXT denoiser call's c_noise
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP * Cpad) return;
    const long long r = i / Cpad;
    const int c = (int)(i % Cpad);
    if (c >= C) {
        if (write_xin) st_as<T>(xin + i, 0.f);
        return;
    }
    const double x1v = x1[r * C + c];
    const float x32 = (float)x1v;
    const float D = __fadd_rn(__fmul_rn(F[r * Cf + c], c_out), __fmul_rn(c_skip, x32));  // no FMA: torch rounds each op
    const double dp = (x1v - (double)D) / (double)sigma_next;
    const double xn = x[r * C + c] + (double)dt * (0.5 * d[r * C + c] + 0.5 * dp);
    x[r * C + c] = xn;
    if (write_xin) st_as<T>(xin + i, (float)xn * c_in_next);
}

// Grouped variants (Cpad % 8 == 0): one thread per 8 consecutive padded channels of a position -- the padded
// denoiser input leaves as one 16 B (bf16) / two 16 B (fp32) stores, and no thread exists per padding element
// (the per-element kernels above launched NP * Cpad threads: 16.8 M for the 8-channel latent state padded to 64).
template <typename T>
__device__ __forceinline__ void store_group8(T* dst, const float (&v)[8]);
template <>
__device__ __forceinline__ void store_group8<float>(float* dst, const float (&v)[8]) {
    reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store_group8<__nv_bfloat16>(__nv_bfloat16* dst, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        w[k] = *reinterpret_cast<const uint32_t*>(&b2);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}

// MODE 0: precondition, 1: Euler step, 2: Heun correction (same arithmetic, op for op, as the kernels above)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) edm_group_kernel(double* x, double* x1, double* d, const float* F, int Cf, T* xin,
                                                        long long NP, int C, int Cpad, float c_out, float c_skip, float sigma,
                                                        float dt, float c_in_next, int write_xin, float* t_next, float t_value) {
    if (t_next != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *t_next = t_value;   // the NEXT denoiser call's c_noise
    const int gpr = Cpad >> 3;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP * gpr) return;
    const long long r = i / gpr;
    const int c0 = (int)(i - r * gpr) * 8;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        o[j] = 0.f;
        const int c = c0 + j;
        if (c < C) {
            const long long e = r * C + c;
            if constexpr (MODE == 0) {
                o[j] = (float)x[e] * c_in_next;
            } else if constexpr (MODE == 1) {
                const double xv = x[e];
                const float D = __fadd_rn(__fmul_rn(F[r * Cf + c], c_out), __fmul_rn(c_skip, (float)xv));
                const double dv = (xv - (double)D) / (double)sigma;
                const double xn = xv + dv * (double)dt;
                if (d) d[e] = dv;
                x1[e] = xn;
                o[j] = (float)xn * c_in_next;
            } else {
                const double x1v = x1[e];
                const float D = __fadd_rn(__fmul_rn(F[r * Cf + c], c_out), __fmul_rn(c_skip, (float)x1v));
                const double dp = (x1v - (double)D) / (double)sigma;
                const double xn = x[e] + (double)dt * (0.5 * d[e] + 0.5 * dp);
                x[e] = xn;
                o[j] = (float)xn * c_in_next;
            }
        }
    }
    if (write_xin) store_group8<T>(xin + r * Cpad + c0, o);
}

template <int MODE>
int launch_edm_group(double* x, double* x1, double* d, const float* F, int Cf, void* xin, int dtype, long long NP, int C,
                     int Cpad, float c_out, float c_skip, float sigma, float dt, float c_in_next, int write_xin,
                     float* t_next, float t_value, cudaStream_t st) {
    const long long threads = write_xin ? NP * (Cpad >> 3) : NP * ((C + 7) >> 3);
    const unsigned g = (unsigned)((threads + 255) / 256);
    // without the padded output only the groups that hold real channels are needed
    const int cp = write_xin ? Cpad : ((C + 7) >> 3) * 8;
    if (dtype == TQ_F32)
        edm_group_kernel<float, MODE><<<g, 256, 0, st>>>(x, x1, d, F, Cf, static_cast<float*>(xin), NP, C, cp, c_out, c_skip,
                                                         sigma, dt, c_in_next, write_xin, t_next, t_value);
    else
        edm_group_kernel<__nv_bfloat16, MODE><<<g, 256, 0, st>>>(x, x1, d, F, Cf, static_cast<__nv_bfloat16*>(xin), NP, C, cp,
                                                                 c_out, c_skip, sigma, dt, c_in_next, write_xin, t_next, t_value);
    return 0;
}
inline bool group_ok(const void* xin, int Cpad, int dtype) {
    return Cpad % 8 == 0 && (reinterpret_cast<uintptr_t>(xin) & 15) == 0 && (dtype == TQ_F32 || dtype == TQ_BF16);
}

__global__ void add_noise_kernel(double* x, const double* noise, double scale, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += noise[i] * scale;
}

// ---------------------------------------------------------------------------------------------------
template <typename TI>
__device__ __forceinline__ double ld_d(const TI* p);
template <> __device__ __forceinline__ double ld_d<double>(const double* p) { return *p; }
template <> __device__ __forceinline__ double ld_d<float>(const float* p) { return (double)*p; }
template <> __device__ __forceinline__ double ld_d<__nv_bfloat16>(const __nv_bfloat16* p) { return (double)__bfloat162float(*p); }
template <typename TO>
__device__ __forceinline__ void st_d(TO* p, double v);
template <> __device__ __forceinline__ void st_d<double>(double* p, double v) { *p = v; }
template <> __device__ __forceinline__ void st_d<float>(float* p, double v) { *p = (float)v; }
template <> __device__ __forceinline__ void st_d<__nv_bfloat16>(__nv_bfloat16* p, double v) { *p = __float2bfloat16_rn((float)v); }

// [N,C,P] -> [N,P,Cpad]: 32x32 smem transpose tiles over (c, p)
template <typename TI, typename TO>
__global__ void nchw_to_nhwc_kernel(const TI* src, TO* dst, int C, long long P, int Cpad) {
    __shared__ double tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j;
        const long long pp = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && pp < P) ? ld_d<TI>(src + ((long long)n * C + c) * P + pp) : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long pp = p0 + j;
        const int c = c0 + threadIdx.x;
        if (pp < P && c < Cpad) st_d<TO>(dst + ((long long)n * P + pp) * Cpad + c, tile[threadIdx.x][j]);
    }
}
// [N,P,Cld] (first C channels) -> [N,C,P]
template <typename TI, typename TO>
__global__ void nhwc_to_nchw_kernel(const TI* src, int Cld, TO* dst, int C, long long P) {
    __shared__ double tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long pp = p0 + j;
        const int c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && pp < P) ? ld_d<TI>(src + ((long long)n * P + pp) * Cld + c) : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j;
        const long long pp = p0 + threadIdx.x;
        if (c < C && pp < P) st_d<TO>(dst + ((long long)n * C + c) * P + pp, tile[threadIdx.x][j]);
    }
}

// Reference: MovingAverageEnvelope.invert_representation (tqdne/representation.py:57-60)
__global__ void mavg_inverse_kernel(const float* rep, float* wave, int Cw, long long L, long long total, double half_log_eps,
                                    double eps) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long l = i % L;
    const long long r = i / L;
    const int c = (int)(r % Cw);
    const long long n = r / Cw;
    const double sw = rep[(n * 2 * Cw + c) * L + l];
    const double le = rep[(n * 2 * Cw + Cw + c) * L + l];
    wave[i] = (float)(sw * (exp(le + half_log_eps) + eps));
}

// Reference: MovingAverageEnvelope.get_representation (tqdne/representation.py:47-55).  np.convolve(|x|, ones(w)/w,
// mode="same") = mean of |x| over [n - w/2, n + (w-1)/2 ... ] -- for even w: indices n - w/2 .. n + w/2 - 1, zero outside;
// rep = cat([x / (env + eps), log(env + log_eps) - log(log_eps)/2], axis=-2)
__global__ void mavg_forward_kernel(const float* wave, float* rep, int Cw, long long L, long long total, int window,
                                    double log_eps, double half_log_eps, double eps) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long l = i % L;
    const long long r = i / L;
    const int c = (int)(r % Cw);
    const long long n = r / Cw;
    const float* x = wave + r * L;
    // 'same' keeps full[(w-1)/2 : (w-1)/2 + L]:  env[l] = sum_{k = l + (w-1)/2 - (w-1)}^{l + (w-1)/2} |x[k]| / w
    const long long hi = l + (window - 1) / 2, lo = hi - (window - 1);
    double acc = 0.0;
    const double inv = 1.0 / (double)window;
    for (long long kk = (lo < 0 ? 0 : lo); kk <= hi && kk < L; ++kk) acc += fabs((double)x[kk]) * inv;
    const double xv = (double)x[l];
    rep[(n * 2 * Cw + c) * L + l] = (float)(xv / (acc + eps));
    rep[(n * 2 * Cw + Cw + c) * L + l] = (float)(log(acc + log_eps) - half_log_eps);
}

inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

int build_linear(std::vector<Op>& ops, const tq_linear_desc& d) {
    TQ_CHECK(d.M > 0 && d.K > 0 && d.Nout > 0 && d.x && d.W, "linear: bad arguments");
    TQ_CHECK(d.x_rows == 1 || d.x_rows == d.M, "linear: x_rows must be 1 or M");
    TQ_CHECK(!d.add || d.add_rows == 1 || d.add_rows == d.M, "linear: add_rows must be 1 or M");
    auto p = std::make_shared<LinParams>();
    p->M = d.M; p->K = d.K; p->Nout = d.Nout; p->x_rows = d.x_rows; p->x = d.x; p->W = d.W; p->b = d.b;
    p->act_in = d.act_in; p->add = d.add; p->add_rows = d.add_rows; p->y = d.y;
    p->y_act = d.y_act; p->y_act_dtype = d.y_act_dtype;
    TQ_CHECK(!d.y_act || d.y_act_dtype == TQ_F32 || d.y_act_dtype == TQ_BF16, "linear: bad y_act_dtype");
    Op op;
    op.name = "linear_f32";
    op.small = true;
    op.launch = [p](cudaStream_t st) -> int {
        if (p->x_rows != 1 && p->M >= 8 && p->K <= LIN_KMAX) {
            const size_t smem = (size_t)LIN_MB * p->K * sizeof(float);   // <= 32 KB
            TQ_CUDA(launch_pdl(linear_rows_kernel, dim3((p->Nout + 7) / 8), dim3(256), smem, st, *p));
            TQ_CUDA(cudaGetLastError());
            count_launch();
            return 0;
        }
        const long long warps = p->x_rows == 1 ? p->Nout : (long long)p->M * p->Nout;
        TQ_CUDA(launch_pdl(linear_kernel, dim3(blocks_for(warps * 32, 256)), dim3(256), 0, st, *p));
        TQ_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    };
    ops.push_back(std::move(op));
    return 0;
}

int build_fourier(std::vector<Op>& ops, const float* t, const float* W, int M, int half, float* feat) {
    TQ_CHECK(t && W && feat && M > 0 && half > 0, "fourier: bad arguments");
    Op op;
    op.name = "fourier_features";
    op.small = true;
    op.launch = [=](cudaStream_t st) -> int {
        TQ_CUDA(launch_pdl(fourier_kernel, dim3(blocks_for((long long)M * half, 128)), dim3(128), 0, st, t, W, M, half, feat));
        TQ_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    };
    ops.push_back(std::move(op));
    return 0;
}

// conv_resample=False resamplers on channels-last tensors (tqdne/blocks.py:59-64: F.interpolate(scale_factor=2, "nearest");
// blocks.py:104: AvgPool{1,2}d(kernel 2, stride 2)).  One thread per 8 channels of an OUTPUT position; fp32 averaging.
template <typename T>
__device__ __forceinline__ void load_vec8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load_vec8<float>(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load_vec8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        v[2 * k] = __low2float(b2);
        v[2 * k + 1] = __high2float(b2);
    }
}
// mode 0: average pool (Ho = H / 2 unless H == 1, Wo = W / 2), mode 1: nearest x2 (Ho = 2 H unless H == 1, Wo = 2 W)
template <typename T>
__global__ void __launch_bounds__(256) resample2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C,
                                                        int Ho, int Wo, int mode) {
    pdl_launch_dependents();
    pdl_wait();
    const int cv = C >> 3;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)N * Ho * Wo * cv) return;
    const int c0 = (int)(i % cv) * 8;
    long long r = i / cv;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const int n = (int)(r / Ho);
    float o[8];
    if (mode == 1) {
        const int yi = H == 1 ? 0 : yo >> 1, xi = xo >> 1;
        load_vec8<T>(x + (((long long)n * H + yi) * W + xi) * C + c0, o);
    } else {
        const int ny = H == 1 ? 1 : 2;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
        for (int dy = 0; dy < ny; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                float v[8];
                load_vec8<T>(x + (((long long)n * H + (ny * yo + dy)) * W + (2 * xo + dx)) * C + c0, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] += v[j];
            }
        const float inv = 1.f / (float)(2 * ny);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] *= inv;
    }
    store_group8<T>(y + (((long long)n * Ho + yo) * Wo + xo) * C + c0, o);
}

int build_resample2(std::vector<Op>& ops, int dtype, const void* x, void* y, int N, int H, int W, int C, int mode) {
    TQ_CHECK(dtype == TQ_BF16 || dtype == TQ_F32, "resample2: bad dtype");
    TQ_CHECK(x && y && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "resample2: bad arguments (C must be a multiple of 8)");
    TQ_CHECK(mode == 0 || mode == 1, "resample2: mode must be 0 (average pool) or 1 (nearest x2)");
    TQ_CHECK(mode == 1 || (W >= 2 && (H == 1 || H >= 2)), "resample2: average pool needs at least two positions per axis");
    const int Ho = H == 1 ? 1 : (mode == 1 ? 2 * H : H / 2), Wo = mode == 1 ? 2 * W : W / 2;
    const long long threads = (long long)N * Ho * Wo * (C / 8);
    Op op;
    op.name = mode == 1 ? "nearest_upsample2" : "avg_pool2";
    op.small = threads * 8 * 4 < 40e6;
    op.launch = [=](cudaStream_t st) -> int {
        if (dtype == TQ_F32)
            TQ_CUDA(launch_pdl(resample2_kernel<float>, dim3(blocks_for(threads, 256)), dim3(256), 0, st, static_cast<const float*>(x),
                               static_cast<float*>(y), N, H, W, C, Ho, Wo, mode));
        else
            TQ_CUDA(launch_pdl(resample2_kernel<__nv_bfloat16>, dim3(blocks_for(threads, 256)), dim3(256), 0, st,
                               static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), N, H, W, C, Ho, Wo, mode));
        TQ_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    };
    ops.push_back(std::move(op));
    return 0;
}

// mean over the positions of a channels-last fp32 tensor: y[n][c] = mean_p x[n][p][c]   (row stride ld >= C)
// one CTA per (sample, 32-channel slab): 8 position lanes x 32 channels, coalesced 128 B rows, smem tree at the end
__global__ void __launch_bounds__(256) spatial_mean_kernel(const float* __restrict__ x, int P, int C, int ld, float* __restrict__ y) {
    __shared__ float red[8][33];
    pdl_launch_dependents();
    pdl_wait();
    const int n = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), pl = threadIdx.x >> 5;
    float a = 0.f;
    if (c < C) {
        const float* xb = x + (long long)n * P * ld + c;
        for (int p = pl; p < P; p += 8) a += __ldg(xb + (long long)p * ld);
    }
    red[pl][threadIdx.x & 31] = a;
    __syncthreads();
    if (pl == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        y[(long long)n * C + c] = t / (float)P;
    }
}

int build_spatial_mean(std::vector<Op>& ops, const float* x, int N, int P, int C, int ld, float* y) {
    TQ_CHECK(x && y && N > 0 && P > 0 && C > 0 && ld >= C, "spatial_mean: bad arguments");
    Op op;
    op.name = "spatial_mean";
    op.launch = [=](cudaStream_t st) -> int {
        TQ_CUDA(launch_pdl(spatial_mean_kernel, dim3((C + 31) / 32, N), dim3(256), 0, st, x, P, C, ld, y));
        TQ_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    };
    ops.push_back(std::move(op));
    return 0;
}

}  // namespace tq

using namespace tq;

extern "C" int tq_edm_precondition(const double* x, void* xin, int32_t dtype, int64_t NP, int32_t C, int32_t Cpad,
                                   float c_in, float* t_next, float t_value, void* stream) {
    TQ_CHECK(x && xin && NP > 0 && C > 0 && Cpad >= C, "edm_precondition: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (group_ok(xin, Cpad, dtype)) {
        launch_edm_group<0>(const_cast<double*>(x), nullptr, nullptr, nullptr, 0, xin, dtype, NP, C, Cpad, 0.f, 0.f, 1.f, 0.f,
                            c_in, 1, t_next, t_value, st);
        TQ_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    }
    const unsigned g = blocks_for(NP * Cpad, 256);
    if (dtype == TQ_F32) precondition_kernel<float><<<g, 256, 0, st>>>(x, static_cast<float*>(xin), NP, C, Cpad, c_in, t_next, t_value);
    else if (dtype == TQ_BF16)
        precondition_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(x, static_cast<__nv_bfloat16*>(xin), NP, C, Cpad, c_in, t_next,
                                                               t_value);
    else TQ_CHECK(false, "edm_precondition: bad dtype");
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int tq_edm_euler(const double* x, const float* F, int32_t Cf, double* d, double* x1, void* xin, int32_t dtype,
                            int64_t NP, int32_t C, int32_t Cpad, float c_out, float c_skip, float sigma, float dt,
                            float c_in_next, int32_t write_xin, float* t_next, float t_value, void* stream) {
    TQ_CHECK(x && F && x1 && NP > 0 && C > 0 && Cpad >= C && Cf >= C, "edm_euler: bad arguments");
    TQ_CHECK(!write_xin || xin, "edm_euler: xin missing");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (group_ok(write_xin ? xin : nullptr, Cpad, dtype)) {
        launch_edm_group<1>(const_cast<double*>(x), x1, d, F, Cf, xin, dtype, NP, C, Cpad, c_out, c_skip, sigma, dt, c_in_next,
                            write_xin, t_next, t_value, st);
        TQ_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    }
    const unsigned g = blocks_for(NP * Cpad, 256);
    if (dtype == TQ_F32)
        euler_kernel<float><<<g, 256, 0, st>>>(x, F, Cf, d, x1, static_cast<float*>(xin), NP, C, Cpad, c_out, c_skip, sigma,
                                               dt, c_in_next, write_xin, t_next, t_value);
    else if (dtype == TQ_BF16)
        euler_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(x, F, Cf, d, x1, static_cast<__nv_bfloat16*>(xin), NP, C, Cpad, c_out,
                                                       c_skip, sigma, dt, c_in_next, write_xin, t_next, t_value);
    else TQ_CHECK(false, "edm_euler: bad dtype");
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int tq_edm_heun(double* x, const double* x1, const double* d, const float* F, int32_t Cf, void* xin,
                           int32_t dtype, int64_t NP, int32_t C, int32_t Cpad, float c_out, float c_skip, float sigma_next,
                           float dt, float c_in_next, int32_t write_xin, float* t_next, float t_value, void* stream) {
    TQ_CHECK(x && x1 && d && F && NP > 0 && C > 0 && Cpad >= C && Cf >= C, "edm_heun: bad arguments");
    TQ_CHECK(!write_xin || xin, "edm_heun: xin missing");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (group_ok(write_xin ? xin : nullptr, Cpad, dtype)) {
        launch_edm_group<2>(x, const_cast<double*>(x1), const_cast<double*>(d), F, Cf, xin, dtype, NP, C, Cpad, c_out, c_skip,
                            sigma_next, dt, c_in_next, write_xin, t_next, t_value, st);
        TQ_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    }
    const unsigned g = blocks_for(NP * Cpad, 256);
    if (dtype == TQ_F32)
        heun_kernel<float><<<g, 256, 0, st>>>(x, x1, d, F, Cf, static_cast<float*>(xin), NP, C, Cpad, c_out, c_skip,
                                              sigma_next, dt, c_in_next, write_xin, t_next, t_value);
    else if (dtype == TQ_BF16)
        heun_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(x, x1, d, F, Cf, static_cast<__nv_bfloat16*>(xin), NP, C, Cpad, c_out,
                                                      c_skip, sigma_next, dt, c_in_next, write_xin, t_next, t_value);
    else TQ_CHECK(false, "edm_heun: bad dtype");
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int tq_edm_add_noise(double* x, const double* noise, double scale, int64_t n, void* stream) {
    TQ_CHECK(x && noise && n > 0, "edm_add_noise: bad arguments");
    add_noise_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, noise, scale, n);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

namespace {
template <typename TI>
int launch_to_nhwc(const void* src, void* dst, int dtype_out, int N, int C, long long P, int Cpad, cudaStream_t st) {
    dim3 grid(blocks_for(P, 32), (Cpad + 31) / 32, N), block(32, 8);
    if (dtype_out == TQ_BF16)
        nchw_to_nhwc_kernel<TI, __nv_bfloat16><<<grid, block, 0, st>>>(static_cast<const TI*>(src), static_cast<__nv_bfloat16*>(dst), C, P, Cpad);
    else if (dtype_out == TQ_F32)
        nchw_to_nhwc_kernel<TI, float><<<grid, block, 0, st>>>(static_cast<const TI*>(src), static_cast<float*>(dst), C, P, Cpad);
    else if (dtype_out == TQ_F64)
        nchw_to_nhwc_kernel<TI, double><<<grid, block, 0, st>>>(static_cast<const TI*>(src), static_cast<double*>(dst), C, P, Cpad);
    else TQ_CHECK(false, "nchw_to_nhwc: bad dtype_out");
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}
template <typename TI>
int launch_to_nchw(const void* src, int Cld, void* dst, int dtype_out, int N, int C, long long P, cudaStream_t st) {
    dim3 grid(blocks_for(P, 32), (C + 31) / 32, N), block(32, 8);
    if (dtype_out == TQ_BF16)
        nhwc_to_nchw_kernel<TI, __nv_bfloat16><<<grid, block, 0, st>>>(static_cast<const TI*>(src), Cld, static_cast<__nv_bfloat16*>(dst), C, P);
    else if (dtype_out == TQ_F32)
        nhwc_to_nchw_kernel<TI, float><<<grid, block, 0, st>>>(static_cast<const TI*>(src), Cld, static_cast<float*>(dst), C, P);
    else if (dtype_out == TQ_F64)
        nhwc_to_nchw_kernel<TI, double><<<grid, block, 0, st>>>(static_cast<const TI*>(src), Cld, static_cast<double*>(dst), C, P);
    else TQ_CHECK(false, "nhwc_to_nchw: bad dtype_out");
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}
}  // namespace

extern "C" int tq_nchw_to_nhwc(const void* src, int32_t dtype_in, void* dst, int32_t dtype_out, int32_t N, int32_t C,
                               int64_t P, int32_t Cpad, void* stream) {
    TQ_CHECK(src && dst && N > 0 && C > 0 && P > 0 && Cpad >= C, "nchw_to_nhwc: bad arguments");
    TQ_CHECK(N <= 65535, "nchw_to_nhwc: batch too large for one launch");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype_in == TQ_F64) return launch_to_nhwc<double>(src, dst, dtype_out, N, C, P, Cpad, st);
    if (dtype_in == TQ_F32) return launch_to_nhwc<float>(src, dst, dtype_out, N, C, P, Cpad, st);
    if (dtype_in == TQ_BF16) return launch_to_nhwc<__nv_bfloat16>(src, dst, dtype_out, N, C, P, Cpad, st);
    TQ_CHECK(false, "nchw_to_nhwc: bad dtype_in");
}

extern "C" int tq_nhwc_to_nchw(const void* src, int32_t dtype_in, int32_t Cld, void* dst, int32_t dtype_out, int32_t N,
                               int32_t C, int64_t P, void* stream) {
    TQ_CHECK(src && dst && N > 0 && C > 0 && P > 0 && Cld >= C, "nhwc_to_nchw: bad arguments");
    TQ_CHECK(N <= 65535, "nhwc_to_nchw: batch too large for one launch");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype_in == TQ_F64) return launch_to_nchw<double>(src, Cld, dst, dtype_out, N, C, P, st);
    if (dtype_in == TQ_F32) return launch_to_nchw<float>(src, Cld, dst, dtype_out, N, C, P, st);
    if (dtype_in == TQ_BF16) return launch_to_nchw<__nv_bfloat16>(src, Cld, dst, dtype_out, N, C, P, st);
    TQ_CHECK(false, "nhwc_to_nchw: bad dtype_in");
}

extern "C" int tq_mavg_envelope_inverse(const float* rep, float* wave, int32_t N, int32_t Cw, int64_t L, double log_eps,
                                        double eps, void* stream) {
    TQ_CHECK(rep && wave && N > 0 && Cw > 0 && L > 0, "mavg_envelope_inverse: bad arguments");
    const long long total = (long long)N * Cw * L;
    mavg_inverse_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rep, wave, Cw, L, total,
                                                                                              log(log_eps) / 2.0, eps);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int tq_mavg_envelope_forward(const float* wave, float* rep, int32_t N, int32_t Cw, int64_t L, int32_t window,
                                        double log_eps, double eps, void* stream) {
    TQ_CHECK(wave && rep && N > 0 && Cw > 0 && L > 0 && window > 0, "mavg_envelope_forward: bad arguments");
    const long long total = (long long)N * Cw * L;
    mavg_forward_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        wave, rep, Cw, L, total, window, log_eps, log(log_eps) / 2.0, eps);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}
