// tq_train.cu -- the small kernels of the training step: resampling helpers of the strided /
// upsampled convolution gradients, dense-layer backward of the embedding MLPs, EDM noising + loss, dropout, Adam + EMA.
// Reference: LightningEDM.step / configure_optimizers (tqdne/edm.py:115-134,240-251), EMA (tqdne/ema.py:24-28),
// Downsample / Upsample (tqdne/blocks.py:29-108), ResBlock dropout (tqdne/unet.py:100-108).  All HBM- or latency-bound.
#include <cuda_bf16.h>

#include <algorithm>

#include "tq_common.h"

namespace tq {
namespace {

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        v[2 * k] = __low2float(b2);
        v[2 * k + 1] = __high2float(b2);
    }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        w[k] = *reinterpret_cast<const uint32_t*>(&b2);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

// rows of a channels-last bf16 tensor [N][L][C], 8 channels per thread; index space = destination vectors
//   mode 0  zero_stuff   dst[n][2j] = src[n][j], dst[n][2j+1] = 0           (gradient of "take the even positions")
//   mode 1  upsample     dst[n][l]  = src[n][l / 2]                         (F.interpolate(scale_factor=2, "nearest"))
//   mode 2  pair_sum     dst[n][j]  = src[n][2j] + src[n][2j+1]             (its gradient)
//   mode 3  mul          dst[n][l]  = src[n][l] * aux[n][l]                 (dropout mask, forward and backward)
//   mode 4  add          dst[n][l]  = src[n][l] + aux[n][l]                 (two gradient paths of one tensor)
__global__ void __launch_bounds__(256) rows_op_kernel(const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* __restrict__ aux,
                                                      __nv_bfloat16* __restrict__ dst, int mode, long long N, long long Ld, int C) {
    const int cv = C >> 3;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= N * Ld * cv) return;
    const int vi = (int)(i % cv);
    const long long r = i / cv, n = r / Ld, l = r % Ld;
    float v[8] = {};
    if (mode == 0) {
        if ((l & 1) == 0) ld8(src + ((n * (Ld / 2) + l / 2) * C + vi * 8), v);
    } else if (mode == 1) {
        ld8(src + ((n * (Ld / 2) + l / 2) * C + vi * 8), v);
    } else if (mode == 2) {
        float a[8], b[8];
        ld8(src + ((n * (2 * Ld) + 2 * l) * C + vi * 8), a);
        ld8(src + ((n * (2 * Ld) + 2 * l + 1) * C + vi * 8), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = a[j] + b[j];
    } else {
        float a[8], b[8];
        ld8(src + (r * C + vi * 8), a);
        ld8(aux + (r * C + vi * 8), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = mode == 3 ? a[j] * b[j] : a[j] + b[j];
    }
    st8(dst + (r * C + vi * 8), v);
}

__device__ __forceinline__ float silu_d(float x) {
    const float s = 1.f / (1.f + expf(-x));
    return s * (1.f + x * (1.f - s));
}
__device__ __forceinline__ float silu_v(float x) { return x / (1.f + expf(-x)); }

// dense layer v = act(x) W^T + b, fp32 (time / conditioning MLPs, emb_layers): dW[j][k] += sum_m dy[m][j] act(x[m][k])
// One block = LBW_J output rows j of dW x all K columns: act(x) [M x K] and dy[:, j-tile] are staged in shared memory once,
// thread = column k (strided), so the M-long sums read shared memory only (the one-thread-per-element version it replaces
// walked M dependent global loads per element: 153 us for the 4032 x 256 emb_layers gradient at M = 64).
constexpr int LBW_J = 16;
__global__ void __launch_bounds__(256) linear_bwd_dw_kernel(const float* __restrict__ dy, const float* __restrict__ x, int act_in,
                                                            float* __restrict__ dW, float* __restrict__ db, int M, int K, int Nout) {
    extern __shared__ float sm_lbw[];
    float* xs = sm_lbw;              // [M][K]
    float* dys = xs + (((size_t)M * K + 3) & ~(size_t)3);  // [M][LBW_J], 16 B aligned
    const int j0 = blockIdx.x * LBW_J, nj = min(LBW_J, Nout - j0);
    for (int i = threadIdx.x; i < M * K; i += 256) {
        const float xv = __ldg(x + i);
        xs[i] = act_in ? silu_v(xv) : xv;
    }
    for (int i = threadIdx.x; i < M * LBW_J; i += 256) {
        const int m = i / LBW_J, jj = i % LBW_J;
        dys[i] = jj < nj ? __ldg(dy + (long long)m * Nout + j0 + jj) : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += 256) {
        float a[LBW_J];
#pragma unroll
        for (int jj = 0; jj < LBW_J; ++jj) a[jj] = 0.f;
        for (int m = 0; m < M; ++m) {
            const float xv = xs[m * K + k];
            const float4* d4 = reinterpret_cast<const float4*>(dys + m * LBW_J);   // broadcast reads
#pragma unroll
            for (int q = 0; q < LBW_J / 4; ++q) {
                const float4 d = d4[q];
                a[4 * q] = fmaf(d.x, xv, a[4 * q]);
                a[4 * q + 1] = fmaf(d.y, xv, a[4 * q + 1]);
                a[4 * q + 2] = fmaf(d.z, xv, a[4 * q + 2]);
                a[4 * q + 3] = fmaf(d.w, xv, a[4 * q + 3]);
            }
        }
#pragma unroll
        for (int jj = 0; jj < LBW_J; ++jj)
            if (jj < nj) dW[(long long)(j0 + jj) * K + k] += a[jj];
    }
    if (db && threadIdx.x < nj) {
        float bsum = 0.f;
        for (int m = 0; m < M; ++m) bsum += dys[m * LBW_J + threadIdx.x];
        db[j0 + threadIdx.x] += bsum;
    }
}
// dx[m][k] = act'(x[m][k]) * sum_j dy[m][j] W[j][k]: the j range is split over blocks (the emb_layers GEMM has ~4000 rows
// against M*K = 16 K outputs) and accumulated with fp32 atomics into the zeroed dx; a second pass applies act'
constexpr int LIN_JB = 64;
__global__ void __launch_bounds__(256) linear_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ W,
                                                            float* __restrict__ dx, int M, int K, int Nout) {
    const int m = blockIdx.y, j0 = blockIdx.x * LIN_JB, j1 = min(Nout, j0 + LIN_JB);
    __shared__ float dys[LIN_JB];
    for (int j = threadIdx.x; j < j1 - j0; j += 256) dys[j] = dy[(long long)m * Nout + j0 + j];
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += 256) {
        float a = 0.f;
        for (int j = j0; j < j1; ++j) a = fmaf(dys[j - j0], __ldg(W + (long long)j * K + k), a);
        atomicAdd(dx + (long long)m * K + k, a);
    }
}
__global__ void __launch_bounds__(256) silu_grad_scale_kernel(float* __restrict__ dx, const float* __restrict__ x, long long n) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dx[i] *= silu_d(x[i]);
}

// EDM noising (edm.py:125-128): xn = y + sigma_n * noise; network input = bf16(c_in(sigma_n) * xn), channels padded
__global__ void __launch_bounds__(256) edm_noise_kernel(const float* __restrict__ y, const float* __restrict__ noise,
                                                        const float* __restrict__ sigma, float* __restrict__ xn,
                                                        __nv_bfloat16* __restrict__ xin, long long P, int C, int Cpad, long long total,
                                                        float sigma_data) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;   // over [N][P][Cpad]
    if (i >= total) return;
    const int c = (int)(i % Cpad);
    const long long r = i / Cpad, n = r / P;
    float v = 0.f;
    if (c < C) {
        const float s = sigma[n];
        const float x = y[r * C + c] + s * noise[r * C + c];
        xn[r * C + c] = x;
        v = x * rsqrtf(s * s + sigma_data * sigma_data);
    }
    xin[i] = __float2bfloat16_rn(v);
}
// EDM loss (edm.py:105-113,129-134): pred = c_out F + c_skip xn; loss = mean(w (pred - y)^2), w = (s^2 + sd^2) / (s sd)^2;
// dF = 2 w (pred - y) c_out / count, written as bf16 with the channels padded (the output conv's dY)
__global__ void __launch_bounds__(256) edm_loss_kernel(const float* __restrict__ F, int Cf, const float* __restrict__ xn,
                                                       const float* __restrict__ y, const float* __restrict__ sigma,
                                                       __nv_bfloat16* __restrict__ dF, float* __restrict__ loss, long long P, int C,
                                                       int Cpad, long long total, float sigma_data, float inv_count) {
    float contrib = 0.f;
    // thread = 8 consecutive padded channels of a position (one 16 B store; only the first ceil(C / 8) groups of a row hold
    // signal), grid-stride over a bounded grid: the loss then takes ~1 200 atomics on its one address -- one per 256 elements
    // (65 024 serialised atomics) was 100 of the old kernel's 140 us
    const int gpr = Cpad >> 3;
    const long long groups = total >> 3;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < groups; i += (long long)gridDim.x * 256) {
        const int c0 = (int)(i % gpr) * 8;
        const long long r = i / gpr;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (c0 < C) {
            const float s = sigma[r / P], sd = sigma_data;
            const float den = s * s + sd * sd;
            const float c_out = s * sd * rsqrtf(den), c_skip = sd * sd / den, wt = den / (s * sd * s * sd);
            float g[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                g[j] = 0.f;
                if (c0 + j < C) {
                    const float diff = c_out * F[r * Cf + c0 + j] + c_skip * xn[r * C + c0 + j] - y[r * C + c0 + j];
                    contrib += wt * diff * diff * inv_count;
                    g[j] = 2.f * wt * diff * c_out * inv_count;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __nv_bfloat162 b2 = __floats2bfloat162_rn(g[2 * k], g[2 * k + 1]);
                w[k] = *reinterpret_cast<const uint32_t*>(&b2);
            }
        }
        *reinterpret_cast<uint4*>(dF + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = contrib;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k];
        atomicAdd(loss, t);
    }
}

// dropout: element i is dropped when hash(seed, i) < p (counter-based splitmix64: the backward pass regenerates the same
// decisions from the same seed, no mask tensor is stored); kept elements are scaled by 1 / (1 - p)   (nn.Dropout)
__global__ void __launch_bounds__(256) dropout_mask_kernel(__nv_bfloat16* __restrict__ mask, long long n, unsigned long long seed,
                                                           float p) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    mask[i] = __float2bfloat16_rn(dropout_scale(seed, i, p, 1.f / (1.f - p)));
}
// dst = src * dropout_scale: 8 elements per thread
__global__ void __launch_bounds__(256) dropout_apply_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                            long long nvec, unsigned long long seed, float p) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= nvec) return;
    float v[8];
    ld8(src + i * 8, v);
    const float ks = 1.f / (1.f - p);
    float mk[8];
    dropout_scale8(seed, i * 8, p, ks, mk);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= mk[j];
    st8(dst + i * 8, v);
}

// bf16 operand copies of one convolution from its fp32 master [Op][k][Ip] (engine layout):
//   fwd[co][t][ci]      = master[co][t][ci]                       (the forward igemm's [cout_pad, k * cin_pad] matrix)
//   bwd[ci][t][co]      = master[co][k - 1 - t][ci_off + ci]      (input gradient: taps flipped, in / out transposed)
__global__ void __launch_bounds__(256) repack_fwd_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, long long n) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16_rn(w[i]);
}
__global__ void __launch_bounds__(256) repack_bwd_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Op, int k,
                                                         int Ip, int ci_off, int Cs) {
    // 32 x 32 (co, ci) tile transpose through shared memory per tap: coalesced on both sides
    __shared__ float tile[32][33];
    const int t = blockIdx.z;
    const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows per pass
    for (int r = ty; r < 32; r += 8) {
        const int co = co0 + r, ci = ci0 + tx;
        tile[r][tx] = (co < Op && ci < Cs) ? w[((long long)co * k + (k - 1 - t)) * Ip + ci_off + ci] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int ci = ci0 + r, co = co0 + tx;
        if (ci < Cs && co < Op) dst[((long long)ci * k + t) * Op + co] = __float2bfloat16_rn(tile[tx][r]);
    }
}

// Every operand copy of a model in one launch: block b finds its job by binary search over the jobs' first blocks
// (tq_repack_batch_prepare), then does what the two kernels above do -- a forward job casts 2048 consecutive elements per
// block (8 per thread), a backward job transposes one 32 x 32 (co, ci) tile of one tap through shared memory.
__global__ void __launch_bounds__(256) repack_batch_kernel(const tq_repack_job* __restrict__ jobs, int n_jobs) {
    __shared__ float tile[32][33];
    int lo = 0, hi = n_jobs - 1;
    const int b = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].block0 <= b) lo = mid;
        else hi = mid - 1;
    }
    const tq_repack_job j = jobs[lo];
    const int local = b - j.block0;
    if (j.fwd != nullptr) {
        const long long n = (long long)j.Op * j.k * j.Ip;   // a multiple of 8 (Ip is padded to 64)
        const long long i = ((long long)local * 256 + threadIdx.x) * 8;
        if (i < n) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(j.master + i));
            const float4 c = __ldg(reinterpret_cast<const float4*>(j.master + i) + 1);
            const __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
            const __nv_bfloat162 p2 = __floats2bfloat162_rn(c.x, c.y), p3 = __floats2bfloat162_rn(c.z, c.w);
            uint4 o;
            o.x = *reinterpret_cast<const unsigned int*>(&p0); o.y = *reinterpret_cast<const unsigned int*>(&p1);
            o.z = *reinterpret_cast<const unsigned int*>(&p2); o.w = *reinterpret_cast<const unsigned int*>(&p3);
            *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(j.fwd) + i) = o;
        }
        return;
    }
    const int tiles_ci = (j.Cs + 31) / 32, tiles_co = (j.Op + 31) / 32;
    const int t = local / (tiles_ci * tiles_co);
    const int r0 = local % (tiles_ci * tiles_co);
    const int co0 = (r0 / tiles_ci) * 32, ci0 = (r0 % tiles_ci) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(j.bwd);
    for (int r = ty; r < 32; r += 8) {
        const int co = co0 + r, ci = ci0 + tx;
        tile[r][tx] = (co < j.Op && ci < j.Cs) ? __ldg(j.master + ((long long)co * j.k + (j.k - 1 - t)) * j.Ip + j.ci_off + ci) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int ci = ci0 + r, co = co0 + tx;
        if (ci < j.Cs && co < j.Op) dst[((long long)ci * j.k + t) * j.Op + co] = __float2bfloat16_rn(tile[tx][r]);
    }
}

// Adam (torch.optim.Adam defaults: no weight decay, no amsgrad) + EMA lerp (ema.py:24-28), one pass over flat fp32 arrays
__global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ ema, long long n, float lr, float b1,
                                                       float b2, float eps, float bc1, float bc2, float ema_w, float grad_scale) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float g = grad[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * g;
    const float vi = b2 * v[i] + (1.f - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    // torch: param -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
    const float pnew = param[i] - (lr / bc1) * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
    param[i] = pnew;
    if (ema) ema[i] += ema_w * (pnew - ema[i]);
}

inline unsigned grid_for(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace
}  // namespace tq

using namespace tq;

extern "C" {

int tq_rows_op(const void* src, const void* aux, void* dst, int32_t mode, int64_t N, int64_t L_dst, int32_t C, void* stream) {
    TQ_CHECK(src && dst && mode >= 0 && mode <= 4 && N > 0 && L_dst > 0 && C > 0 && C % 8 == 0, "rows_op: bad arguments");
    TQ_CHECK(mode < 3 || aux, "rows_op: mul / add need aux");
    TQ_CHECK(mode >= 2 || L_dst % 2 == 0, "rows_op: destination length must be even");
    rows_op_kernel<<<grid_for(N * L_dst * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(src), static_cast<const __nv_bfloat16*>(aux), static_cast<__nv_bfloat16*>(dst), mode, N,
        L_dst, C);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int tq_linear_backward(const float* dy, const float* x, const float* W, int32_t act_in, float* dx, float* dW, float* db, int32_t M,
                       int32_t K, int32_t Nout, void* stream) {
    TQ_CHECK(dy && x && W && M > 0 && K > 0 && Nout > 0, "linear_backward: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dW) {
        const size_t smem = ((((size_t)M * K + 3) & ~(size_t)3) + (size_t)M * LBW_J) * sizeof(float);
        TQ_CHECK(smem <= 200 * 1024, "linear_backward: M * K too large for the shared-memory staging (%zu bytes)", smem);
        static PerDeviceMax attr;
        if (attr.raise(smem)) TQ_CUDA(cudaFuncSetAttribute(linear_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        linear_bwd_dw_kernel<<<(Nout + LBW_J - 1) / LBW_J, 256, smem, st>>>(dy, x, act_in, dW, db, M, K, Nout);
        TQ_CUDA(cudaGetLastError());
        count_launch();
    }
    if (dx) {
        TQ_CUDA(cudaMemsetAsync(dx, 0, (size_t)M * K * sizeof(float), st));
        linear_bwd_dx_kernel<<<dim3((Nout + LIN_JB - 1) / LIN_JB, M), 256, 0, st>>>(dy, W, dx, M, K, Nout);
        TQ_CUDA(cudaGetLastError());
        if (act_in) silu_grad_scale_kernel<<<grid_for((long long)M * K), 256, 0, st>>>(dx, x, (long long)M * K);
        TQ_CUDA(cudaGetLastError());
        count_launch(act_in ? 2 : 1);
    }
    return 0;
}

int tq_edm_noise(const float* y, const float* noise, const float* sigma, float* xn, void* xin, int64_t N, int64_t P, int32_t C,
                 int32_t Cpad, float sigma_data, void* stream) {
    TQ_CHECK(y && noise && sigma && xn && xin && N > 0 && P > 0 && C > 0 && Cpad >= C, "edm_noise: bad arguments");
    const long long total = N * P * Cpad;
    edm_noise_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, noise, sigma, xn,
                                                                                     static_cast<__nv_bfloat16*>(xin), P, C, Cpad,
                                                                                     total, sigma_data);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int tq_edm_loss(const float* F, int32_t Cf, const float* xn, const float* y, const float* sigma, void* dF, float* loss, int64_t N,
                int64_t P, int32_t C, int32_t Cpad, float sigma_data, void* stream) {
    TQ_CHECK(F && xn && y && sigma && dF && loss && N > 0 && P > 0 && C > 0 && Cpad >= C && Cf >= C, "edm_loss: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TQ_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    const long long total = N * P * Cpad;
    TQ_CHECK(Cpad % 8 == 0 && (reinterpret_cast<uintptr_t>(dF) & 15) == 0, "edm_loss: dF must be 16 B aligned with Cpad %% 8 == 0");
    const unsigned loss_grid = std::min<unsigned>(grid_for(total / 8), 8u * (unsigned)device_sm_count());
    edm_loss_kernel<<<loss_grid, 256, 0, st>>>(F, Cf, xn, y, sigma, static_cast<__nv_bfloat16*>(dF), loss, P, C, Cpad, total,
                                                    sigma_data, 1.f / (float)(N * P * C));
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int tq_dropout_mask(void* mask, int64_t n, uint64_t seed, float p, void* stream) {
    TQ_CHECK(mask && n > 0 && p >= 0.f && p < 1.f, "dropout_mask: bad arguments");
    dropout_mask_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<__nv_bfloat16*>(mask), n, seed, p);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int tq_dropout_apply(const void* src, void* dst, int64_t n, uint64_t seed, float p, void* stream) {
    TQ_CHECK(src && dst && n > 0 && n % 8 == 0 && p >= 0.f && p < 1.f, "dropout_apply: bad arguments");
    dropout_apply_kernel<<<grid_for(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), n / 8, seed, p);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int tq_repack_conv_weights(const float* master, void* fwd, void* bwd, int32_t Op, int32_t k, int32_t Ip, int32_t ci_off,
                           int32_t Cs, void* stream) {
    TQ_CHECK(master && Op > 0 && k > 0 && Ip > 0 && ci_off >= 0 && (bwd == nullptr || ci_off + Cs <= Ip), "repack_conv_weights: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (fwd) {
        const long long n = (long long)Op * k * Ip;
        repack_fwd_kernel<<<grid_for(n), 256, 0, st>>>(master, static_cast<__nv_bfloat16*>(fwd), n);
        TQ_CUDA(cudaGetLastError());
        count_launch();
    }
    if (bwd) {
        repack_bwd_kernel<<<dim3((Cs + 31) / 32, (Op + 31) / 32, k), 256, 0, st>>>(master, static_cast<__nv_bfloat16*>(bwd), Op, k, Ip,
                                                                                    ci_off, Cs);
        TQ_CUDA(cudaGetLastError());
        count_launch();
    }
    return 0;
}

int64_t tq_repack_batch_prepare(tq_repack_job* jobs, int32_t n_jobs) {
    if (jobs == nullptr || n_jobs <= 0) { set_error("repack_batch_prepare: no jobs"); return -1; }
    long long total = 0;
    for (int i = 0; i < n_jobs; ++i) {
        tq_repack_job& j = jobs[i];
        const bool one = (j.fwd != nullptr) != (j.bwd != nullptr);
        if (!j.master || !one || j.Op <= 0 || j.k <= 0 || j.Ip <= 0 || j.Ip % 8 != 0 || j.ci_off < 0 ||
            (j.bwd != nullptr && (j.Cs <= 0 || j.ci_off + j.Cs > j.Ip))) {
            set_error("repack_batch_prepare: bad job %d", i);
            return -1;
        }
        long long nb;
        if (j.fwd) nb = ((long long)j.Op * j.k * j.Ip + 2047) / 2048;
        else nb = (long long)((j.Cs + 31) / 32) * ((j.Op + 31) / 32) * j.k;
        if (total + nb > 0x7fffffffLL) { set_error("repack_batch_prepare: too many blocks"); return -1; }
        j.block0 = (int32_t)total;
        j.nblocks = (int32_t)nb;
        total += nb;
    }
    return total;
}

int tq_repack_batch_run(const tq_repack_job* jobs_device, int32_t n_jobs, int64_t total_blocks, void* stream) {
    TQ_CHECK(jobs_device && n_jobs > 0 && total_blocks > 0 && total_blocks <= 0x7fffffffLL, "repack_batch_run: bad arguments");
    repack_batch_kernel<<<(unsigned)total_blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(jobs_device, n_jobs);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int tq_adam_ema_step(float* param, const float* grad, float* m, float* v, float* ema, int64_t n, float lr, float beta1, float beta2,
                     float eps, int64_t step, float ema_decay, float grad_scale, void* stream) {
    TQ_CHECK(param && grad && m && v && n > 0 && step >= 1, "adam_ema_step: bad arguments");
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    adam_ema_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, m, v, ema, n, lr, beta1, beta2, eps, bc1,
                                                                                 bc2, 1.f - ema_decay, grad_scale);
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // extern "C"
