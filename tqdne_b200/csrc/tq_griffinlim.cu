// tq_griffinlim.cu -- batched log-spectrogram inverse: un-normalise, exp, append the zero Nyquist row,
// then fast Griffin-Lim (momentum 0.99) phase reconstruction, all iterations in ONE launch.
//
// Reference: LogSpectrogram.invert_representation / invert_spectrogram (tqdne/representation.py:152-175)
// calling librosa 0.11 griffinlim(S, n_iter=128, hop_length=32, n_fft=256, random_state=0)
// (representation.py:103-108).  librosa is a third-party dependency absent from /root/reference; the
// algorithm restated here (and in oracle/griffinlim_ref.py) is its published one:
//     angles = S * exp(i*phase0)
//     repeat n_iter: y = istft(angles); R = stft(y); a = R - m/(1+m) * R_prev;
//                    angles = S * a / (|a| + tiny);  R_prev = R
//     return istft(angles)
// stft/istft: n_fft = win = 256, hop 32, periodic Hann, center=True with zero padding, irfft * window
// overlap-add divided by the window sum-of-squares where it exceeds tiny, 128 samples trimmed per side.
//
// One CTA (16 warps) owns one (sample, component) item for the whole reconstruction.  The padded signal
// (4320 samples) and the per-warp FFT buffers live in shared memory; a frame's real 256-point transform is
// a 128-point complex radix-2 FFT held by one warp plus the usual even/odd split.  The overlap-add runs in
// 8 barrier-separated groups of non-overlapping frames (t mod 8), so the summation order is fixed and no
// atomics are needed.  The complex spectra (angles, R_prev) and S stream through L2.
#include <math_constants.h>

#include <cstdlib>

#include "tq_common.h"

namespace tq {
namespace {

constexpr int NFFT = 256;
constexpr int HOP = 32;
constexpr int NBIN = NFFT / 2 + 1;  // 129
constexpr int GL_THREADS = 512;
constexpr int GL_WARPS = GL_THREADS / 32;

template <typename R>
struct alignas(2 * sizeof(R)) Cx {
    R x, y;
};

struct GlParams {
    const float* rep;      // [items][128][frames]
    const double* phase0;  // [129][frames][2]: unit phasors (cos, sin) of the initial phase
    void* wave;            // [items][hop*(frames-1)]  (float or double = R)
    void* ws;
    int items, frames, n_iter;
    double log_clip, log_max, mom;  // mom = momentum / (1 + momentum)
};

template <typename R> __device__ __forceinline__ R r_tiny();
template <> __device__ __forceinline__ float r_tiny<float>() { return 1.17549435e-38f; }
template <> __device__ __forceinline__ double r_tiny<double>() { return 2.2250738585072014e-308; }
__device__ __forceinline__ float r_hypot(float a, float b) { return hypotf(a, b); }
__device__ __forceinline__ double r_hypot(double a, double b) { return hypot(a, b); }

__device__ __forceinline__ int bitrev7(int v) { return (int)(__brev((unsigned)v) >> 25); }

// in-place 128-point complex FFT, decimation in time: input bit-reversed, output natural order
template <typename R, bool INV>
__device__ __forceinline__ void fft128(Cx<R>* b, const Cx<R>* tw, int lane) {
#pragma unroll
    for (int s = 1; s <= 7; ++s) {
        const int half = 1 << (s - 1);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int j = lane + 32 * i;
            const int pos = j & (half - 1);
            const int i0 = ((j >> (s - 1)) << s) + pos;
            const int i1 = i0 + half;
            Cx<R> w = tw[pos << (7 - s)];
            if (INV) w.y = -w.y;
            const Cx<R> v = b[i1];
            const Cx<R> u = b[i0];
            const R tr = w.x * v.x - w.y * v.y;
            const R ti = w.x * v.y + w.y * v.x;
            b[i0] = Cx<R>{u.x + tr, u.y + ti};
            b[i1] = Cx<R>{u.x - tr, u.y - ti};
        }
        __syncwarp();
    }
}

template <typename R>
__global__ void __launch_bounds__(GL_THREADS) griffinlim_kernel(const GlParams p) {
    extern __shared__ __align__(16) unsigned char gl_smem[];
    const int frames = p.frames;
    const int len = NFFT + HOP * (frames - 1);  // padded signal length
    R* ypad = reinterpret_cast<R*>(gl_smem);
    R* win = ypad + ((len + 3) & ~3);
    Cx<R>* tw128 = reinterpret_cast<Cx<R>*>(win + NFFT);
    Cx<R>* tw256 = tw128 + 64;
    Cx<R>* bufs = tw256 + 132;

    const int item = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    Cx<R>* buf = bufs + warp * 128;

    const size_t per_item = (size_t)frames * NBIN;
    R* S = static_cast<R*>(p.ws) + (size_t)item * per_item * 5;
    Cx<R>* A = reinterpret_cast<Cx<R>*>(S + per_item);
    Cx<R>* Tp = A + per_item;

    for (int i = tid; i < NFFT; i += GL_THREADS) win[i] = (R)(0.5 - 0.5 * cospi(2.0 * i / NFFT));
    for (int i = tid; i < 64; i += GL_THREADS) {
        double s, c;
        sincospi(-2.0 * i / 128.0, &s, &c);
        tw128[i] = Cx<R>{(R)c, (R)s};
    }
    for (int i = tid; i <= 128; i += GL_THREADS) {
        double s, c;
        sincospi(-2.0 * i / 256.0, &s, &c);
        tw256[i] = Cx<R>{(R)c, (R)s};
    }
    // S = exp(((rep + 1) / 2) * (log_max - log_clip) + log_clip), Nyquist row = 0; angles = S * exp(i phase0)
    const float* rep = p.rep + (size_t)item * 128 * frames;
    for (int idx = tid; idx < frames * NBIN; idx += GL_THREADS) {
        const int f = idx / frames, t = idx % frames;  // coalesced read of rep[f][t]
        R sv = 0;
        if (f < 128) {
            const float nls = (rep[(size_t)f * frames + t] + 1.f) / 2.f;  // float32 like the reference input
            sv = (R)exp((double)nls * (p.log_max - p.log_clip) + p.log_clip);
        }
        const double2 u = reinterpret_cast<const double2*>(p.phase0)[(size_t)f * frames + t];   // (cos, sin) of the phase
        S[(size_t)t * NBIN + f] = sv;
        A[(size_t)t * NBIN + f] = Cx<R>{(R)((double)sv * u.x), (R)((double)sv * u.y)};
    }
    __syncthreads();

    const R mom = (R)p.mom;
    const R inv_n = (R)(1.0 / 128.0);
    for (int it = 0; it <= p.n_iter; ++it) {
        for (int i = tid; i < len; i += GL_THREADS) ypad[i] = 0;
        __syncthreads();
        // ---------------- istft: irfft each frame, window, overlap-add (8 groups of disjoint frames)
        for (int g = 0; g < 8; ++g) {
            for (int t = g + 8 * warp; t < frames; t += 8 * GL_WARPS) {
                const Cx<R>* At = A + (size_t)t * NBIN;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = lane + 32 * i;
                    Cx<R> xk = At[k];
                    Cx<R> xn = At[128 - k];
                    if (k == 0) {  // c2r transforms ignore the imaginary parts of DC and Nyquist
                        xk.y = 0;
                        xn.y = 0;
                    }
                    const R er = (R)0.5 * (xk.x + xn.x), ei = (R)0.5 * (xk.y - xn.y);
                    const R dr = (R)0.5 * (xk.x - xn.x), di = (R)0.5 * (xk.y + xn.y);
                    const R c = tw256[k].x, s = -tw256[k].y;  // e^{+2 pi i k / 256}
                    const R orr = dr * c - di * s, oi = dr * s + di * c;
                    buf[bitrev7(k)] = Cx<R>{er - oi, ei + orr};
                }
                __syncwarp();
                fft128<R, true>(buf, tw128, lane);
                R* yt = ypad + HOP * t;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = lane + 32 * i;
                    const Cx<R> z = buf[n];
                    yt[2 * n] += win[2 * n] * (z.x * inv_n);
                    yt[2 * n + 1] += win[2 * n + 1] * (z.y * inv_n);
                }
                __syncwarp();
            }
            __syncthreads();
        }
        // ---------------- divide by the window sum of squares; samples outside [128, len-128) are the
        // centre padding: discarded by istft's trim and zero for the next stft
        for (int n = tid; n < len; n += GL_THREADS) {
            R v = 0;
            if (n >= NFFT / 2 && n < len - NFFT / 2) {
                R wss = 0;
                const int r = n & (HOP - 1);
#pragma unroll
                for (int j = 0; j < NFFT / HOP; ++j) {
                    const int o = r + HOP * j;       // offset inside a frame
                    const int t = (n - o) / HOP;     // exact: n - o is a multiple of HOP
                    if (n - o >= 0 && t < frames) wss += win[o] * win[o];
                }
                v = ypad[n];
                if (wss > r_tiny<R>()) v /= wss;
            }
            ypad[n] = v;
        }
        __syncthreads();
        if (it == p.n_iter) break;
        // ---------------- stft + fast Griffin-Lim phase update
        for (int t = warp; t < frames; t += GL_WARPS) {
            const R* yt = ypad + HOP * t;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int n = lane + 32 * i;
                buf[bitrev7(n)] = Cx<R>{yt[2 * n] * win[2 * n], yt[2 * n + 1] * win[2 * n + 1]};
            }
            __syncwarp();
            fft128<R, false>(buf, tw128, lane);
            const size_t row = (size_t)t * NBIN;
            for (int k = lane; k <= 128; k += 32) {
                const Cx<R> zk = buf[k & 127];
                const Cx<R> zn = buf[(128 - k) & 127];
                const R er = (R)0.5 * (zk.x + zn.x), ei = (R)0.5 * (zk.y - zn.y);
                const R dr = (R)0.5 * (zk.x - zn.x), di = (R)0.5 * (zk.y + zn.y);
                const R c = tw256[k].x, s = tw256[k].y;  // e^{-2 pi i k / 256}
                // X = Xe + tw * Xo,  Xo = -i * (dr + i di) = di - i dr
                const R xr = er + (di * c + dr * s);
                const R xi = ei + (di * s - dr * c);
                R ar = xr, ai = xi;
                if (it > 0) {
                    const Cx<R> tp = Tp[row + k];
                    ar -= mom * tp.x;
                    ai -= mom * tp.y;
                }
                const R den = r_hypot(ar, ai) + r_tiny<R>();
                const R sv = S[row + k];
                A[row + k] = Cx<R>{ar / den * sv, ai / den * sv};
                Tp[row + k] = Cx<R>{xr, xi};
            }
            __syncwarp();
        }
        __syncthreads();
    }
    const int out_len = HOP * (frames - 1);
    R* w = static_cast<R*>(p.wave) + (size_t)item * out_len;
    for (int n = tid; n < out_len; n += GL_THREADS) w[n] = ypad[n + NFFT / 2];
}

// =====================================================================================================
// Forward representation: LogSpectrogram.get_representation (tqdne/representation.py:140-150,163-169) with
// librosa 0.11 stft(x, n_fft=256, hop_length=32) semantics (center=True, zero padding, periodic Hann):
//   rep[f][t] = 2 * (log(max(|STFT(x)[f][t]|, clip)) - log_clip) / (log_max - log_clip) - 1,   f < n_fft/2
// One CTA per (sample, component): the padded signal sits in shared memory, a warp owns a frame (the same
// 128-point complex FFT + even/odd split as the inverse), and the [128 x frames] tile leaves coalesced.
template <typename R, typename TO>
__global__ void __launch_bounds__(GL_THREADS) logspec_forward_kernel(const float* wave, TO* rep, long long L, int frames,
                                                                     double clip, double log_clip, double log_max) {
    extern __shared__ __align__(16) unsigned char gl_smem[];
    const int len = NFFT + HOP * (frames - 1);
    R* ypad = reinterpret_cast<R*>(gl_smem);
    R* win = ypad + ((len + 3) & ~3);
    Cx<R>* tw128 = reinterpret_cast<Cx<R>*>(win + NFFT);
    Cx<R>* tw256 = tw128 + 64;
    Cx<R>* bufs = tw256 + 132;
    TO* tile = reinterpret_cast<TO*>(bufs + GL_WARPS * 128);  // [128][frames + 1]
    const int item = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    Cx<R>* buf = bufs + warp * 128;
    const float* x = wave + (size_t)item * L;
    for (int i = tid; i < len; i += GL_THREADS) {
        const long long j = (long long)i - NFFT / 2;
        ypad[i] = (j >= 0 && j < L) ? (R)x[j] : (R)0;
    }
    for (int i = tid; i < NFFT; i += GL_THREADS) win[i] = (R)(0.5 - 0.5 * cospi(2.0 * i / NFFT));
    for (int i = tid; i < 64; i += GL_THREADS) {
        double sn, cs;
        sincospi(-2.0 * i / 128.0, &sn, &cs);
        tw128[i] = Cx<R>{(R)cs, (R)sn};
    }
    for (int i = tid; i <= 128; i += GL_THREADS) {
        double sn, cs;
        sincospi(-2.0 * i / 256.0, &sn, &cs);
        tw256[i] = Cx<R>{(R)cs, (R)sn};
    }
    __syncthreads();
    const int pitch = frames + 1;
    const double scale = 2.0 / (log_max - log_clip);
    for (int t = warp; t < frames; t += GL_WARPS) {
        const R* yt = ypad + HOP * t;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = lane + 32 * i;
            buf[bitrev7(n)] = Cx<R>{yt[2 * n] * win[2 * n], yt[2 * n + 1] * win[2 * n + 1]};
        }
        __syncwarp();
        fft128<R, false>(buf, tw128, lane);
        for (int k = lane; k < 128; k += 32) {
            const Cx<R> zk = buf[k & 127];
            const Cx<R> zn = buf[(128 - k) & 127];
            const R er = (R)0.5 * (zk.x + zn.x), ei = (R)0.5 * (zk.y - zn.y);
            const R dr = (R)0.5 * (zk.x - zn.x), di = (R)0.5 * (zk.y + zn.y);
            const R c = tw256[k].x, sn = tw256[k].y;  // e^{-2 pi i k / 256}
            const R xr = er + (di * c + dr * sn);
            const R xi = ei + (di * sn - dr * c);
            double mag = (double)r_hypot(xr, xi);
            mag = mag < clip ? clip : mag;
            tile[k * pitch + t] = (TO)((log(mag) - log_clip) * scale - 1.0);
        }
        __syncwarp();
    }
    __syncthreads();
    TO* out = rep + (size_t)item * 128 * frames;
    for (int idx = tid; idx < 128 * frames; idx += GL_THREADS) {
        const int f = idx / frames, t = idx - f * frames;
        out[idx] = tile[f * pitch + t];
    }
}

// =====================================================================================================
// Fused Griffin-Lim (the product path, fp64 by default):  STFT(frame t) -> phase update -> inverse FFT(frame t) ->
// overlap-add, per frame, templated on the arithmetic type.
//
// The rebuilt spectrum never leaves the SM: a warp transforms frame t of the current signal, applies the fast
// Griffin-Lim update against R_prev (the only per-iteration global traffic besides S: 129 complex values read + written
// per frame, L2 resident), inverse-transforms the new angles and overlap-adds into the NEXT signal buffer.  Frames are
// processed in 8 rounds (t mod 8) separated by block barriers, so concurrently added frames never overlap and the
// summation order is fixed (bit-reproducible, no atomics).  The 128-point complex FFT is held in registers (4 points per
// lane): radix 4 x 4 x 4 x 2 with two shared-memory transposes and one shuffle stage.
//
// The kernel is bound by instruction issue (fp64: by the FP64 pipe), so the arithmetic is kept lean:
//   * a / (|a| + tiny) * S is evaluated as a * (S * rsqrt(|a|^2)) -- one reciprocal square root instead of hypot + two
//     divisions; |a|^2 that would underflow takes the literal formula.  Under fp64 the difference (1 ulp) sits ten orders
//     of magnitude below the 1e-9 parity bound, even after the ~1e3 amplification of 128 iterations.
//   * the unit phasors of the initial phase arrive as a table (host: cos / sin of 2 pi U in float64) -- no sincos in
//     the kernel; S = exp(..) once per item in the prologue.
//   * 1 / window-sum-of-squares is a per-sample table in shared memory.
namespace fused {

// Warps per CTA (template parameter FWT of the kernel): 16 with one CTA per SM, or -- fp64, where the signal buffers leave
// room for it -- 8 with TWO CTAs per SM: the same 16 warps per SM, but an item is half as wide, so the 768 items of a batch
// of 256 fill the last wave better (5.19 waves of 148 one-CTA slots was 6 rounds) and one CTA's frame-round barriers are
// covered by the other's work.  What makes two CTAs fit is the window-sum table below instead of a per-sample array.
constexpr int EX = 152;         // per-warp exchange buffer (complex elements): >= 152 (transposes), >= 129 (spectrum)

template <typename R> __device__ __forceinline__ Cx<R> cmul(Cx<R> a, Cx<R> b) { return Cx<R>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename R> __device__ __forceinline__ Cx<R> cadd(Cx<R> a, Cx<R> b) { return Cx<R>{a.x + b.x, a.y + b.y}; }
template <typename R> __device__ __forceinline__ Cx<R> csub(Cx<R> a, Cx<R> b) { return Cx<R>{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ float r_rsqrt(float v) { return rsqrtf(v); }
__device__ __forceinline__ double r_rsqrt(double v) { return rsqrt(v); }
// smallest |a|^2 the rsqrt form is used for (below it |a|^2 may have lost bits to underflow)
template <typename R> __device__ __forceinline__ R r_small2();
template <> __device__ __forceinline__ float r_small2<float>() { return 1e-30f; }
template <> __device__ __forceinline__ double r_small2<double>() { return 1e-280; }

// z[k] = sum_j z[j] W4^(jk), W4 = -i (forward) / +i (inverse)
template <typename R, bool INV>
__device__ __forceinline__ void radix4(Cx<R> (&z)[4]) {
    const Cx<R> t0 = cadd(z[0], z[2]), t1 = csub(z[0], z[2]), t2 = cadd(z[1], z[3]), t3 = csub(z[1], z[3]);
    z[0] = cadd(t0, t2);
    z[2] = csub(t0, t2);
    const Cx<R> it3 = INV ? Cx<R>{-t3.y, t3.x} : Cx<R>{t3.y, -t3.x};  // (+i or -i) * t3
    z[1] = cadd(t1, it3);
    z[3] = csub(t1, it3);
}
template <typename R, bool INV>
__device__ __forceinline__ Cx<R> twid(const Cx<R>* tw, int m) {
    Cx<R> w = tw[m];
    if (INV) w.y = -w.y;
    return w;
}
__device__ __forceinline__ float shfl16(float v) { return __shfl_xor_sync(0xffffffffu, v, 16); }
__device__ __forceinline__ double shfl16(double v) { return __shfl_xor_sync(0xffffffffu, v, 16); }

// Twiddle tables of the 128-point FFT, laid out so that every load is bank-conflict free (the generic table
// exp(-2 pi i m / 128) read at m = lane k or 4 l2 m puts the eight lanes of a quarter warp on 1-2 banks: measured 104 extra
// shared-memory wavefronts per frame, 20 % of the kernel's shared-memory traffic, which is what bounds it):
//   tw1[k - 1][lane] = W^(lane k), k = 1..3     tw2[m - 1][l2] = W^(4 l2 m), l2 < 8, m = 1..3     (W = exp(-2 pi i / 128))
constexpr int TW_ELEMS = 3 * 32 + 3 * 8;

// in : z[j] = x[lane + 32 j]            (natural order)
// out: z[p] = X[(lane & 15) + 16 p + 64 (lane >> 4)]
// ex: this warp's exchange buffer
template <typename R, bool INV>
__device__ __forceinline__ void fft128(Cx<R> (&z)[4], Cx<R>* ex, const Cx<R>* tw, int lane) {
    const Cx<R>* tw1 = tw;
    const Cx<R>* tw2 = tw + 96;
    radix4<R, INV>(z);
#pragma unroll
    for (int k = 1; k < 4; ++k) z[k] = cmul(z[k], twid<R, INV>(tw1, (k - 1) * 32 + lane));
    const int a = lane >> 3, l2 = lane & 7;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) ex[k * 40 + lane] = z[k];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = ex[a * 40 + l2 + 8 * j];
    radix4<R, INV>(z);
#pragma unroll
    for (int m = 1; m < 4; ++m) z[m] = cmul(z[m], twid<R, INV>(tw2, (m - 1) * 8 + l2));
    const int g = lane & 15, l3 = lane >> 4;
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 4; ++m) ex[9 * (a + 4 * m) + l2] = z[m];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = ex[9 * g + l3 + 2 * j];
    radix4<R, INV>(z);
    if (l3) {
        // W^16 = (1 - i) / sqrt 2, W^32 = -i, W^48 = -(1 + i) / sqrt 2 (forward); conjugates for the inverse
        const R h = (R)0.70710678118654752440;
        const Cx<R> z1 = z[1], z2 = z[2], z3 = z[3];
        if (!INV) {
            z[1] = Cx<R>{h * (z1.x + z1.y), h * (z1.y - z1.x)};
            z[2] = Cx<R>{z2.y, -z2.x};
            z[3] = Cx<R>{h * (z3.y - z3.x), -h * (z3.x + z3.y)};
        } else {
            z[1] = Cx<R>{h * (z1.x - z1.y), h * (z1.x + z1.y)};
            z[2] = Cx<R>{-z2.y, z2.x};
            z[3] = Cx<R>{-h * (z3.x + z3.y), h * (z3.x - z3.y)};
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const R ox = shfl16(z[q].x);
        const R oy = shfl16(z[q].y);
        z[q] = l3 ? Cx<R>{ox - z[q].x, oy - z[q].y} : Cx<R>{z[q].x + ox, z[q].y + oy};
    }
    __syncwarp();
}

// a / (|a| + tiny) * sv  (librosa: angles /= |angles| + tiny; angles *= S)
template <typename R>
__device__ __forceinline__ Cx<R> unit_times(R ar, R ai, R sv) {
    const R d2 = ar * ar + ai * ai;
    if (d2 > r_small2<R>()) {
        const R f = sv * r_rsqrt(d2);
        return Cx<R>{ar * f, ai * f};
    }
    const R den = r_hypot(ar, ai) + r_tiny<R>();
    return Cx<R>{ar / den * sv, ai / den * sv};
}

// 1 / (window sum of squares) of padded sample n, 0 outside the kept range [NFFT/2, len - NFFT/2).  Away from the ends all
// NFFT/HOP frames cover a sample and the sum depends on n mod HOP only; the first / last NFFT/2 kept samples have their own
// entries: table = [HOP interior | NFFT/2 head | NFFT/2 tail].  (Needs len >= 2 NFFT; shorter signals use WSS_FULL.)
constexpr int WSS_TAB = HOP + NFFT;
template <typename R>
__device__ __forceinline__ R inv_wss_at(const R* tab, int n, int len, bool full) {
    if (full) return tab[n];
    if (n < NFFT / 2 || n >= len - NFFT / 2) return (R)0;
    if (n < NFFT) return tab[HOP + n - NFFT / 2];
    if (n >= len - NFFT) return tab[HOP + NFFT / 2 + n - (len - NFFT)];
    return tab[n & (HOP - 1)];
}

template <typename R>
struct Smem {
    R* ya;        // signal buffers (padded length)
    R* yb;
    R* inv_wss;   // window-sum table (WSS_TAB entries), or one entry per padded sample (`full`)
    R* win;       // periodic Hann, 256
    Cx<R>* tw128; // TW_ELEMS: the conflict-free stage tables of fft128
    Cx<R>* tw256; // 129 (+pad)
    Cx<R>* ex;    // FW x EX
};

template <typename R, int FWT>
__global__ void __launch_bounds__(FWT * 32, (sizeof(R) == 4 || FWT == 8) ? 2 : 1) griffinlim_fused_kernel(const GlParams p) {
    constexpr int FTHREADS = FWT * 32;
    constexpr int FW = FWT;
    extern __shared__ __align__(16) unsigned char gl_smem[];
    const int frames = p.frames;
    const int len = NFFT + HOP * (frames - 1);
    const int len4 = (len + 3) & ~3;
    const bool wss_full = len < 2 * NFFT;
    const int wss_n = wss_full ? len4 : WSS_TAB;
    Smem<R> sm;
    sm.ya = reinterpret_cast<R*>(gl_smem);
    sm.yb = sm.ya + len4;
    sm.inv_wss = sm.yb + len4;
    sm.win = sm.inv_wss + wss_n;
    sm.tw128 = reinterpret_cast<Cx<R>*>(sm.win + NFFT);
    sm.tw256 = sm.tw128 + TW_ELEMS;
    sm.ex = sm.tw256 + 132;

    const int item = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    Cx<R>* ex = sm.ex + warp * EX;
    const size_t per_item = (size_t)frames * NBIN;
    R* S = static_cast<R*>(p.ws) + (size_t)item * per_item * 5;
    Cx<R>* Tp = reinterpret_cast<Cx<R>*>(S + per_item);

    for (int i = tid; i < NFFT; i += FTHREADS) sm.win[i] = (R)(0.5 - 0.5 * cospi(2.0 * i / NFFT));
    for (int i = tid; i < TW_ELEMS; i += FTHREADS) {
        // exponent of W = exp(-2 pi i / 128): lane * k for the stage-1 table, 4 * l2 * m for the stage-2 table
        const int e = i < 96 ? (i & 31) * (i / 32 + 1) : 4 * ((i - 96) & 7) * ((i - 96) / 8 + 1);
        double s, c;
        sincospi(-2.0 * e / 128.0, &s, &c);
        sm.tw128[i] = Cx<R>{(R)c, (R)s};
    }
    for (int i = tid; i <= 128; i += FTHREADS) {
        double s, c;
        sincospi(-2.0 * i / 256.0, &s, &c);
        sm.tw256[i] = Cx<R>{(R)c, (R)s};
    }
    for (int i = tid; i < len4; i += FTHREADS) {
        sm.ya[i] = 0;
        sm.yb[i] = 0;
    }
    __syncthreads();
    for (int i = tid; i < wss_n; i += FTHREADS) {
        // the padded sample this entry stands for (interior entries: any sample with all NFFT / HOP frames over it)
        int n = i;
        if (!wss_full) n = i < HOP ? NFFT + i : (i < HOP + NFFT / 2 ? NFFT / 2 + (i - HOP) : len - NFFT + (i - HOP - NFFT / 2));
        R v = 0;
        if (n >= NFFT / 2 && n < len - NFFT / 2) {
            R wss = 0;
            const int r = n & (HOP - 1);
#pragma unroll
            for (int j = 0; j < NFFT / HOP; ++j) {
                const int o = r + HOP * j;
                const int t = (n - o) / HOP;
                if (n - o >= 0 && t < frames) wss += sm.win[o] * sm.win[o];
            }
            v = wss > r_tiny<R>() ? (R)1 / wss : (R)1;
        }
        sm.inv_wss[i] = v;
    }
    // S = exp(((rep + 1) / 2) * (log_max - log_clip) + log_clip), Nyquist row = 0
    const float* rep = p.rep + (size_t)item * 128 * frames;
    for (int idx = tid; idx < frames * NBIN; idx += FTHREADS) {
        const int f = idx / frames, t = idx % frames;  // coalesced read of rep[f][t]
        R sv = 0;
        if (f < 128) {
            const float nls = (rep[(size_t)f * frames + t] + 1.f) / 2.f;
            sv = (R)exp((double)nls * (p.log_max - p.log_clip) + p.log_clip);
        }
        S[(size_t)t * NBIN + f] = sv;
    }
    __syncthreads();

    const R mom = (R)p.mom;
    const R inv_n = (R)(1.0 / 128.0);
    const R half = (R)0.5;
    // twiddles of the bins this lane owns (k = lane + 32 j), kept in registers: e^{-2 pi i k / 256}
    Cx<R> w256[4];
    int n_out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        w256[j] = sm.tw256[lane + 32 * j];
        n_out[j] = (lane & 15) + 16 * j + 64 * (lane >> 4);
    }
    R* ycur = sm.ya;
    R* ynext = sm.yb;

    for (int it = 0; it <= p.n_iter; ++it) {
        for (int round = 0; round < 8; ++round) {
            for (int t = round + 8 * warp; t < frames; t += 8 * FW) {
                const size_t row = (size_t)t * NBIN;
                // bins owned by this lane: k = lane + 32 j (j < 4); lane 0 also owns k = 128
                R sv[4], sv128 = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) sv[j] = __ldcg(S + row + lane + 32 * j);
                if (lane == 0) sv128 = __ldcg(S + row + 128);
                if (it > 0) {
                    Cx<R> tp[4], tp128 = Cx<R>{0, 0};
                    if (it > 1) {   // R_prev of the first update is absent (librosa: tprev is None)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const double2* q = reinterpret_cast<const double2*>(Tp + row + lane + 32 * j);
                            if constexpr (sizeof(R) == 8) {
                                const double2 v = __ldcg(q);
                                tp[j] = Cx<R>{(R)v.x, (R)v.y};
                            } else {
                                const float2 v = __ldcg(reinterpret_cast<const float2*>(Tp + row + lane + 32 * j));
                                tp[j] = Cx<R>{(R)v.x, (R)v.y};
                            }
                        }
                        if (lane == 0) tp128 = Tp[row + 128];
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) tp[j] = Cx<R>{0, 0};
                    }
                    // ---- forward: frame t of the current signal
                    Cx<R> z[4];
                    const R* yt = ycur + HOP * t;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ni = lane + 32 * j;
                        const Cx<R> v = *reinterpret_cast<const Cx<R>*>(yt + 2 * ni);
                        const Cx<R> w = *reinterpret_cast<const Cx<R>*>(sm.win + 2 * ni);
                        z[j] = Cx<R>{v.x * w.x, v.y * w.y};
                    }
                    fft128<R, false>(z, ex, sm.tw128, lane);
#pragma unroll
                    for (int q = 0; q < 4; ++q) ex[n_out[q]] = z[q];
                    __syncwarp();
                    Cx<R> zk[4], zn[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = lane + 32 * j;
                        zk[j] = ex[k];
                        zn[j] = ex[(128 - k) & 127];
                    }
                    const Cx<R> z0 = ex[0];
                    __syncwarp();
                    // ---- split + fast Griffin-Lim update; new angles into ex[0..128]
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = lane + 32 * j;
                        const R er = half * (zk[j].x + zn[j].x), ei = half * (zk[j].y - zn[j].y);
                        const R dr = half * (zk[j].x - zn[j].x), di = half * (zk[j].y + zn[j].y);
                        const Cx<R> w = w256[j];
                        const R xr = er + (di * w.x + dr * w.y);
                        const R xi = ei + (di * w.y - dr * w.x);
                        const R ar = xr - mom * tp[j].x, ai = xi - mom * tp[j].y;
                        ex[k] = unit_times<R>(ar, ai, sv[j]);
                        Tp[row + k] = Cx<R>{xr, xi};
                    }
                    if (lane == 0) {
                        const R xr = z0.x - z0.y;  // Nyquist bin: real
                        const R ar = xr - mom * tp128.x, ai = -mom * tp128.y;
                        ex[128] = unit_times<R>(ar, ai, sv128);
                        Tp[row + 128] = Cx<R>{xr, 0};
                    }
                } else {
                    // initial angles: S * exp(i phase0), unit phasors from the host table [129][frames]
                    const double2* ph = reinterpret_cast<const double2*>(p.phase0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = lane + 32 * j;
                        const double2 u = __ldg(ph + (size_t)k * frames + t);
                        ex[k] = Cx<R>{(R)((double)sv[j] * u.x), (R)((double)sv[j] * u.y)};
                    }
                    if (lane == 0) {
                        const double2 u = __ldg(ph + (size_t)128 * frames + t);
                        ex[128] = Cx<R>{(R)((double)sv128 * u.x), (R)((double)sv128 * u.y)};
                    }
                }
                __syncwarp();
                // ---- inverse: c2r pre-twiddle, FFT, window, overlap-add into the next signal
                Cx<R> z[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = lane + 32 * j;
                    Cx<R> xk = ex[k], xn = ex[128 - k];
                    if (k == 0) {  // c2r transforms ignore the imaginary parts of DC and Nyquist
                        xk.y = 0;
                        xn.y = 0;
                    }
                    const R er = half * (xk.x + xn.x), ei = half * (xk.y - xn.y);
                    const R dr = half * (xk.x - xn.x), di = half * (xk.y + xn.y);
                    const R c = w256[j].x, s = -w256[j].y;  // e^{+2 pi i k / 256}
                    const R orr = dr * c - di * s, oi = dr * s + di * c;
                    z[j] = Cx<R>{er - oi, ei + orr};
                }
                fft128<R, true>(z, ex, sm.tw128, lane);
                R* yo = ynext + HOP * t;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    Cx<R>* dst = reinterpret_cast<Cx<R>*>(yo + 2 * n_out[q]);
                    const Cx<R> w = *reinterpret_cast<const Cx<R>*>(sm.win + 2 * n_out[q]);
                    Cx<R> acc = *dst;
                    acc.x += w.x * (z[q].x * inv_n);
                    acc.y += w.y * (z[q].y * inv_n);
                    *dst = acc;
                }
            }
            __syncthreads();
        }
        // ---- divide by the window sum of squares (centre padding -> 0), clear the consumed buffer, swap
        for (int n = tid; n < len4; n += FTHREADS) {
            ynext[n] *= inv_wss_at<R>(sm.inv_wss, n, len, wss_full);
            ycur[n] = 0;
        }
        __syncthreads();
        R* tmp = ycur;
        ycur = ynext;
        ynext = tmp;
    }
    const int out_len = HOP * (frames - 1);
    R* w = static_cast<R*>(p.wave) + (size_t)item * out_len;
    for (int n = tid; n < out_len; n += FTHREADS) w[n] = ycur[n + NFFT / 2];
}

template <typename R>
size_t smem_bytes(int frames, int warps) {
    const int len = NFFT + HOP * (frames - 1);
    const int len4 = (len + 3) & ~3;
    const size_t wss_n = len < 2 * NFFT ? (size_t)len4 : (size_t)WSS_TAB;
    return sizeof(R) * (2 * (size_t)len4 + wss_n + NFFT) + sizeof(Cx<R>) * (TW_ELEMS + 132 + (size_t)warps * EX);
}

}  // namespace fused

template <typename R>
size_t gl_smem_bytes(int frames) {
    const int len = NFFT + HOP * (frames - 1);
    return sizeof(R) * (((len + 3) & ~3) + NFFT) + sizeof(Cx<R>) * (64 + 132 + GL_WARPS * 128);
}

}  // namespace
}  // namespace tq

using namespace tq;

extern "C" int64_t tq_griffinlim_ws_bytes(int32_t items, int32_t n_fft, int32_t frames, int32_t precision) {
    if (items <= 0 || frames <= 0 || n_fft != NFFT) return -1;
    const int64_t el = precision == TQ_F64 ? 8 : 4;
    return (int64_t)items * frames * NBIN * 5 * el;
}

extern "C" int tq_logspec_griffinlim(const float* rep, const double* phase0, void* wave, int32_t items, int32_t n_fft,
                                     int32_t hop, int32_t frames, int32_t n_iter, double log_clip, double log_max,
                                     double momentum, int32_t precision, void* ws, void* stream) {
    TQ_CHECK(n_fft == NFFT && hop == HOP, "griffinlim: only n_fft=256, hop=32 is built (reference SpectrogramConfig)");
    TQ_CHECK(rep && phase0 && wave && ws && items > 0 && frames > 0 && n_iter >= 0, "griffinlim: bad arguments");
    TQ_CHECK(frames % 2 == 0, "griffinlim: the frame count must be even");
    TQ_CHECK(precision == TQ_F32 || precision == TQ_F64, "griffinlim: precision must be TQ_F32 or TQ_F64");
    GlParams p;
    p.rep = rep; p.phase0 = phase0; p.wave = wave; p.ws = ws; p.items = items; p.frames = frames; p.n_iter = n_iter;
    p.log_clip = log_clip; p.log_max = log_max; p.mom = momentum / (1.0 + momentum);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // product path: the fused kernel in the requested arithmetic; TQ_GL_LEGACY=1 selects the unfused kernels (kept as an
    // independent cross-check of the fused ones)
    const char* legacy = getenv("TQ_GL_LEGACY");
    const bool use_legacy = legacy && legacy[0] == '1';
    if (!use_legacy && precision == TQ_F32) {
        const size_t smem = fused::smem_bytes<float>(frames, 16);
        TQ_CHECK(smem <= 227 * 1024, "griffinlim: too many frames for shared memory");
        TQ_CUDA(cudaFuncSetAttribute(fused::griffinlim_fused_kernel<float, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fused::griffinlim_fused_kernel<float, 16><<<items, 512, smem, st>>>(p);
    } else if (!use_legacy) {
        // fp64: two 8-warp CTAs per SM when both fit (2 x (smem + 1 KB reserved) <= 228 KB), else one 16-warp CTA
        const size_t smem8 = fused::smem_bytes<double>(frames, 8), smem16 = fused::smem_bytes<double>(frames, 16);
        const char* w16 = getenv("TQ_GL_WARPS16");
        if (2 * (smem8 + 1024) <= 228 * 1024 && !(w16 && w16[0] == '1')) {
            TQ_CUDA(cudaFuncSetAttribute(fused::griffinlim_fused_kernel<double, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
            fused::griffinlim_fused_kernel<double, 8><<<items, 256, smem8, st>>>(p);
        } else {
            TQ_CHECK(smem16 <= 227 * 1024, "griffinlim: too many frames for shared memory");
            TQ_CUDA(cudaFuncSetAttribute(fused::griffinlim_fused_kernel<double, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            fused::griffinlim_fused_kernel<double, 16><<<items, 512, smem16, st>>>(p);
        }
    } else if (precision == TQ_F32) {
        const size_t smem = gl_smem_bytes<float>(frames);
        TQ_CHECK(smem <= 227 * 1024, "griffinlim: too many frames for shared memory");
        TQ_CUDA(cudaFuncSetAttribute(griffinlim_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        griffinlim_kernel<float><<<items, GL_THREADS, smem, st>>>(p);
    } else {
        const size_t smem = gl_smem_bytes<double>(frames);
        TQ_CHECK(smem <= 227 * 1024, "griffinlim: too many frames for shared memory");
        TQ_CUDA(cudaFuncSetAttribute(griffinlim_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        griffinlim_kernel<double><<<items, GL_THREADS, smem, st>>>(p);
    }
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int tq_logspec_forward(const float* wave, void* rep, int32_t rep_dtype, int32_t items, int32_t n_fft, int32_t hop,
                                  int64_t L, int32_t frames, double clip, double log_max, int32_t precision, void* stream) {
    TQ_CHECK(n_fft == NFFT && hop == HOP, "logspec_forward: only n_fft=256, hop=32 is built (reference SpectrogramConfig)");
    TQ_CHECK(wave && rep && items > 0 && L > 0, "logspec_forward: bad arguments");
    TQ_CHECK(frames == 1 + (int)(L / HOP), "logspec_forward: frames must be 1 + L / hop (center=True)");
    TQ_CHECK(precision == TQ_F32 || precision == TQ_F64, "logspec_forward: precision must be TQ_F32 or TQ_F64");
    TQ_CHECK(rep_dtype == TQ_F32 || rep_dtype == TQ_F64, "logspec_forward: rep_dtype must be TQ_F32 or TQ_F64");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double log_clip = log(clip);
#define TQ_LSF(R, TO)                                                                                                   \
    do {                                                                                                                \
        const size_t smem = gl_smem_bytes<R>(frames) + (size_t)128 * (frames + 1) * sizeof(TO);                                                          \
        TQ_CHECK(smem <= 227 * 1024, "logspec_forward: too many frames for shared memory");                             \
        TQ_CUDA(cudaFuncSetAttribute(logspec_forward_kernel<R, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        logspec_forward_kernel<R, TO><<<items, GL_THREADS, smem, st>>>(wave, static_cast<TO*>(rep), L, frames, clip, log_clip, \
                                                                      log_max);                                        \
    } while (0)
    if (precision == TQ_F32 && rep_dtype == TQ_F32) TQ_LSF(float, float);
    else if (precision == TQ_F32) TQ_LSF(float, double);
    else if (rep_dtype == TQ_F32) TQ_LSF(double, float);
    else TQ_LSF(double, double);
#undef TQ_LSF
    TQ_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}
