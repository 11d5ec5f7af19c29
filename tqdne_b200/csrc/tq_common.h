// tq_common.h -- host-side internals shared by the translation units of libtqdne_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "../../include/tqdne_b200.h"

namespace tq {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define TQ_CHECK(cond, ...)          \
    do {                             \
        if (!(cond)) {               \
            tq::set_error(__VA_ARGS__); \
            return 1;                \
        }                            \
    } while (0)

#define TQ_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (void)cudaGetLastError(); /* do not leave a stale error for the next launch check */ \
            tq::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

struct Op {
    std::string name;
    std::function<int(cudaStream_t)> launch;
    bool small = false;  // latency-bound launch (single wave, few microseconds): a candidate for a programmatic edge
};

int device_sm_count();
// cudaFuncSetAttribute (opt-in shared-memory sizes) is PER DEVICE: `static PerDeviceOnce once; if (once.first()) { ... }`
// runs an initialiser once per device ordinal (a process-wide `static bool` left every device but the first without the
// attribute: launches on cuda:1 failed with "invalid argument")
// the same for attributes that grow with the request: the largest size set so far on the current device
struct PerDeviceMax {
    size_t cur[64] = {};
    bool raise(size_t v) {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (v <= cur[d]) return false;
        cur[d] = v;
        return true;
    }
};
struct PerDeviceOnce {
    bool done[64] = {};
    bool first() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};
// Programmatic dependent launch (opt-in: TQ_PDL=1): consecutive kernels of a plan overlap the next kernel's
// launch latency and prologue with the previous kernel's drain; every kernel launched this way executes
// griddepcontrol.wait before it touches memory a predecessor may still be writing.
bool pdl_enabled();
// TQ_PDL: 0 = never, 1 = every kernel of a plan, 2 = only where a small op follows another op (set per op by the plan
// runner through pdl_set_current)
int pdl_mode();
void pdl_set_current(bool on);

#ifdef __CUDACC__
// dropout (nn.Dropout, tqdne/unet.py:100-108): element i is dropped when hash(seed, i) < p -- a counter-based splitmix64,
// so the backward pass regenerates the forward's decisions from (seed, element index) and no mask tensor is stored
__device__ __forceinline__ float dropout_scale(unsigned long long seed, long long i, float p, float keep_scale) {
    // 32-bit counter hash (two multiply / xor-shift rounds): ~8 integer instructions per element -- the 64-bit splitmix
    // version made the GroupNorm backward passes, which evaluate it twice per element, measurably slower
    unsigned int x = (unsigned int)i * 0x9E3779B1u + (unsigned int)seed;
    x ^= (unsigned int)(seed >> 32) + (unsigned int)((unsigned long long)i >> 32);
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    // (float)(x >> 8) * 2^-24 < p, decided on the integers (both conversions are exact, so this is the same decision
    // bit for bit; the int -> float conversion runs on the quarter-rate XU pipe and was 1/4 of the hash's cost)
    const unsigned int thr = (unsigned int)ceilf(p * 16777216.f);
    return (x >> 8) < thr ? 0.f : keep_scale;
}
// the same decisions for 8 consecutive elements idx0 .. idx0 + 7 (idx0 a multiple of 8, so the low word never carries):
// the 64-bit index arithmetic and the seed folding are done once per vector, m[j] = 0 or keep_scale
__device__ __forceinline__ void dropout_scale8(unsigned long long seed, long long idx0, float p, float keep_scale, float (&m)[8]) {
    const unsigned int base = (unsigned int)idx0 * 0x9E3779B1u + (unsigned int)seed;
    const unsigned int hi = (unsigned int)(seed >> 32) + (unsigned int)((unsigned long long)idx0 >> 32);
    const unsigned int thr = (unsigned int)ceilf(p * 16777216.f);
    const unsigned int thr8 = thr >= 0x1000000u ? 0xFFFFFFFFu : thr << 8;   // (x >> 8) < thr  <=>  x < thr * 256
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        unsigned int x = (base + (unsigned int)j * 0x9E3779B1u) ^ hi;
        x ^= x >> 16;
        x *= 0x7FEB352Du;
        x ^= x >> 15;
        x *= 0x846CA68Bu;
        x ^= x >> 16;
        m[j] = x < thr8 ? 0.f : keep_scale;
    }
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// builders implemented in the kernel translation units; each appends >= 1 Op
int build_conv_sm100(std::vector<Op>& ops, const tq_conv_desc& d);
int build_conv_simt(std::vector<Op>& ops, const tq_conv_desc& d);
// statistics parts per sample (tq_conv_desc.stats) each conv kernel writes for a geometry
int conv_stats_parts_sm100(const tq_conv_desc& d);
int conv_stats_parts_simt(const tq_conv_desc& d);
int build_groupnorm(std::vector<Op>& ops, const tq_gn_desc& d);
int build_attention(std::vector<Op>& ops, const tq_attn_desc& d);
// tcgen05 attention (tq_attn_sm100.cu): bf16, head dim 64 / 128, 32 < T <= 512, all keys resident in shared memory
bool attention_tc_supported(const tq_attn_desc& d);
bool attention_writes_lse(const tq_attn_desc& d);
int build_attention_tc(std::vector<Op>& ops, const tq_attn_desc& d);
int build_linear(std::vector<Op>& ops, const tq_linear_desc& d);
int build_fourier(std::vector<Op>& ops, const float* t, const float* W, int M, int half, float* feat);
int build_resample2(std::vector<Op>& ops, int dtype, const void* x, void* y, int N, int H, int W, int C, int mode);
int build_spatial_mean(std::vector<Op>& ops, const float* x, int N, int P, int C, int ld, float* y);

}  // namespace tq

struct tq_plan {
    std::vector<tq::Op> ops;
    bool use_graph = false;
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraph_t graph = nullptr;
    size_t graph_ops = 0;
};
