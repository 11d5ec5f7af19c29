// pipe_probe.cu -- what does one producer/consumer mbarrier handshake cost on sm_100a?
//
// Standalone micro-benchmark (no torch): warp 0 = producer, warp 1 = consumer, S stages, K iterations, no data
// movement at all.  Variants isolate the wait primitive (try_wait / test_wait / with a %globaltimer read in the
// slow path), single-lane vs whole-warp polling, and plain arrive vs tcgen05.commit on the consumer side.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I tqdne_b200/csrc tools/pipe_probe.cu -o tools/pipe_probe
//   ./tools/pipe_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tq_ptx.cuh"

using namespace tq;

enum WaitKind { W_TRY = 0, W_TRY_GT = 1, W_TEST = 2 };

template <int WK>
__device__ __forceinline__ void wait_kind(uint32_t bar, uint32_t parity) {
    if constexpr (WK == W_TRY) {
        while (!mbar_try_wait(bar, parity)) {}
    } else if constexpr (WK == W_TRY_GT) {
        mbar_wait(bar, parity);  // the library's bounded wait: %globaltimer read per failed poll
    } else {
        uint32_t ok;
        do {
            asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        } while (!ok);
    }
}

// WARP: whole warp polls, one elected lane acts; else lane 0 alone runs the loop.  COMMIT: consumer releases the
// stage with tcgen05.commit instead of a plain arrive.  WORK: dependent integer ops per iteration in each role (to see
// whether the handshake overlaps with independent work).
// SPIN: 8 extra warps (like the idle epilogue warps of the igemm kernel) wait on a barrier that completes only at the
// end: 0 none, 1 library mbar_wait (try_wait + %globaltimer), 2 bare try_wait loop, 3 try_wait + nanosleep(256) backoff,
// 4 try_wait with a 1 ms suspend-time hint, 5 lane 0 polls with nanosleep backoff then __syncwarp
template <int WK, bool WARP, bool COMMIT, int S, int SPIN>
__global__ void __launch_bounds__(96 + 256, 1) probe_kernel(int K, unsigned long long* out) {
    __shared__ __align__(8) unsigned long long bars[2 * S + 1];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b0 = smem_u32(bars);
    auto full = [&](int s) { return b0 + 8u * s; };
    auto empty = [&](int s) { return b0 + 8u * (S + s); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(b0 + 8u * (2 * S), 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(&tmem_slot), 32);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_slot;
    const long long t0 = clock64();
    if (warp == 0) {
        if (WARP || lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < K; ++i) {
                wait_kind<WK>(empty(stage), phase ^ 1u);
                if (!WARP || elect_one()) mbar_arrive(full(stage));
                if (WARP) __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (WARP || lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < K; ++i) {
                wait_kind<WK>(full(stage), phase);
                if (COMMIT) tc_fence_after();
                if (!WARP || elect_one()) {
                    if (COMMIT) umma_commit(empty(stage));
                    else mbar_arrive(empty(stage));
                }
                if (WARP) __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
            // drain: the last commits must land before the barriers die
            if (COMMIT) {
                for (int s = 0; s < S; ++s) {
                    const int st = (stage + s) % S;
                    (void)st;
                }
            }
            if (lane == 0) {
                out[blockIdx.x] = (unsigned long long)(clock64() - t0);
                mbar_arrive(b0 + 8u * (2 * S));
            }
        }
    } else if (warp >= 3 && SPIN != 0) {
        const uint32_t done = b0 + 8u * (2 * S);
        if constexpr (SPIN == 1) {
            mbar_wait(done, 0);
        } else if constexpr (SPIN == 2) {
            while (!mbar_try_wait(done, 0)) {}
        } else if constexpr (SPIN == 3) {
            while (!mbar_try_wait(done, 0)) __nanosleep(256);
        } else if constexpr (SPIN == 4) {
            uint32_t ok;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(done), "r"(0u), "r"(1000000u) : "memory");
            } while (!ok);
        } else {
            if (lane == 0) while (!mbar_try_wait(done, 0)) __nanosleep(256);
            __syncwarp();
        }
    }
    __syncthreads();
    if (COMMIT) {
        // wait until every empty barrier completed its last phase (K/S or K/S+1 completions): simply spin a while
        if (threadIdx.x == 0) {
            const long long t1 = clock64();
            while (clock64() - t1 < 20000) {}
        }
        __syncthreads();
    }
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tm, 32);
    }
}

template <int WK, bool WARP, bool COMMIT, int S, int SPIN = 0>
void run(const char* name, int grid, int K) {
    unsigned long long* d;
    cudaMalloc(&d, grid * sizeof(unsigned long long));
    cudaMemset(d, 0, grid * sizeof(unsigned long long));
    for (int rep = 0; rep < 2; ++rep) probe_kernel<WK, WARP, COMMIT, S, SPIN><<<grid, SPIN ? 96 + 256 : 96>>>(K, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%-58s grid=%3d  CUDA error: %s\n", name, grid, cudaGetErrorString(e));
        exit(1);
    }
    std::vector<unsigned long long> h(grid);
    cudaMemcpy(h.data(), d, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double s = 0, mx = 0;
    for (auto v : h) { s += (double)v; if ((double)v > mx) mx = (double)v; }
    printf("%-58s grid=%3d S=%d  avg %.1f  max %.1f cycles/handshake\n", name, grid, S, s / grid / K, mx / K);
    cudaFree(d);
}

int main() {
    const int K = 20000;
    for (int grid : {148}) {
        run<W_TRY, true, true, 4, 1>("8 spinners: mbar_wait (globaltimer) | warp+elect try_wait commit", grid, K);
        run<W_TRY, true, true, 4, 2>("8 spinners: bare try_wait        | warp+elect try_wait commit", grid, K);
        run<W_TRY, true, true, 4, 3>("8 spinners: try_wait+nanosleep   | warp+elect try_wait commit", grid, K);
        run<W_TRY, true, true, 4, 4>("8 spinners: try_wait 1ms hint    | warp+elect try_wait commit", grid, K);
        run<W_TRY, true, true, 4, 5>("8 spinners: lane0 poll+nanosleep | warp+elect try_wait commit", grid, K);
        run<W_TRY_GT, true, true, 4, 1>("8 spinners: mbar_wait (gt)      | warp+elect mbar_wait(gt) commit", grid, K);
        run<W_TRY_GT, false, true, 4, 1>("8 spinners: mbar_wait (gt)      | lane0 mbar_wait(gt) commit", grid, K);
        run<W_TRY_GT, true, true, 4, 5>("8 spinners: lane0 poll+nanosleep | warp+elect mbar_wait(gt) commit", grid, K);
        run<W_TRY, false, false, 4>("lane0, try_wait, plain arrive", grid, K);
        run<W_TRY_GT, false, false, 4>("lane0, try_wait + globaltimer slow path, plain arrive", grid, K);
        run<W_TEST, false, false, 4>("lane0, test_wait spin, plain arrive", grid, K);
        run<W_TRY, true, false, 4>("warp+elect, try_wait, plain arrive", grid, K);
        run<W_TRY_GT, true, false, 4>("warp+elect, try_wait + globaltimer, plain arrive", grid, K);
        run<W_TEST, true, false, 4>("warp+elect, test_wait spin, plain arrive", grid, K);
        run<W_TRY, false, true, 4>("lane0, try_wait, tcgen05.commit", grid, K);
        run<W_TRY, true, true, 4>("warp+elect, try_wait, tcgen05.commit", grid, K);
        run<W_TEST, true, true, 4>("warp+elect, test_wait, tcgen05.commit", grid, K);
        run<W_TRY, true, true, 8>("warp+elect, try_wait, tcgen05.commit", grid, K);
        run<W_TRY, true, true, 2>("warp+elect, try_wait, tcgen05.commit", grid, K);
        run<W_TRY, true, false, 8>("warp+elect, try_wait, plain arrive", grid, K);
        run<W_TRY, true, false, 2>("warp+elect, try_wait, plain arrive", grid, K);
    }
    return 0;
}
