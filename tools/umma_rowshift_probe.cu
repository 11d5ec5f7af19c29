// umma_rowshift_probe.cu -- can a tcgen05.mma A operand (K-major, SWIZZLE_128B) start at an arbitrary 128 B ROW of a
// TMA-written tile, i.e. may one shared-memory halo buffer serve all taps of a 1-D / single-image-row convolution
// through descriptor row offsets?
//
// A[256 x 64] bf16 is loaded by ONE TMA box (SWIZZLE_128B) to a 1024 B aligned buffer; B[64 x 64] = identity, so
// D[m][n] = A[shift + m][n].  For shift = 0..9 the MMA (M = 128, N = 64, K = 64) is issued with the A descriptor start
// address advanced by shift * 128 B, once with matrix-base-offset 0 and once with base offset = (start >> 7) & 7
// (descriptor bits 49-51), and the result is compared with the expected rows.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I tqdne_b200/csrc tools/umma_rowshift_probe.cu -o tools/umma_rowshift_probe
//   ./tools/umma_rowshift_probe
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tq_ptx.cuh"

using namespace tq;

constexpr int SHIFTS = 10;

struct Params {
    CUtensorMap amap, bmap;
    float* out;  // [2 variants][SHIFTS][128][64]
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base, b_smem = base + 256 * 128;
    const uint32_t bar_ld = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar_ld, 1);
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(&tmem_slot), 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar_ld, 256 * 128 + 64 * 128);
        tma_load_2d(a_smem, &p.amap, bar_ld, 0, 0);
        tma_load_2d(b_smem, &p.bmap, bar_ld, 0, 0);
    }
    mbar_wait(bar_ld, 0);
    uint32_t phase = 0;
    for (int variant = 0; variant < 2; ++variant) {
        for (int s = 0; s < SHIFTS; ++s) {
            if (threadIdx.x == 0) {
                tc_fence_after();
                const uint32_t start = a_smem + s * 128;
                uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
                if (variant == 1) hi |= ((start >> 7) & 7u) << 17;  // matrix base offset, descriptor bits 49-51
                const uint32_t a_lo = ((start & 0x3FFFFu) >> 4) | (1u << 16);
                const uint32_t b_lo = ((b_smem & 0x3FFFFu) >> 4) | (1u << 16);
                constexpr uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem, umma_desc_pack(a_lo + 2u * kk, hi), umma_desc_pack(b_lo + 2u * kk, b_hi),
                              umma_idesc_bf16(128, 64), kk != 0);
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, phase);
            phase ^= 1;
            tc_fence_after();
            float* o = p.out + ((size_t)(variant * SHIFTS + s) * 128 + warp * 32 + lane) * 64;
            for (int c = 0; c < 64; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem + (uint32_t(warp * 32) << 16) + c, r);
                tmem_ld_wait();
                for (int j = 0; j < 32; ++j) o[c + j] = __uint_as_float(r[j]);
            }
            tc_fence_before();
            __syncthreads();
        }
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 64);
    }
}

static bool encode(CUtensorMap* m, void* ptr, int rows) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) return false;
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main() {
    std::vector<__nv_bfloat16> A(256 * 64), B(64 * 64);
    std::vector<float> Af(256 * 64);
    for (int r = 0; r < 256; ++r)
        for (int c = 0; c < 64; ++c) {
            Af[r * 64 + c] = (float)((r * 7 + c * 3) % 61 - 30);
            A[r * 64 + c] = __float2bfloat16(Af[r * 64 + c]);
        }
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 64; ++k) B[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *dA, *dB;
    float* dO;
    const size_t on = (size_t)2 * SHIFTS * 128 * 64;
    cudaMalloc(&dA, A.size() * 2);
    cudaMalloc(&dB, B.size() * 2);
    cudaMalloc(&dO, on * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0, on * 4);
    Params p;
    if (!encode(&p.amap, dA, 256) || !encode(&p.bmap, dB, 64)) {
        printf("tensor map encode failed\n");
        return 1;
    }
    p.out = dO;
    const int smem = 1024 + 256 * 128 + 64 * 128;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("kernel failed: %s\n", cudaGetErrorString(e));
        return 1;
    }
    std::vector<float> O(on);
    cudaMemcpy(O.data(), dO, on * 4, cudaMemcpyDeviceToHost);
    for (int v = 0; v < 2; ++v)
        for (int s = 0; s < SHIFTS; ++s) {
            int bad = 0, first = -1;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n)
                    if (O[((size_t)(v * SHIFTS + s) * 128 + m) * 64 + n] != Af[(s + m) * 64 + n]) {
                        if (first < 0) first = m * 64 + n;
                        ++bad;
                    }
            printf("base_offset %-14s shift %d rows: %s (%d of 8192 wrong%s)\n", v ? "(start>>7)&7" : "0", s,
                   bad ? "MISMATCH" : "exact", bad, bad ? "" : "");
            if (bad && first >= 0) {
                const int m = first / 64, n = first % 64;
                // which source row did this output row come from?
                int src = -1;
                for (int r = 0; r < 256 && src < 0; ++r) {
                    bool all = true;
                    for (int c = 0; c < 64 && all; ++c) all = O[((size_t)(v * SHIFTS + s) * 128 + m) * 64 + c] == Af[r * 64 + c];
                    if (all) src = r;
                }
                printf("    first wrong element: D[%d][%d]; output row %d equals A row %d (expected %d)\n", m, n, m, src, s + m);
            }
        }
    return 0;
}
