// Does the LAUNCH GEOMETRY of gn_apply (one CTA per (sample, position chunk), each streaming its own contiguous chunk, 4 x 16 B
// loads in flight per thread) cap its bandwidth, or its arithmetic?  Plain bf16 copies of a [256, 1024, 128] tensor (67 MB in,
// 67 MB out, rotating over 4 buffer pairs so that nothing stays in the 126 MB L2) in three geometries:
//   A  grid-stride over the whole tensor (one moving front), 16 B per thread, 4 in flight
//   B  gn_apply's geometry: grid (chunks, N), thread = (16 B vector of a row, row lane), rows p0 + pl + u * lanes
//   C  as B with 4 CTAs per SM forced to 8 (more warps)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o copy_geometry_probe tools/copy_geometry_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) copy_stride(const uint4* __restrict__ s, uint4* __restrict__ d, long long n) {
    const long long step = (long long)gridDim.x * 256;
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    for (; i + 3 * step < n; i += 4 * step) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(s + i + u * step);
#pragma unroll
        for (int u = 0; u < 4; ++u) d[i + u * step] = v[u];
    }
    for (; i < n; i += step) d[i] = __ldg(s + i);
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) copy_chunked(const uint4* __restrict__ s, uint4* __restrict__ d, int P, int cv, int chunks) {
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int lanes = 256 / cv, vi = threadIdx.x % cv, pl = threadIdx.x / cv;
    const int per = (P + chunks - 1) / chunks, p0 = chunk * per, p1 = min(P, p0 + per);
    const uint4* sb = s + (long long)n * P * cv + vi;
    uint4* db = d + (long long)n * P * cv + vi;
    int pix = p0 + pl;
    for (; pix + 3 * lanes < p1; pix += 4 * lanes) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(sb + (long long)(pix + u * lanes) * cv);
#pragma unroll
        for (int u = 0; u < 4; ++u) db[(long long)(pix + u * lanes) * cv] = v[u];
    }
    for (; pix < p1; pix += lanes) db[(long long)pix * cv] = __ldg(sb + (long long)pix * cv);
}

int main() {
    const int N = 256, P = 1024, C = 128, cv = C / 8;
    const long long vecs = (long long)N * P * cv;
    const size_t bytes = vecs * 16;
    uint4 *s[4], *d[4];
    for (int i = 0; i < 4; ++i) { cudaMalloc(&s[i], bytes); cudaMalloc(&d[i], bytes); cudaMemset(s[i], i, bytes); }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        for (int it = 0; it < 8; ++it) launch(s[it & 3], d[it & 3]);
        cudaEventRecord(e0);
        const int iters = 40;
        for (int it = 0; it < iters; ++it) launch(s[it & 3], d[it & 3]);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-58s %7.2f us  %6.2f TB/s (read + write)\n", name, ms * 1e3 / iters, 2.0 * bytes / (ms * 1e-3 / iters) / 1e12);
    };
    for (int g : {148 * 2, 148 * 4, 148 * 8, 148 * 16})
        { char nm[96]; snprintf(nm, sizeof nm, "A grid-stride, %d CTAs", g);
          run(nm, [&](const uint4* a, uint4* b) { copy_stride<<<g, 256>>>(a, b, vecs); }); }
    for (int ch : {1, 2, 4, 8})
        { char nm[96]; snprintf(nm, sizeof nm, "B chunked (gn_apply geometry), %d chunks x %d samples, 4 CTA/SM", ch, N);
          run(nm, [&](const uint4* a, uint4* b) { copy_chunked<4><<<dim3(ch, N), 256>>>(a, b, P, cv, ch); }); }
    for (int ch : {2, 4, 8})
        { char nm[96]; snprintf(nm, sizeof nm, "C chunked, %d chunks x %d samples, 8 CTA/SM", ch, N);
          run(nm, [&](const uint4* a, uint4* b) { copy_chunked<8><<<dim3(ch, N), 256>>>(a, b, P, cv, ch); }); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
