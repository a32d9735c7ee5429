// Instruction-fetch microbenchmark: do warps of one SM that walk the same large straight-line body at the same time
// share its instruction-cache misses?  Body = NCHUNK chunks of 256 independent-ish FFMAs (4 KB of SASS each).
//   mode 0: every warp enters the body at chunk 0 (warps walk it together)
//   mode 1: warp w enters at chunk (w * stride) % NCHUNK (warps spread over the body, like unsynchronised optimizer warps)
// Prints cycles per chunk per warp for 1, 4, 8, 12 warps per SM.      nvcc -arch=sm_100a -O3 -o icache_share icache_share.cu
#include <cstdio>
#include <cuda_runtime.h>
#define F4 a = fmaf(a, c, d); b = fmaf(b, c, d); e = fmaf(e, c, d); f = fmaf(f, c, d);
#define F16 F4 F4 F4 F4
#define F64 F16 F16 F16 F16
#define CHUNK F64 F64 F64 F64          // 256 FFMA = 4 KB
#ifndef NCHUNK
#define NCHUNK 16
#endif
#define CASE(k) case k: CHUNK
template <int NC>
__global__ void __launch_bounds__(384, 1) walk(float *out, int iters, int mode, long long *cyc)
{
    float a = threadIdx.x * 1e-9f, b = a + 1.f, e = a + 2.f, f = a + 3.f;
    const float c = 1.0000001f, d = 1e-9f;
    const int warp = threadIdx.x >> 5;
    int start = mode ? (warp * 5) % NC : 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        switch (start) {
            CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7)
#if NCHUNK > 8
            CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15)
#endif
#if NCHUNK > 16
            CASE(16) CASE(17) CASE(18) CASE(19) CASE(20) CASE(21) CASE(22) CASE(23)
#endif
        }
        start = 0;
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 12 + warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + e + f;
}
int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 384 * 4); cudaMallocManaged(&cyc, 148 * 12 * 8);
    const int iters = 200;
    for (int mode = 0; mode < 2; mode++)
        for (int warps : {1, 4, 8, 12}) {
            for (int rep = 0; rep < 2; rep++) { walk<NCHUNK><<<148, warps * 32>>>(out, iters, mode, cyc); cudaDeviceSynchronize(); }
            double s = 0; for (int w = 0; w < warps; w++) s += cyc[w];
            printf("body %3d KB mode %d warps/SM %2d: %7.1f cycles per 4 KB chunk per warp (%.2f IPC per SM)\n", NCHUNK * 4, mode, warps,
                   s / warps / iters / NCHUNK, 256.0 * warps / (s / warps / iters / NCHUNK));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
