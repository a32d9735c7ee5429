// development probe: dependent-chain latencies on this GPU (one warp)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *t, const double *in)
{
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = in[lane]; sm[lane + 32] = in[lane];
    __syncwarp();
    double a = in[lane], b = in[32 + lane];
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) a = fma(a, b, 1e-9);
    long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) a = a + b;
    long long t2 = clock64();
#pragma unroll
    for (int i = 0; i < 32; i++) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1e-30;
    long long t3 = clock64();
    int idx = lane;
#pragma unroll
    for (int i = 0; i < 32; i++) { double v = sm[idx & 63]; idx = (int)v + lane; a += v; }
    long long t4 = clock64();
#pragma unroll
    for (int i = 0; i < 16; i++) a = b / (a + 3.0);
    long long t5 = clock64();
#pragma unroll
    for (int i = 0; i < 16; i++) a = sqrt(a + 3.0);
    long long t6 = clock64();
    int q = (int)a + 12345;
#pragma unroll
    for (int i = 0; i < 16; i++) q = (int)((double)q * b) + lane;
    long long t7 = clock64();
    float f = (float)a;
#pragma unroll
    for (int i = 0; i < 64; i++) f = fmaf(f, 1.0001f, 1e-9f);
    long long t8 = clock64();
    out[lane] = a + idx + q + f;
    if (lane == 0) { t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; t[5] = t6 - t5; t[6] = t7 - t6; t[7] = t8 - t7; }
}
int main()
{
    double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + i * 1e-3;
    double *in, *out; long long *t;
    cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, 32 * 8); cudaMalloc(&t, 64);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; rep++) k<<<1, 32>>>(out, t, in);
    long long ht[8]; cudaMemcpy(ht, t, 64, cudaMemcpyDeviceToHost);
    printf("DFMA dep chain %.1f cyc/op\nDADD dep chain %.1f\nSHFL(64-bit)+DADD %.1f\nLDS.64 -> F2I -> addr -> LDS chain %.1f\n1.0/x (+DADD) %.1f\nsqrt (+DADD) %.1f\nI2F.F64 * -> F2I chain %.1f\nFFMA dep chain %.1f\n",
           ht[0] / 64.0, ht[1] / 64.0, ht[2] / 32.0, ht[3] / 32.0, ht[4] / 16.0, ht[5] / 16.0, ht[6] / 16.0, ht[7] / 64.0);
}
