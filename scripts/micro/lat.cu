// Latency microbenchmarks on B200 (development aid): dependent-issue latencies that drive the M-step.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double* out, long long* cyc, int iters) {
    __shared__ double sh[256];
    __shared__ int shi[64];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 256; i += blockDim.x) sh[i] = 1.0 + 1e-9 * i;
    for (int i = tid; i < 64; i += blockDim.x) shi[i] = (i + 1) & 63;
    __syncthreads();
    double a = 1.0 + tid * 1e-12, b = 0.999999;
    long long t0, t1;
    // 1. dependent DFMA chain
    t0 = clock64();
    for (int i = 0; i < iters; i++) { a = fma(a, b, 1e-9); a = fma(a, b, 1e-9); a = fma(a, b, 1e-9); a = fma(a, b, 1e-9); }
    t1 = clock64(); if (tid == 0) cyc[0] = (t1 - t0) / (4 * iters);
    // 2. dependent LDS chain (pointer chase)
    int p = lane;
    t0 = clock64();
    for (int i = 0; i < iters; i++) { p = shi[p]; p = shi[p]; p = shi[p]; p = shi[p]; }
    t1 = clock64(); if (tid == 0) cyc[1] = (t1 - t0) / (4 * iters);
    // 3. __syncthreads round trip
    t0 = clock64();
    for (int i = 0; i < iters; i++) { __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); }
    t1 = clock64(); if (tid == 0) cyc[2] = (t1 - t0) / (4 * iters);
    // 4. __drcp_rn chain
    t0 = clock64();
    for (int i = 0; i < iters; i++) { a = __drcp_rn(a); a = __drcp_rn(a); a = __drcp_rn(a); a = __drcp_rn(a); }
    t1 = clock64(); if (tid == 0) cyc[3] = (t1 - t0) / (4 * iters);
    // 5. redux chain
    unsigned u = tid;
    t0 = clock64();
    for (int i = 0; i < iters; i++) { u = __reduce_max_sync(0xffffffffu, u + lane); u = __reduce_max_sync(0xffffffffu, u + lane); u = __reduce_max_sync(0xffffffffu, u + lane); u = __reduce_max_sync(0xffffffffu, u + lane); }
    t1 = clock64(); if (tid == 0) cyc[4] = (t1 - t0) / (4 * iters);
    // 6. shfl double chain
    t0 = clock64();
    for (int i = 0; i < iters; i++) { a = __shfl_xor_sync(0xffffffffu, a, 1); a = __shfl_xor_sync(0xffffffffu, a, 2); a = __shfl_xor_sync(0xffffffffu, a, 4); a = __shfl_xor_sync(0xffffffffu, a, 8); }
    t1 = clock64(); if (tid == 0) cyc[5] = (t1 - t0) / (4 * iters);
    // 7. STS -> barrier -> LDS -> DFMA (one "publish/consume" round)
    t0 = clock64();
    for (int i = 0; i < iters; i++) { sh[tid & 255] = a; __syncthreads(); a = fma(sh[(tid + 1) & 255], b, a); __syncthreads(); }
    t1 = clock64(); if (tid == 0) cyc[6] = (t1 - t0) / iters;
    // 8. dependent IMAD chain
    int q = tid;
    t0 = clock64();
    for (int i = 0; i < iters; i++) { q = q * 3 + 1; q = q * 5 + 7; q = q * 3 + 1; q = q * 5 + 7; }
    t1 = clock64(); if (tid == 0) cyc[7] = (t1 - t0) / (4 * iters);
    // 9. 22 independent LDS + 22 DFMA (the register-solver inner body)
    double r[22];
    for (int k = 0; k < 22; k++) r[k] = a + k;
    t0 = clock64();
    for (int i = 0; i < iters; i++) {
        double f[22];
#pragma unroll
        for (int k = 0; k < 22; k++) f[k] = sh[(3 * k + (i & 1)) & 255];
#pragma unroll
        for (int k = 0; k < 22; k++) r[k] = fma(f[k], b, r[k]);
    }
    t1 = clock64(); if (tid == 0) cyc[8] = (t1 - t0) / iters;
    for (int k = 0; k < 22; k++) a += r[k];
    out[tid] = a + p + u + q;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 16 * 8);
    const char* names[] = {"DFMA dep", "LDS dep", "__syncthreads", "__drcp_rn dep", "REDUX dep", "SHFL.64 dep", "STS+bar+LDS+DFMA+bar", "IMAD dep", "22xLDS+22xDFMA body"};
    for (int threads : {32, 224}) {
        k_lat<<<1, threads>>>(out, cyc, 2000); cudaDeviceSynchronize();
        printf("threads=%d:", threads);
        for (int i = 0; i < 9; i++) printf("  %s=%lld", names[i], cyc[i]);
        printf("\n");
    }
    return 0;
}
