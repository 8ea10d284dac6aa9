// Development aid: does instruction-level parallelism inside ONE warp pay on the FP64 pipe?  K interleaved exp(-z)
// chains per trip, 1 warp and 8 warps per CTA, one CTA per SM.
#include "../../trackdlo_b200/csrc/tdlo_common.cuh"
#include <cstdio>
using namespace tdlo;
template <int K>
__global__ void k(double* out, long long* cyc, int trips, double seed) {
    __shared__ double tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = exp2((double)threadIdx.x / 64.0);
    __syncthreads();
    double acc = 0.0;
    double z[K];
    for (int u = 0; u < K; u++) z[u] = seed * (1 + u) + 1e-3 * threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < trips; it++) {
        double p[K];
#pragma unroll
        for (int u = 0; u < K; u++) p[u] = exp_neg(z[u] * z[u], tab);
#pragma unroll
        for (int u = 0; u < K; u++) { acc += p[u]; z[u] += 1e-4; }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (acc == 1.2345) out[0] = acc;
}
template <int K> void run(int threads) {
    double* out; long long* cyc; cudaMalloc(&out, 8); cudaMallocManaged(&cyc, 8);
    const int trips = 2000;
    k<K><<<148, threads>>>(out, cyc, trips, 0.7); cudaDeviceSynchronize();
    k<K><<<148, threads>>>(out, cyc, trips, 0.7); cudaDeviceSynchronize();
    printf("threads/CTA %3d  K=%d: %6.1f cycles per trip, %6.1f cycles per exp (per warp)\n", threads, K, (double)cyc[0] / trips, (double)cyc[0] / trips / K);
}
int main() {
    for (int threads : {64, 256, 512, 768, 1024}) { run<1>(threads); run<4>(threads); run<8>(threads); }
    return 0;
}
