// Development aid: per-SM throughput of SHFL vs shared-memory loads (do they share a pipe?).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double a0 = lane, a1 = lane + 1, a2 = lane + 2, a3 = lane + 3;
    int idx = lane;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 2) {       // 4 x 64-bit shuffles (= 8 SHFL)
            a0 += __shfl_sync(0xffffffffu, a0, (lane + 1) & 31); a1 += __shfl_sync(0xffffffffu, a1, (lane + 2) & 31);
            a2 += __shfl_sync(0xffffffffu, a2, (lane + 3) & 31); a3 += __shfl_sync(0xffffffffu, a3, (lane + 5) & 31);
        }
        if (MODE == 1 || MODE == 2) {       // 4 x LDS.64 conflict-free (2 wavefronts each)
            a0 += sm[idx]; a1 += sm[idx + 32]; a2 += sm[idx + 64]; a3 += sm[idx + 96];
            idx = (idx + 128) & 2047;
        }
        if (MODE == 3) {                    // 4 x LDS.128 broadcast per 8-lane group (w4-like): 4 distinct addresses per warp
            const double2* p = reinterpret_cast<const double2*>(sm) + ((lane >> 3) * 2 + (idx & 1023));
            double2 v0 = p[0], v1 = p[1], v2 = p[16], v3 = p[17];
            a0 += v0.x + v0.y; a1 += v1.x + v1.y; a2 += v2.x + v2.y; a3 += v3.x + v3.y;
            idx = (idx + 32) & 1023;
        }
        if (MODE == 4) {                    // 4 x LDS.64 all lanes same address
            a0 += sm[idx & 1023]; a1 += sm[(idx & 1023) + 1]; a2 += sm[(idx & 1023) + 2]; a3 += sm[(idx & 1023) + 3];
            idx += 4;
            idx -= lane; idx += lane;
        }
        if (MODE == 5) {                    // 4 x LDS.64 random-ish (table lookup like)
            a0 += sm[(idx * 7) & 63]; a1 += sm[(idx * 13 + 5) & 63]; a2 += sm[(idx * 11 + 3) & 63]; a3 += sm[(idx * 5 + 9) & 63];
            idx = idx * 3 + 1;
        }
        if (MODE == 6) {                    // 4 x LDS.64 from a 16-entry table (one entry per 8-byte bank)
            a0 += sm[(idx * 7) & 15]; a1 += sm[(idx * 13 + 5) & 15]; a2 += sm[(idx * 11 + 3) & 15]; a3 += sm[(idx * 5 + 9) & 15];
            idx = idx * 3 + 1;
        }
    }
    if (a0 + a1 + a2 + a3 == 1.2345) out[0] = a0;
}
template <int MODE> void run(const char* name) {
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, blocks = 148 * 2, threads = 512;
    k<MODE><<<blocks, threads>>>(out, 10); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_iters_per_sm = (double)blocks * threads / 32 * iters / 148;
    const double cyc = ms * 1e-3 * 1.965e9;
    printf("%-44s %7.3f ms  %6.2f SM-cycles per warp-iteration (4 ops)\n", name, ms, cyc / warp_iters_per_sm);
}
int main() {
    run<0>("4 x shfl.f64 (8 SHFL)");
    run<1>("4 x LDS.64 conflict-free");
    run<2>("both");
    run<3>("4 x LDS.128, 4 addresses/warp (8-lane groups)");
    run<4>("4 x LDS.64 uniform address");
    run<5>("4 x LDS.64 random in 64 entries");
    run<6>("4 x LDS.64 random in 16 entries");
    return 0;
}
