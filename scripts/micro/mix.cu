// Development aid: what does an FP64 instruction cost the SMSP issue port when it is MIXED with integer / FP32 / shared-memory
// instructions?  Each mode is a loop body of independent chains (8 FP64 chains, 8 integer chains ...), timed at 4 and 8 warps
// per scheduler; the figure printed is SMSP cycles per loop trip per warp-slot (= cycles the scheduler spends on one warp's trip).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, double seed, int iseed) {
    __shared__ double sh[1024];
    double f[8];
    int g[16];
    float h[8];
    for (int i = 0; i < 8; i++) { f[i] = seed + i + threadIdx.x; h[i] = (float)(seed * i) + threadIdx.x; }
    for (int i = 0; i < 16; i++) g[i] = iseed * i + threadIdx.x;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = seed;
    __syncthreads();
    const double a = seed * 1.0001, b = seed * 0.9999;
    const float af = (float)a, bf = (float)b;
    for (int it = 0; it < iters; it++) {
        // FP64 part
        if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 5 || MODE == 6 || MODE == 9 || MODE == 10) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
        }
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[i]) : "d"(a));
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(f[i]) : "d"(a));
        }
        // integer part: 8 / 16 independent IMADs
        if (MODE == 3 || MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(g[i]) : "r"(iseed), "r"(it));
        }
        if (MODE == 4 || MODE == 8) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(g[i]) : "r"(iseed), "r"(it));
        }
        if (MODE == 5) {            // 8 FP32 FMAs
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(h[i]) : "f"(af), "f"(bf));
        }
        if (MODE == 6) {            // 8 LOP3 (alu pipe)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g[i]) : "r"(iseed), "r"(it));
        }
        if (MODE == 9) {            // 8 shared loads (64-bit), independent of the FP64 chains
#pragma unroll
            for (int i = 0; i < 8; i++) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(sh + ((g[i] + it) & 1023)))); g[i] += (int)__double2loint(v); }
        }
        if (MODE == 10) {           // 8 FP64 compares + selects
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("{ .reg .pred p; setp.lt.f64 p, %1, %2; selp.b32 %0, %0, %3, p; }" : "+r"(g[i]) : "d"(f[i]), "d"(b), "r"(it));
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += f[i] + h[i];
    for (int i = 0; i < 16; i++) s += g[i];
    if (s == 1.2345) out[0] = s;
}
template <int MODE>
void run(const char* name) {
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int threads = 256; threads <= 1024; threads *= 2) {
        if (threads == 512 + 256) continue;
        const int iters = 20000, blocks = 148;
        k<MODE><<<blocks, threads>>>(out, 100, 1.0, 3); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters, 1.0, 3); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double wps = threads / 32 / 4.0;                       // warps per scheduler
        const double cyc = ms * 1e-3 * clk * 1e3 / iters / wps;      // SMSP cycles per warp trip
        printf("%-34s %2.0f warps/SMSP  %7.2f cycles per trip  (%s)\n", name, wps, cyc, cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    run<0>("8 DFMA");
    run<1>("8 DADD");
    run<2>("8 DMUL");
    run<7>("8 IMAD");
    run<8>("16 IMAD");
    run<3>("8 DFMA + 8 IMAD");
    run<4>("8 DFMA + 16 IMAD");
    run<5>("8 DFMA + 8 FFMA");
    run<6>("8 DFMA + 8 LOP3");
    run<9>("8 DFMA + 8 (LDS.64 + IADD)");
    run<10>("8 DFMA + 8 (DSETP + SEL)");
    return 0;
}
