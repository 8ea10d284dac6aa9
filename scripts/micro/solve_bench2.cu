// Development aid: times gj_solve_regs (register-resident) against gj_solve_small (shared-memory) in isolation.
#include "../../trackdlo_b200/csrc/tdlo_kernels.cuh"
#include <cstdio>
#include <vector>
#include <cmath>
using namespace tdlo;
__global__ void k_solve(const double* Ain, int n, double* Wout, long long* cyc, int which, int reps) {
    extern __shared__ __align__(16) double smem[];
    double* AB = smem;                       // n x (n+3)
    double* buf = smem + 64 * 67;            // 512 doubles
    double* wsol = buf + 512;                // 192
    int* pivs = reinterpret_cast<int*>(wsol + 192);
    const int ld = n + 3;
    long long tot = 0;
    for (int r = 0; r < reps; r++) {
        for (int i = threadIdx.x; i < n * ld; i += blockDim.x) AB[i] = Ain[i];
        __syncthreads();
        const long long t0 = clock64();
        int bad;
        if (which == 0) bad = gj_solve_regs<8, false>(AB, n, ld, buf, wsol);
        else if (which == 1) bad = gj_solve_regs<8, true>(AB, n, ld, buf, wsol);
        else if (which == 2) bad = gj_solve_regs<10, false>(AB, n, ld, buf, wsol);
        else if (which == 3) bad = gj_solve_regs<10, true>(AB, n, ld, buf, wsol);
        else if (which == 4) bad = gj_solve_small(AB, n, ld, buf, pivs, wsol, true);
        else bad = gj_solve_small(AB, n, ld, buf, pivs, wsol, false);
        tot += clock64() - t0;
        if (bad && threadIdx.x == 0 && r == 0) printf("singular! which=%d\n", which);
    }
    if (threadIdx.x == 0) cyc[which] = tot / reps;
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) Wout[which * 192 + i] = wsol[i];
}
int main() {
    for (int n : {50, 30, 53, 64, 7}) {
        const int ld = n + 3;
        std::vector<double> A(n * ld), G(n * n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double d = fabs(i - j) * 0.017; G[i * n + j] = exp(-1.414 * d / 0.35) * (2 * d + 0.5) / 0.49; }
        for (int i = 0; i < n; i++) {
            const double di = sqrt(300.0 + 50 * sin(i));
            for (int j = 0; j < n; j++) { const double dj = sqrt(300.0 + 50 * sin(j)); A[i * ld + j] = di * G[i * n + j] * dj + (i == j ? 0.5 : 0.0); }
            for (int d = 0; d < 3; d++) A[i * ld + n + d] = sin(0.3 * i + d);
        }
        double *dA, *dW; long long* cyc;
        cudaMalloc(&dA, A.size() * 8); cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
        cudaMallocManaged(&dW, 12 * 192 * 8); cudaMallocManaged(&cyc, 16 * 8);
        const int smem = (64 * 67 + 512 + 192) * 8 + 1024;
        cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const char* names[] = {"regs<8,nopivot>", "regs<8,pivot>", "regs<10,nopivot>", "regs<10,pivot>", "small(pivot)", "small(nopivot)"};
        printf("n = %d, 224 threads\n", n);
        for (int w = 0; w < 6; w++) {
            cyc[w] = 0;
            if ((w < 2) && 7 * 8 < n + 3) continue;
            k_solve<<<1, 224, smem>>>(dA, n, dW, cyc, w, 20); cudaError_t e = cudaDeviceSynchronize(); if (e) printf("err %s\n", cudaGetErrorString(e));
        }
        for (int w = 0; w < 6; w++) {
            if (!cyc[w]) continue;
            double err = 0; for (int i = 0; i < 3 * n; i++) err = fmax(err, fabs(dW[w * 192 + i] - dW[4 * 192 + i]));
            printf("  %-18s %8lld cycles/solve (%5lld per step)  max|W - W_small(pivot)| = %.2e\n", names[w], cyc[w], cyc[w] / n, err);
        }
        cudaFree(dA); cudaFree(dW); cudaFree(cyc);
    }
    return 0;
}
