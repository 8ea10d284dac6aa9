// Development aid: times the in-kernel dense solvers of tdlo_kernels.cuh in isolation (one CTA).
#include "../../trackdlo_b200/csrc/tdlo_kernels.cuh"
#include <cstdio>
#include <vector>
#include <cmath>
using namespace tdlo;

// ---- experiments: where do the cycles of a step go?
__device__ int exp_floor(double* AB, int n, int ld, double* buf) {          // barrier + warp-0 rcp chain only
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = 0; k < n; k++) {
        if (warp == 0) { const double rp = __drcp_rn(AB[k * ld + k]); buf[(k & 1) * 64 + lane] = rp * AB[lane * ld + k]; }
        __syncthreads();
    }
    return 0;
}
__device__ int exp_update_only(double* AB, int n, int ld, double* buf) {     // barrier + 17-row update, fixed multipliers
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncol = n + 3;
    for (int k = 0; k < n; k++) {
        if (warp >= 1) {
            const int c = (warp - 1) & 1, g = (warp - 1) >> 1;
            const int j = k + 2 + 32 * c + lane;
            if (j < ncol) {
                const double q = -AB[k * ld + j];
                double* pp = AB + g * ld + j;
                const double* fp = buf + g;
                const int sr = 3 * ld;
                int cnt = (n - g + 2) / 3;
                for (; cnt >= 6; cnt -= 6) {
                    const double x0 = pp[0], x1 = pp[sr], x2 = pp[2 * sr], x3 = pp[3 * sr], x4 = pp[4 * sr], x5 = pp[5 * sr];
                    const double f0 = fp[0], f1 = fp[3], f2 = fp[6], f3 = fp[9], f4 = fp[12], f5 = fp[15];
                    pp[0] = fma(f0, q, x0); pp[sr] = fma(f1, q, x1); pp[2 * sr] = fma(f2, q, x2);
                    pp[3 * sr] = fma(f3, q, x3); pp[4 * sr] = fma(f4, q, x4); pp[5 * sr] = fma(f5, q, x5);
                    pp += 6 * sr; fp += 18;
                }
                for (; cnt > 0; cnt--) { pp[0] = fma(fp[0], q, pp[0]); pp += sr; fp += 3; }
            }
        }
        __syncthreads();
    }
    return 0;
}

template <int MODE>
__device__ int exp_update_var(double* AB, int n, int ld, double* buf) {
    // MODE 0: one batch of 6 rows only; 1: full rows, loads+fma but NO stores (accumulate); 2: full rows, f from registers (no f loads)
    // MODE 3: register-resident body (22 LDS f + 22 DFMA), inlined, shared-space pointers
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncol = n + 3;
    double acc = 0.0;
    double a[22];
    for (int r = 0; r < 22; r++) a[r] = lane + r;
    for (int k = 0; k < n; k++) {
        if (warp >= 1) {
            const int c = (warp - 1) & 1, g = (warp - 1) >> 1;
            const int j = k + 2 + 32 * c + lane;
            if (j < ncol) {
                const double q = -AB[k * ld + j];
                double* pp = AB + g * ld + j;
                const double* fp = buf + g;
                const int sr = 3 * ld;
                if (MODE == 3) {
                    double f[22];
#pragma unroll
                    for (int r = 0; r < 22; r++) f[r] = fp[3 * r];
#pragma unroll
                    for (int r = 0; r < 22; r++) a[r] = fma(f[r], q, a[r]);
                } else {
                    int cnt = MODE == 0 ? 6 : (n - g + 2) / 3;
                    for (; cnt >= 6; cnt -= 6) {
                        const double x0 = pp[0], x1 = pp[sr], x2 = pp[2 * sr], x3 = pp[3 * sr], x4 = pp[4 * sr], x5 = pp[5 * sr];
                        double f0 = 1.0, f1 = 1.1, f2 = 1.2, f3 = 1.3, f4 = 1.4, f5 = 1.5;
                        if (MODE != 2) { f0 = fp[0]; f1 = fp[3]; f2 = fp[6]; f3 = fp[9]; f4 = fp[12]; f5 = fp[15]; }
                        if (MODE == 1) acc += fma(f0, q, x0) + fma(f1, q, x1) + fma(f2, q, x2) + fma(f3, q, x3) + fma(f4, q, x4) + fma(f5, q, x5);
                        else {
                            pp[0] = fma(f0, q, x0); pp[sr] = fma(f1, q, x1); pp[2 * sr] = fma(f2, q, x2);
                            pp[3 * sr] = fma(f3, q, x3); pp[4 * sr] = fma(f4, q, x4); pp[5 * sr] = fma(f5, q, x5);
                        }
                        pp += 6 * sr; fp += 18;
                    }
                }
            }
        }
        __syncthreads();
    }
    for (int r = 0; r < 22; r++) acc += a[r];
    if (acc == 12345.678) buf[0] = acc;
    return 0;
}
__global__ void k_solve(const double* Ain, int n, double* Wout, long long* cyc, int which, int reps) {
    extern __shared__ __align__(16) double smem[];
    double* AB = smem;                       // n x (n+3)
    double* buf = smem + 64 * 67;            // 512 doubles
    double* wsol = buf + 512;                // 192
    int* pivs = reinterpret_cast<int*>(wsol + 192);
    int* used = pivs + 64;
    const int ld = n + 3;
    long long tot = 0;
    for (int r = 0; r < reps; r++) {
        for (int i = threadIdx.x; i < n * ld; i += blockDim.x) AB[i] = Ain[i];
        __syncthreads();
        const long long t0 = clock64();
        int bad;
        if (which == 0) bad = gj_solve_reg<false, 18>(AB, n, ld, buf, wsol);
        else if (which == 1) bad = gj_solve_reg<true, 18>(AB, n, ld, buf, wsol);
        else if (which == 2) bad = gj_solve_small(AB, n, ld, buf, pivs, wsol, true);
        else if (which == 3) bad = gj_solve_small(AB, n, ld, buf, pivs, wsol, false);
        else if (which == 4) bad = gj_solve(AB, n, ld, pivs, used, buf + 300, wsol);
        else if (which == 5) bad = exp_floor(AB, n, ld, buf);
        else if (which == 6) bad = exp_update_only(AB, n, ld, buf);
        else if (which == 7) bad = exp_update_var<0>(AB, n, ld, buf);
        else if (which == 8) bad = exp_update_var<1>(AB, n, ld, buf);
        else if (which == 9) bad = exp_update_var<2>(AB, n, ld, buf);
        else bad = exp_update_var<3>(AB, n, ld, buf);
        tot += clock64() - t0;
        if (bad && threadIdx.x == 0) printf("singular!\n");
    }
    if (threadIdx.x == 0) cyc[which] = tot / reps;
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) Wout[which * 192 + i] = wsol[i];
}
int main() {
    const int n = 50, ld = n + 3;
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
    const char* names[] = {"reg<nopivot>", "reg<pivot>", "small(pivot)", "small(nopivot)", "generic", "exp:floor", "exp:update-only", "exp:1 batch", "exp:no stores", "exp:no f loads", "exp:reg body"};
    for (int w = 0; w < 11; w++) { k_solve<<<1, 224, smem>>>(dA, n, dW, cyc, w, 20); cudaError_t e = cudaDeviceSynchronize(); if (e) printf("err %s\n", cudaGetErrorString(e)); }
    for (int w = 0; w < 11; w++) {
        double err = 0; for (int i = 0; i < 3 * n; i++) err = fmax(err, fabs(dW[w * 192 + i] - dW[4 * 192 + i]));
        printf("%-16s %8lld cycles/solve (%5lld per step)  max|W - W_generic| = %.2e\n", names[w], cyc[w], cyc[w] / n, err);
    }
    return 0;
}
