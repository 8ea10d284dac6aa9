#include "../../trackdlo_b200/csrc/tdlo_kernels.cuh"
#include <cstdio>
#include <vector>
#include <cmath>
using namespace tdlo;
template <int CW, bool PIVOT>
__device__ int gj_dbg(double* dbg, const double* __restrict__ AB, int n, int ld, double* __restrict__ buf, double* __restrict__ wsol) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nt = blockDim.x, nwarp = nt >> 5;
    const int ncol = n + 3;
    double* mcol = buf;                 // [2][64]
    double* pvv = buf + 128;            // [2]
    double* pivots = buf + 130;         // [64]
    double* out = buf + 194;            // [64][3]
    int* ibuf = reinterpret_cast<int*>(buf + 386);
    int* pivi = ibuf;                   // [2]
    int* prow = ibuf + 2;               // [64]
    int* flag = ibuf + 66;
    const int r0 = lane, r1 = lane + 32;
    const int cbase = w * CW;
    double a0[CW], a1[CW];
#pragma unroll
    for (int c = 0; c < CW; c++) {
        const int col = cbase + c;
        a0[c] = (r0 < n && col < ncol) ? AB[r0 * ld + col] : 0.0;
        a1[c] = (r1 < n && col < ncol) ? AB[r1 * ld + col] : 0.0;
    }
    bool used0 = r0 >= n, used1 = r1 >= n;
    int bad = 0;
    if (tid == 0) *flag = 0;
    if (w == 0) {                       // publish column 0
        int pn; double pvn;
        if (PIVOT) gj_pick(a0[0], a1[0], !used0, !used1, lane, pn, pvn);
        else { pn = 0; pvn = __shfl_sync(0xffffffffu, a0[0], 0); }
        mcol[r0] = a0[0]; mcol[r1] = a1[0];
        if (lane == 0) { pivi[0] = pn; pvv[0] = pvn; pivots[0] = pvn; prow[0] = pn; }
    }
    __syncthreads();
    const int nkb = (n + CW - 1) / CW;
    for (int kb = 0; kb < nkb; kb++) {
#pragma unroll
        for (int kk = 0; kk < CW; kk++) {
            const int k = kb * CW + kk;
            if (k >= n) break;
            const int cur = k & 1, nxt = cur ^ 1;
            const int p = pivi[cur];
            const double pv = pvv[cur];
            double m0 = mcol[cur * 64 + r0], m1 = mcol[cur * 64 + r1];
            if (p == r0) { used0 = true; m0 = 0.0; }
            if (p == r1) { used1 = true; m1 = 0.0; }
            const double rd = -rcp_fast(pv);
            bad |= !(fabs(rd) <= 1.79e308);
            const int pl = p & 31;
            const bool ph = p >= 32;
            // owner of column k+1: update it first and publish
            constexpr bool same = true;
            (void)same;
            const int co = (kk + 1 < CW) ? kk + 1 : 0;            // static
            const int wo = (kk + 1 < CW) ? kb : kb + 1;
            if (w == wo && k + 1 < n) {
                const double rj = __shfl_sync(0xffffffffu, ph ? a1[co] : a0[co], pl);
                const double q = rj * rd;
                a0[co] = fma(m0, q, a0[co]); a1[co] = fma(m1, q, a1[co]);
                int pn; double pvn;
                if (PIVOT) gj_pick(a0[co], a1[co], !used0, !used1, lane, pn, pvn);
                else { pn = k + 1; pvn = __shfl_sync(0xffffffffu, pn < 32 ? a0[co] : a1[co], pn & 31); }
                mcol[nxt * 64 + r0] = a0[co]; mcol[nxt * 64 + r1] = a1[co];
                if (lane == 0) { pivi[nxt] = pn; pvv[nxt] = pvn; pivots[k + 1] = pvn; prow[k + 1] = pn; }
            }
            // the remaining columns > k of this warp
            if (w > kb) {
#pragma unroll
                for (int c = 0; c < CW; c++) {
                    if (w == wo && c == co && k + 1 < n) continue;     // done above (only when kk == CW-1: co == 0)
                    const double rj = __shfl_sync(0xffffffffu, ph ? a1[c] : a0[c], pl);
                    const double q = rj * rd;
                    a0[c] = fma(m0, q, a0[c]); a1[c] = fma(m1, q, a1[c]);
                }
            } else if (w == kb) {
#pragma unroll
                for (int c = 0; c < CW; c++) {
                    if (c <= kk) continue;                             // static: columns <= k are finished
                    if (c == co && kk + 1 < CW && k + 1 < n) continue; // done above
                    const double rj = __shfl_sync(0xffffffffu, ph ? a1[c] : a0[c], pl);
                    const double q = rj * rd;
                    a0[c] = fma(m0, q, a0[c]); a1[c] = fma(m1, q, a1[c]);
                }
            }
            for (int c = 0; c < CW; c++) { const int col = cbase + c; if (col < ncol) { if (r0 < n) dbg[(k * 8 + r0) * 16 + col] = a0[c]; } }
            __syncthreads();
        }
    }
    // right-hand sides -> shared, solution off the pivot rows
#pragma unroll
    for (int c = 0; c < CW; c++) {
        const int col = cbase + c;
        if (col >= n && col < ncol) {
            if (r0 < n) out[r0 * 3 + (col - n)] = a0[c];
            if (r1 < n) out[r1 * 3 + (col - n)] = a1[c];
        }
    }
    if (bad) *flag = 1;
    __syncthreads();
    for (int i = tid; i < 3 * n; i += nt) {
        const int k = i / 3, d = i - 3 * k;
        wsol[i] = out[prow[k] * 3 + d] / pivots[k];
    }
    (void)nwarp;
    __syncthreads();
    return *flag;
}


__global__ void k(const double* Ain, int n, double* dbg) {
    extern __shared__ __align__(16) double smem[];
    double* AB = smem; double* buf = smem + 64 * 67; double* wsol = buf + 512;
    const int ld = n + 3;
    for (int i = threadIdx.x; i < n * ld; i += blockDim.x) AB[i] = Ain[i];
    __syncthreads();
    gj_dbg<2, false>(dbg, AB, n, ld, buf, wsol);
}
int main() {
    const int n = 4, ld = n + 3;
    std::vector<double> A(n * ld);
    for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) A[i * ld + j] = (i == j ? 4.0 : 1.0 / (1 + abs(i - j))); for (int d = 0; d < 3; d++) A[i * ld + n + d] = i + d + 1; }
    double *dA, *dbg; cudaMalloc(&dA, A.size() * 8); cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice); cudaMallocManaged(&dbg, 8 * 8 * 16 * 8);
    for (int i = 0; i < 8 * 8 * 16; i++) dbg[i] = -999;
    const int smem = (64 * 67 + 512 + 192) * 8 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<1, 224, smem>>>(dA, n, dbg); printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    for (int k2 = 0; k2 < n; k2++) { printf("after step %d\n", k2); for (int r = 0; r < n; r++) { for (int c = 0; c < ld; c++) printf(" %9.5f", dbg[(k2 * 8 + r) * 16 + c]); printf("\n"); } }
}
