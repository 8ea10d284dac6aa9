// Microbenchmark for the question BASELINE.json's north_star asks about the Nn = 200 dense solve: is a tcgen05 (5th-gen
// tensor core, TMEM accumulators) trailing update worth having next to the FP64 DMMA one?  One CTA (the M-step of a frame
// runs on one CTA) computes the symmetric update C[256 x 256] = A A^T, A = [256 x 96] (the panel-update GEMM of a blocked
// Cholesky at Nn <= 256 with 96 accumulated panel columns), two ways:
//   (a) tcgen05.mma kind::tf32, operands in shared memory (canonical K-major no-swizzle layout, UMMA descriptors), FP32
//       accumulators in TMEM, 3xTF32 split (hi*hi + hi*lo + lo*hi) for ~2^-21 relative accuracy, tcgen05.ld epilogue;
//   (b) mma.sync.m8n8k4.f64 (DMMA) with operands in shared memory, FP64 accumulators in registers.
// Prints the time per GEMM (SM cycles and microseconds), the achieved rate per SM and the error against FP64.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_trailing tc_trailing.cu ; run on one B200.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int MR = 256, KT = 96;
constexpr uint32_t LBO = 128;                    // bytes between the two K core matrices (8 rows x 16 B each)
constexpr uint32_t SBO = (KT / 4) * 128;         // bytes between 8-row groups

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }

// UMMA shared-memory descriptor (SWIZZLE_NONE, K-major): start address, leading / stride byte offsets in 16-byte units, version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((LBO >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((SBO >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    return d;
}
// instruction descriptor: D = F32 (c_format 1, bits 4-5), A = B = TF32 (format 2, bits 7-9 / 10-12), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u));
}

__global__ void __launch_bounds__(128) tc_gemm_kernel(const double* __restrict__ A, float* __restrict__ C, int reps, unsigned long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint32_t* a_hi = reinterpret_cast<uint32_t*>(smem);
    uint32_t* a_lo = a_hi + MR * KT;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int idx = tid; idx < MR * KT; idx += blockDim.x) {
        const int r = idx / KT, k = idx - r * KT;
        const float a = (float)A[idx];
        const uint32_t hi = to_tf32(a);
        const uint32_t lo = to_tf32(a - __uint_as_float(hi));
        const int w = (r >> 3) * (SBO / 4) + (k >> 2) * (LBO / 4) + (r & 7) * 4 + (k & 3);
        a_hi[w] = hi; a_lo[w] = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core's async proxy
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t hi_s = smem_u32(a_hi), lo_s = smem_u32(a_lo);
    uint32_t phase = 0;
    float chk = 0.f;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; rep++) {
        if (tid == 0) {
            for (int mt = 0; mt < 2; mt++) {
                for (int prod = 0; prod < 3; prod++) {
                    const uint32_t as = (prod == 2 ? lo_s : hi_s) + mt * 16 * SBO;      // rows 128 mt ..
                    const uint32_t bs = (prod == 1 ? lo_s : hi_s);
                    for (int ks = 0; ks < KT / 8; ks++)
                        umma_tf32(tmem + mt * 256, umma_desc(as + ks * 2 * LBO), umma_desc(bs + ks * 2 * LBO), (prod | ks) != 0);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        }
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        } while (!ok);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;");
        // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (row = lane), 8 columns per load
        for (int mt = 0; mt < 2; mt++) {
            for (int c0 = 0; c0 < 256; c0 += 8) {
                uint32_t v[8];
                const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + mt * 256 + c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (rep == reps - 1) {
                    float* dst = C + (size_t)(mt * 128 + 32 * warp + lane) * MR + c0;
#pragma unroll
                    for (int j = 0; j < 8; j++) dst[j] = __uint_as_float(v[j]);
                } else chk += __uint_as_float(v[0]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    const long long t1 = clock64();
    if (tid == 0) { cyc[0] = (unsigned long long)(t1 - t0); cyc[1] = (unsigned long long)__float_as_uint(chk); }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}

// (b) DMMA: C[256 x 256] = A A^T with mma.sync.m8n8k4.f64; A (fp64, [256][96+pad]) in shared memory; 8 warps, each owns 32 rows,
// processes its 32 x 256 strip as 4 x 32 tiles of 8 x 8 in passes of 4 x 8 tiles (64 accumulator registers).
constexpr int LDA = KT + 2;
__global__ void __launch_bounds__(256) dmma_gemm_kernel(const double* __restrict__ A, double* __restrict__ C, int reps, unsigned long long* cyc) {
    extern __shared__ __align__(16) unsigned char smem2[];
    double* sa = reinterpret_cast<double*>(smem2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int idx = tid; idx < MR * KT; idx += blockDim.x) { const int r = idx / KT, k = idx - r * KT; sa[r * LDA + k] = A[idx]; }
    __syncthreads();
    const int fr = lane >> 2, fk = lane & 3;
    double chk = 0.0;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; rep++) {
        for (int cb = 0; cb < 256; cb += 64) {                     // 8 column tiles at a time
            double c0[4][8], c1[4][8];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) { c0[i][j] = 0.0; c1[i][j] = 0.0; }
            for (int kk = 0; kk < KT; kk += 4) {
                double av[4], bv[8];
#pragma unroll
                for (int i = 0; i < 4; i++) av[i] = sa[(warp * 32 + i * 8 + fr) * LDA + kk + fk];
#pragma unroll
                for (int j = 0; j < 8; j++) bv[j] = sa[(cb + j * 8 + fr) * LDA + kk + fk];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i][j]), "+d"(c1[i][j]) : "d"(av[i]), "d"(bv[j]));
            }
            if (rep == reps - 1) {
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        double* dst = C + (size_t)(warp * 32 + i * 8 + fr) * MR + cb + j * 8 + 2 * fk;
                        dst[0] = c0[i][j]; dst[1] = c1[i][j];
                    }
            } else chk += c0[0][0];
        }
    }
    const long long t1 = clock64();
    if (tid == 0) { cyc[0] = (unsigned long long)(t1 - t0); cyc[1] = (unsigned long long)__double_as_longlong(chk); }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main() {
    std::vector<double> hA((size_t)MR * KT);
    srand(7);
    for (auto& v : hA) v = (rand() / (double)RAND_MAX - 0.5) * 2.0;
    std::vector<double> ref((size_t)MR * MR);
    for (int i = 0; i < MR; i++) for (int j = 0; j < MR; j++) { double s = 0; for (int k = 0; k < KT; k++) s += hA[i * KT + k] * hA[j * KT + k]; ref[(size_t)i * MR + j] = s; }
    double *dA, *dC64; float* dC32; unsigned long long* dcyc;
    CK(cudaMalloc(&dA, hA.size() * 8)); CK(cudaMalloc(&dC64, ref.size() * 8)); CK(cudaMalloc(&dC32, ref.size() * 4)); CK(cudaMalloc(&dcyc, 16));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const int reps = 200;
    const double flop = 2.0 * MR * MR * KT;
    unsigned long long cyc[2];
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    // (a) tcgen05
    const int smem_a = 2 * MR * KT * 4;
    CK(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a));
    tc_gemm_kernel<<<1, 128, smem_a>>>(dA, dC32, 2, dcyc); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); tc_gemm_kernel<<<1, 128, smem_a>>>(dA, dC32, reps, dcyc); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    CK(cudaMemcpy(cyc, dcyc, 16, cudaMemcpyDeviceToHost));
    std::vector<float> c32(ref.size()); CK(cudaMemcpy(c32.data(), dC32, c32.size() * 4, cudaMemcpyDeviceToHost));
    double err = 0, mx = 0; for (size_t i = 0; i < ref.size(); i++) { err = fmax(err, fabs(c32[i] - ref[i])); mx = fmax(mx, fabs(ref[i])); }
    std::printf("tcgen05 kind::tf32 3xTF32 (M128 N256 K8 x %d MMAs / GEMM, FP32 accumulators in TMEM, tcgen05.ld epilogue): %.0f cycles / GEMM = %.2f us (event: %.2f us); "
                "%.2f TFLOP/s per SM on the FP64-equivalent flop (%.1f MFLOP), %.2f TF32 TFLOP/s issued; max |err| / max |C| = %.2e\n",
                3 * 2 * KT / 8, (double)cyc[0] / reps, (double)cyc[0] / reps / (clk_khz * 1e-3), ms * 1e3 / reps, flop / ((double)cyc[0] / reps / (clk_khz * 1e3)) / 1e12, flop / 1e6,
                3 * flop / ((double)cyc[0] / reps / (clk_khz * 1e3)) / 1e12, err / mx);
    // (b) DMMA
    const int smem_b = MR * LDA * 8;
    CK(cudaFuncSetAttribute(dmma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
    dmma_gemm_kernel<<<1, 256, smem_b>>>(dA, dC64, 2, dcyc); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); dmma_gemm_kernel<<<1, 256, smem_b>>>(dA, dC64, reps, dcyc); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    CK(cudaMemcpy(cyc, dcyc, 16, cudaMemcpyDeviceToHost));
    std::vector<double> c64(ref.size()); CK(cudaMemcpy(c64.data(), dC64, c64.size() * 8, cudaMemcpyDeviceToHost));
    err = 0; for (size_t i = 0; i < ref.size(); i++) err = fmax(err, fabs(c64[i] - ref[i]));
    std::printf("DMMA mma.sync.m8n8k4.f64 (8 warps, operands in shared memory, FP64 accumulators in registers): %.0f cycles / GEMM = %.2f us (event: %.2f us); "
                "%.2f TFLOP/s per SM; max |err| / max |C| = %.2e\n",
                (double)cyc[0] / reps, (double)cyc[0] / reps / (clk_khz * 1e-3), ms * 1e3 / reps, flop / ((double)cyc[0] / reps / (clk_khz * 1e3)) / 1e12, err / mx);
    std::printf("SM clock used for the conversion: %.0f MHz\n", clk_khz * 1e-3);
    return 0;
}
