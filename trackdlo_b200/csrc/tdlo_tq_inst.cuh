// One translation unit per variant of the persistent task-queue kernel, so that the variants compile in parallel
// (a single TU with all of them takes minutes).  Included by tdlo_tq_inst_*.cu only.
#pragma once
#include "tdlo_taskq.cuh"

#define TDLO_TQ_INSTANCE(NAME, NPASS, THREADS, MINB)                                                              \
    namespace tdlo {                                                                                              \
    cudaError_t NAME##_prepare(int smem, int* occ) {                                                              \
        cudaError_t e = cudaFuncSetAttribute(tdlo_tq_kernel<NPASS, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e != cudaSuccess) return e;                                                                           \
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, tdlo_tq_kernel<NPASS, THREADS, MINB>, THREADS, smem);              \
    }                                                                                                             \
    cudaError_t NAME##_launch(int grid, int smem, cudaStream_t s, const TqArgs& t) {                              \
        tdlo_tq_kernel<NPASS, THREADS, MINB><<<grid, THREADS, smem, s>>>(t);                                      \
        return cudaGetLastError();                                                                                \
    }                                                                                                             \
    }
