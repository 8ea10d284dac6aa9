// Nn <= 64: 256 threads, 2 resident CTAs per SM
#include "tdlo_tq_inst.cuh"
TDLO_TQ_INSTANCE(tq_2_256_2, 2, 256, 2)
