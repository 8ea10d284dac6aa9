// Nn <= 128: 256 threads, 2 resident CTAs per SM
#include "tdlo_tq_inst.cuh"
TDLO_TQ_INSTANCE(tq_4_256_2, 4, 256, 2)
