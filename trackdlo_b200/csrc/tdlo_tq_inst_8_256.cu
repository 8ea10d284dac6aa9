// Nn <= 256: 256 threads, 2 resident CTAs per SM
#include "tdlo_tq_inst.cuh"
TDLO_TQ_INSTANCE(tq_8_256_2, 8, 256, 2)
