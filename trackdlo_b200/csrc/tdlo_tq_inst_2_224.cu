// Nn <= 64: 224 threads, 3 resident CTAs per SM
#include "tdlo_tq_inst.cuh"
TDLO_TQ_INSTANCE(tq_2_224_3, 2, 224, 3)
