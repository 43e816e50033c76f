#define FIR_TU_VEC 2
#define FIR_TU_NAME(f) f##_c
#include "fir_dg.inc"
