#define FIR_TU_VEC 1
#define FIR_TU_NAME(f) f##_r
#include "fir_dg.inc"
