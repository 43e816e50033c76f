#define FIR_TU_VEC 1
#define FIR_TU_NAME(f) f##_r23
#define FIR_LL_EACH(X) X(2, 1) X(3, 1) X(2, 3) X(2, 5) X(3, 2) X(3, 4) X(3, 5)
#include "fir_ll.inc"
