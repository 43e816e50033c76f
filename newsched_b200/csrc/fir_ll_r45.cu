#define FIR_TU_VEC 1
#define FIR_TU_NAME(f) f##_r45
#define FIR_LL_EACH(X) X(4, 1) X(4, 3) X(4, 5) X(5, 2) X(5, 3) X(5, 4)
#include "fir_ll.inc"
