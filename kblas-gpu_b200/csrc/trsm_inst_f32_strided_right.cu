// trsm_inst_f32_strided_right.cu -- one of the eight instantiation units of trsm_dispatch.cuh
#include "trsm_dispatch.cuh"
template int kblasx::tri_solve_side<float, true, false>(KBlasHandle *, int, int, int, float, kblasx::BatchRef<const float, true>, int,
                                                   kblasx::BatchRef<float, true>, int, int);
