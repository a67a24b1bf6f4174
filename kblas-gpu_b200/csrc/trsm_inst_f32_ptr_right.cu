// trsm_inst_f32_ptr_right.cu -- one of the eight instantiation units of trsm_dispatch.cuh
#include "trsm_dispatch.cuh"
template int kblasx::tri_solve_side<float, false, false>(KBlasHandle *, int, int, int, float, kblasx::BatchRef<const float, false>, int,
                                                   kblasx::BatchRef<float, false>, int, int);
