// trsm_inst_f32_ptr_left.cu -- one of the eight instantiation units of trsm_dispatch.cuh
#include "trsm_dispatch.cuh"
template int kblasx::tri_solve_side<float, false, true>(KBlasHandle *, int, int, int, float, kblasx::BatchRef<const float, false>, int,
                                                   kblasx::BatchRef<float, false>, int, int);
