// trsm_inst_f64_ptr_right.cu -- one of the eight instantiation units of trsm_dispatch.cuh
#include "trsm_dispatch.cuh"
template int kblasx::tri_solve_side<double, false, false>(KBlasHandle *, int, int, int, double, kblasx::BatchRef<const double, false>, int,
                                                   kblasx::BatchRef<double, false>, int, int);
