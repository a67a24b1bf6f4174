// trsm_inst_f64_ptr_left.cu -- one of the eight instantiation units of trsm_dispatch.cuh
#include "trsm_dispatch.cuh"
template int kblasx::tri_solve_side<double, false, true>(KBlasHandle *, int, int, int, double, kblasx::BatchRef<const double, false>, int,
                                                   kblasx::BatchRef<double, false>, int, int);
