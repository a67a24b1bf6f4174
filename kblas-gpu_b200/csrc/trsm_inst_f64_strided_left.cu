// trsm_inst_f64_strided_left.cu -- one of the eight instantiation units of trsm_dispatch.cuh
#include "trsm_dispatch.cuh"
template int kblasx::tri_solve_side<double, true, true>(KBlasHandle *, int, int, int, double, kblasx::BatchRef<const double, true>, int,
                                                   kblasx::BatchRef<double, true>, int, int);
