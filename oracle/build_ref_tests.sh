#!/usr/bin/env bash
# oracle/build_ref_tests.sh -- TEST INFRASTRUCTURE.  Builds the reference's OWN hot-path test/bench programs
# (testing/batch_triangular/test_X{potrf,trsm,potrs,posv}_batch.cpp + testing/testing_helper.cu) from the sources
# where they lie under /root/reference, unmodified, against the reference's headers, and links them TWICE:
#   oracle/_ref/bin/ours/test_<p><op>_batch  ->  kblas-gpu_b200/lib/libkblas-gpu.so   (the drop-in proof)
#   oracle/_ref/bin/ref/test_<p><op>_batch   ->  oracle/_ref/libkblas_ref.so          (the reference itself)
# The image has no MKL/LAPACKE headers: oracle/shim/mkl.h forwards the five BLAS/LAPACK names the programs call to
# scipy's bundled OpenBLAS (they are compiled -DUSE_MKL, as the authors build them: make.inc:28-29).
# Outputs only under oracle/_ref/ (git-ignored, travels to the GPU box).  Nothing is copied from the reference.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
R=${KBLAS_REFERENCE:-/root/reference}
OUT="$HERE/_ref"
[ -d "$R/testing/batch_triangular" ] || { echo "no reference tree at $R"; exit 0; }
[ -f "$OUT/libkblas_ref.so" ] || "$HERE/build_ref.sh"
mkdir -p "$OUT/obj" "$OUT/bin/ours" "$OUT/bin/ref"
SCIPY_LIBS=$(python -c "import scipy,os;print(os.path.realpath(os.path.join(os.path.dirname(scipy.__file__),'..','scipy.libs')))")
OPENBLAS=$(basename "$(ls "$SCIPY_LIBS"/libscipy_openblas*.so | head -1)")
CUDA=${CUDA_HOME:-/usr/local/cuda}
INC="-I$HERE/shim -I$R/testing -I$OUT/inc -I$R/src -I$CUDA/include"
nvcc -O2 -std=c++14 -Xcompiler -fopenmp,-fPIC -DUSE_MKL -gencode arch=compute_100,code=sm_100 $INC \
     -c "$R/testing/testing_helper.cu" -o "$OUT/obj/testing_helper.o"
for op in potrf trsm potrs posv; do
  for p in s d; do
    # -O0: the test functions are declared int and fall off their end without a return (test_Xpotrf_batch.cpp:413);
    # with optimisation g++ treats that as unreachable and main runs on into kblasDestroy a second time (SIGSEGV in
    # cublasDestroy with the reference library); at -O0 a trap (SIGILL) sits there -- after all output, either way
    g++ -O0 -fopenmp -w -DUSE_MKL -DPREC_$p $INC -c "$R/testing/batch_triangular/test_X${op}_batch.cpp" \
        -o "$OUT/obj/test_${p}${op}_batch.o" &
  done
done
wait
COMMON="-L$CUDA/lib64 -lcublas -lcudart -lcusolver -L$SCIPY_LIBS -l:$OPENBLAS -Wl,-rpath,$SCIPY_LIBS -lgomp -lm"
for op in potrf trsm potrs posv; do
  for p in s d; do
    o="$OUT/obj/test_${p}${op}_batch.o"
    g++ -fopenmp "$o" "$OUT/obj/testing_helper.o" -o "$OUT/bin/ours/test_${p}${op}_batch" \
        -L"$ROOT/kblas-gpu_b200/lib" -l:libkblas-gpu.so -Wl,-rpath,'$ORIGIN/../../../../kblas-gpu_b200/lib' $COMMON
    g++ -fopenmp "$o" "$OUT/obj/testing_helper.o" -o "$OUT/bin/ref/test_${p}${op}_batch" \
        -L"$OUT" -l:libkblas_ref.so -Wl,-rpath,'$ORIGIN/../..' $COMMON
  done
done
ls -la "$OUT/bin/ours" "$OUT/bin/ref"
