#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified reference* hot path into
# oracle/_ref/libkblas_ref.so so that tests/ and bench.py --impl reference can
# run the real KBLAS GPU code beside ours on the B200 box.
#
# Sources are compiled where they lie under /root/reference (never copied into
# the repo).  The only deviation is the one-hunk header patch SURVEY.md §0(8)
# describes: include/kblas_operators.h:126-133 defines
#   void atomicAdd(cuFloatComplex*, cuFloatComplex)
# which CUDA >= 12 rejects (collides with the builtin float2 atomicAdd).  The
# patch is applied to a *copy* of the public headers inside oracle/_ref/inc
# (git-ignored build output).  The hunk is complex-only; s/d paths are untouched.
#
# Usage: oracle/build_ref.sh [REFERENCE_ROOT]     (default /root/reference)
set -euo pipefail
R=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
O=$HERE/_ref
if [ ! -d "$R/src/batch_triangular" ]; then
  echo "build_ref: $R not present; keeping prebuilt $O (if any)"; exit 0
fi
mkdir -p "$O/inc" "$O/obj"
cp "$R"/include/*.h "$O/inc/"
# wrap lines 126-133 (cuFloatComplex atomicAdd) of the COPY in #if 0 / #endif
python3 - "$O/inc/kblas_operators.h" <<'PY'
import sys
p = sys.argv[1]
L = open(p).read().split("\n")
start = next(i for i, l in enumerate(L) if "atomicAdd" in l and "cuFloatComplex" in l)
# walk back to the 'static'/__device__ qualifier line that opens the definition
b = start
while b > 0 and "__device__" not in L[b]:
    b -= 1
# find the closing brace of the function body
depth = 0; e = start; seen = False
for i in range(start, len(L)):
    depth += L[i].count("{"); seen = seen or "{" in L[i]
    depth -= L[i].count("}")
    if seen and depth == 0:
        e = i; break
L.insert(e + 1, "#endif // disabled for CUDA>=12 (oracle/build_ref.sh)")
L.insert(b, "#if 0 // disabled for CUDA>=12 (oracle/build_ref.sh)")
open(p, "w").write("\n".join(L))
print("patched", p, "lines", b + 1, "-", e + 1)
PY
NV="nvcc -O3 -std=c++14 -Xcompiler -fPIC -DTARGET_SM=100 -gencode arch=compute_100,code=sm_100 -I$O/inc -I$R/src -w"
pids=()
for P in s d; do
  for f in Xpotrf_batch Xtrsm_batch Xpotrs_batch Xposv_batch Xsyrk_batch Xgemm_batch Xhelper_funcs; do
    $NV -DPREC_$P -c "$R/src/batch_triangular/$f.cu" -o "$O/obj/$P$f.o" &
    pids+=($!)
  done
done
for f in kblas_common workspace_queries; do
  $NV -c "$R/src/$f.cu" -o "$O/obj/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -shared -o "$O/libkblas_ref.so" "$O"/obj/*.o -lcublas -Xlinker --no-undefined -Xlinker -rpath=/usr/local/cuda/lib64
ls -la "$O/libkblas_ref.so"
