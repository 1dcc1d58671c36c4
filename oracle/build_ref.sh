#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds the *unmodified* reference CPU path
# (/root/reference/src/cpu + src/interface_c, foges/pogs @ 649ba26) into
# oracle/_ref/libpogs_ref.so, straight from the sources where they lie.  No
# reference source is copied into this repo; only the built .so lands in the
# git-ignored oracle/_ref/ directory (it travels to the GPU box with gpurun).
#
# The reference needs a CBLAS (it calls cblas_{s,d}{axpy,dot,nrm2,scal,gemv,
# trsv,syrk,trsm,gemm} and {s,d}syevd_).  This image has no system BLAS, so
# we link the OpenBLAS bundled with scipy (symbols carry a "scipy_" prefix,
# remapped with -D flags below) or, as a fallback, the one bundled with
# opencv (plain symbols).
#
# Usage: oracle/build_ref.sh [--openmp]     (default: both variants are built)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${POGS_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/cpu" ]; then
  echo "build_ref.sh: reference tree not found at $REF (expected on the GPU box) - skipping" >&2
  exit 0
fi
mkdir -p "$OUT"
SP="$(python3 -c 'import site;print(site.getsitepackages()[0])')"
SCIPY_BLAS="$(ls "$SP"/scipy.libs/libscipy_openblas-*.so 2>/dev/null | head -1 || true)"
CV_BLAS="$(ls "$SP"/opencv_python_headless.libs/libopenblasp-*.so 2>/dev/null | head -1 || true)"
DEFS=()
if [ -n "$SCIPY_BLAS" ]; then
  BLAS="$SCIPY_BLAS"
  for s in saxpy daxpy sscal dscal sasum dasum sdot ddot snrm2 dnrm2 sgemv dgemv \
           strsv dtrsv ssyrk dsyrk sgemm dgemm strsm dtrsm; do
    DEFS+=("-Dcblas_$s=scipy_cblas_$s")
  done
  for s in ssyevd_ dsyevd_ strtrs_ dtrtrs_ sgeqrf_ dgeqrf_ sormqr_ dormqr_; do
    DEFS+=("-D$s=scipy_$s")
  done
elif [ -n "$CV_BLAS" ]; then
  BLAS="$CV_BLAS"
else
  echo "build_ref.sh: no bundled OpenBLAS found" >&2; exit 1
fi
BLASDIR="$(dirname "$BLAS")"
SRCS=("$REF/src/cpu/pogs.cpp" "$REF/src/cpu/matrix/matrix_dense.cpp"
      "$REF/src/cpu/matrix/matrix_sparse.cpp"
      "$REF/src/cpu/projector/projector_cgls.cpp"
      "$REF/src/cpu/projector/projector_direct_dense.cpp"
      "$REF/src/interface_c/pogs_c.cpp"
      "$HERE/ref_persistent.cpp")   # our own driver over the reference's C++ API (no reference code)
INC=(-I "$REF/src/include" -I "$REF/src/cpu/include" -I "$REF/src/interface_c")
build() {  # $1 = output name, rest = extra flags
  local out="$1"; shift
  g++ -O2 -std=c++20 -fPIC -shared "$@" "${DEFS[@]}" "${INC[@]}" "${SRCS[@]}" \
      -o "$OUT/$out" "$BLAS" -Wl,-rpath,"$BLASDIR"
  echo "built $OUT/$out against $BLAS"
}
# Serial-loop build (BLAS threads only) and OpenMP build (prox/equil loops threaded).
build libpogs_ref.so &
build libpogs_ref_omp.so -fopenmp &
wait
echo "$BLAS" > "$OUT/BLAS_USED.txt"
