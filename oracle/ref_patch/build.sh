#!/bin/bash
# Build the reference's own GCLIB_GJKEPA.f90 with gfortran next to the oracle.  UNTESTED in this image (no Fortran
# compiler); see README.md.  Usage: REF=/path/to/collision-detect-GJK-EPA bash oracle/ref_patch/build.sh
set -euo pipefail
here=$(cd "$(dirname "$0")" && pwd)
out="$here/../_ref"
ref=${REF:-/root/reference}
FC=${FC:-gfortran}
command -v "$FC" >/dev/null || { echo "no Fortran compiler ($FC): cannot build the reference" >&2; exit 3; }
mkdir -p "$out"
# a working copy under oracle/_ref (git-ignored); comments are GBK
iconv -f gbk -t utf-8 -c "$ref/src/GCLIB_GJKEPA.f90" > "$out/GCLIB_GJKEPA.f90"
sed -i -E \
    -e 's/isExpa[[:space:]]*==[[:space:]]*\.FALSE\./isExpa .EQV. .FALSE./' \
    -e 's/ANY\(zeroMask\)[[:space:]]*==[[:space:]]*\.FALSE\./ANY(zeroMask) .EQV. .FALSE./' \
    -e 's/idFc\(4, 3\) = \[1, 1, 1, 2, &/idFc(4, 3) = RESHAPE([1, 1, 1, 2, \&/' \
    -e 's/^([[:space:]]*4, 4, 3, 4)\]/\1], [4, 3])/' \
    "$out/GCLIB_GJKEPA.f90"
FFLAGS="-O2 -fopenmp -std=legacy -ffree-line-length-none -ffp-contract=off -fno-fast-math -J$out"
/usr/bin/gcc -O3 -std=c99 -fopenmp -ffp-contract=off -c "$here/../gjkepa_oracle.c" -o "$out/gjkepa_oracle.o"
$FC $FFLAGS -c "$here/gclib_stubs.f90" -o "$out/gclib_stubs.o"
$FC $FFLAGS -c "$out/GCLIB_GJKEPA.f90" -o "$out/GCLIB_GJKEPA.o"
$FC $FFLAGS "$here/ref_driver.f90" "$out/GCLIB_GJKEPA.o" "$out/gclib_stubs.o" "$out/gjkepa_oracle.o" -lm -o "$out/gjkepa_ref"
echo "built $out/gjkepa_ref"
