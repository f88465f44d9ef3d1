#!/bin/bash
# build_variant.sh NAME [extra nvcc flags...] — a variant of the library whose march kernels (engine.cu) are compiled with
# extra flags (e.g. -DTM_VARIANT=2), linked against the other objects of the regular build, into tools/_variant/NAME/.
# Select it at run time with SCFTB_LIB=tools/_variant/NAME/libscft_b200.so (A/B measurements of kernel experiments).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
OUT=$ROOT/tools/_variant/$NAME
mkdir -p $OUT
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -rdc=false -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v --fmad=true "$@" -c -o $OUT/engine.o $ROOT/scft_b200/csrc/engine.cu 2> $OUT/engine.ptxas.log || (cat $OUT/engine.ptxas.log; exit 1)
OBJS=""
for u in solvers mixer postproc pcg2d broyden_dev diblock; do OBJS="$OBJS $ROOT/scft_b200/lib/obj/$u.o"; done
$NVCC $ARCH -shared -o $OUT/libscft_b200.so $OUT/engine.o $OBJS
grep -A3 "march_tm_kernel" $OUT/engine.ptxas.log | grep -E "spill|Used" | sed "s/^/[$NAME] /"
