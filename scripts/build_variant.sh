#!/bin/bash
# usage: scripts/build_variant.sh <name> <file.cu> [-DMACRO=VALUE ...]   -> variants/lib<name>.so (development sweeps)
set -e
NAME=$1; SRC=$2; shift 2
cd "$(dirname "$0")/../compactbases_jl_b200/csrc"
mkdir -p ../../variants/obj
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c $SRC -o ../../variants/obj/$NAME.o 2>/dev/null
OBJS=""
for f in errors bspline csc apply assemble assemble_stream density bandlu bandlu_part complex comm; do
  if [ "$f.cu" == "$SRC" ]; then OBJS="$OBJS ../../variants/obj/$NAME.o"; else OBJS="$OBJS build/$f.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib$NAME.so $OBJS
echo built variants/lib$NAME.so
