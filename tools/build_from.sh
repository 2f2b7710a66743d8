#!/bin/bash
# builds agarai_b200/_ab/<tag>.so from the working tree with agar_kernels.cuh replaced by <file>
# usage: tools/build_from.sh <kernels.cuh> <tag> [extra nvcc flags]
set -e
F=$1; TAG=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
mkdir -p $T/agarai_b200/csrc $T/include
cp $ROOT/agarai_b200/csrc/*.cu $ROOT/agarai_b200/csrc/*.cuh $T/agarai_b200/csrc/
cp $ROOT/include/*.h $T/include/
cp $F $T/agarai_b200/csrc/agar_kernels.cuh
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared "$@" \
  -o $ROOT/agarai_b200/_ab/$TAG.so $T/agarai_b200/csrc/agar_b200.cu
rm -rf $T
echo built agarai_b200/_ab/$TAG.so
