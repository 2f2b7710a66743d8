#!/bin/bash
# Builds the CUDA library of a git revision (or of the working tree: WORK) into agarai_b200/_ab/<tag>.so so that
# two kernels can be timed on the same GPU box in one gpurun call (AGAR_B200_LIB=... python bench.py).
# usage: tools/build_variant.sh <rev|WORK> <tag> [extra nvcc flags]
set -e
REV=$1; TAG=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
mkdir -p $T/agarai_b200/csrc $T/include
for f in agarai_b200/csrc/agar_b200.cu agarai_b200/csrc/agar_kernels.cuh agarai_b200/csrc/agar_features.cuh include/agar_b200.h include/agar_spec.h; do
  if [ "$REV" = WORK ]; then cp $ROOT/$f $T/$f; else git -C $ROOT show $REV:$f > $T/$f 2>/dev/null || true; fi
done
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared "$@" \
  -o $ROOT/agarai_b200/_ab/$TAG.so $T/agarai_b200/csrc/agar_b200.cu
rm -rf $T
echo built agarai_b200/_ab/$TAG.so
