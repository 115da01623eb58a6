#!/bin/bash
# builds experiment variants of the split forward kernel (SVS_F3_EXP bit mask) into tools/bin/ (measurement only)
set -e
cd "$(dirname "$0")/../s-volsdf_b200/csrc"
mkdir -p ../../tools/bin build
for v in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I. -D${SVS_VAR:-SVS_F3_EXP}=$v ${SVS_EXTRA:-} -c mlp.cu -o build/mlpexp$v.obj &
done
wait
for v in "$@"; do
  objs=$(ls build/*.o | grep -v "build/mlp.o" | grep -v obj)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/bin/libsvs_exp$v.so $objs build/mlpexp$v.obj
  echo built exp$v
done
