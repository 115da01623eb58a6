#!/bin/bash
# Builds measurement variants of the library into tools/bin/ (never shipped): mlp.cu is recompiled with
#   -D${SVS_VAR:-SVS_F3_EXP}=<value> ${SVS_EXTRA}
# for every value given and linked with the objects of the regular build (run s-volsdf_b200/csrc/build.sh first).
#   bash tools/f3_exp.sh 1 16 17                          # SVS_F3_EXP bit masks of tc_fwd3_kernel
#   SVS_VAR=SVS_CHAIN_EXP bash tools/f3_exp.sh 7          # tc_chain_kernel without any copies
#   SVS_VAR=SVS_F3_TRACE bash tools/f3_exp.sh 1           # clock64 trace of the split forward chain (tools/f3_trace.py)
#   SVS_VAR=SVS_CHAIN_TRACE bash tools/f3_exp.sh 5        # ... of the backward sweep (prologue id; tools/chain_trace.py)
set -euo pipefail
cd "$(dirname "$0")/../s-volsdf_b200/csrc"
mkdir -p ../../tools/bin build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
VAR=${SVS_VAR:-SVS_F3_EXP}
pids=()
for v in "$@"; do
  rm -f build/mlpexp$v.obj
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I. -D$VAR=$v ${SVS_EXTRA:-} \
    -c mlp.cu -o build/mlpexp$v.obj > build/mlpexp$v.log 2>&1 & pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait $p || fail=1; done
for v in "$@"; do
  if [ ! -f build/mlpexp$v.obj ]; then
    grep -E "error" -A3 build/mlpexp$v.log | head -20
    echo "FAILED exp$v ($VAR=$v)"
    fail=1
    continue
  fi
  objs=$(ls build/*.o | grep -v "build/mlp.o")
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/bin/libsvs_exp$v.so $objs build/mlpexp$v.obj
  echo "built exp$v ($VAR=$v)"
done
exit $fail
