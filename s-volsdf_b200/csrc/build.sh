#!/bin/bash
# Builds libsvolsdf_b200.so in-tree for sm_100a.  sampler.cu / composite.cu / rays.cu / mvs.cu use -fmad=false so
# the canonical fp32 arithmetic of the oracle (no fused multiply-add) is reproduced exactly.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I. ${SVS_NVCC_EXTRA:-}"
mkdir -p build
pids=()
for f in capi mlp optim; do
  $NVCC $ARCH $COMMON -Xptxas -v -c $f.cu -o build/$f.o > build/$f.log 2>&1 & pids+=($!)
done
for f in sampler composite rays mvs; do
  $NVCC $ARCH $COMMON -fmad=false -Xptxas -v -c $f.cu -o build/$f.o > build/$f.log 2>&1 & pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait $p || fail=1; done
if [ $fail -ne 0 ]; then cat build/*.log; exit 1; fi
$NVCC $ARCH -shared -o ../libsvolsdf_b200.so build/*.o
echo "built $(cd .. && pwd)/libsvolsdf_b200.so"
