#!/bin/bash
# Build a VARIANT of the library for same-box A/B runs: tools/build_variant.sh <name> [-DFOO=1 ...]
# -> gpurun_variants/lib<name>.so (git-ignored, travels with the gpurun snapshot); select it with TFGNN_B200_LIB=<path>.
set -eu
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OBJ=$(mktemp -d /tmp/tfgnn_variant_${NAME}_XXXX)
mkdir -p $ROOT/gpurun_variants
for f in $ROOT/tf2_gnn_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" \
    -c $f -o $OBJ/$(basename ${f%.cu}).o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a $OBJ/*.o -o $ROOT/gpurun_variants/lib$NAME.so
rm -rf $OBJ
ls -la $ROOT/gpurun_variants/lib$NAME.so
