#!/bin/bash
# builds kernel variants into gpurun-travelling .so files: scripts/build_variants.sh name "-DX=1 -DY=0" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags -o variants/lib_$name.so relate_b200/csrc/paint_api.cu &
done
wait
ls -la variants
