#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG=...]: an experiment build of the library with extra flags for pack_reg_kernels.cu
# into build/variants/lib_NAME.so (B200S_LIB selects it); the other translation units come from build/obj/*.o
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants build/obj
for f in api.cu fit_kernels.cu pack_kernels.cu encoder.cpp; do
  o=build/obj/${f%.*}.o
  if [ ! -f $o ] || [ gpu-provisioner_b200/csrc/$f -nt $o ]; then
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --split-compile 0 -Xcompiler -fPIC -c gpu-provisioner_b200/csrc/$f -o $o
  fi
done
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --split-compile 0 -Xcompiler -fPIC "$@" -c gpu-provisioner_b200/csrc/pack_reg_kernels.cu -o build/variants/pr_$name.o
nvcc -shared -o build/variants/lib_$name.so build/obj/*.o build/variants/pr_$name.o
echo built build/variants/lib_$name.so
