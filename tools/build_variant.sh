#!/bin/bash
# builds the library with extra -D flags into build_variants/libhmme_<name>.so (experiments; select with HMME_B200_LIB)
name=$1; shift
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -Xptxas -v "$@" \
  -o build_variants/libhmme_$name.so hm-opencl_b200/csrc/hmme_b200.cu hm-opencl_b200/csrc/hmme_group.cu -ldl 2>&1 | grep -A3 "me_u8_tile_kernelILi3" | grep -E "spill|registers"
