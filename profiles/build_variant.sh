#!/bin/bash
# Builds a tuning variant of the library next to the product build: profiles/build_variant.sh <name> [-DFLAG=..]...
# -> conditional_ude_b200/csrc/variants/libcude_b200_<name>.so  (select with CUDE_B200_LIB=...)
set -e
cd "$(dirname "$0")/../conditional_ude_b200/csrc"
name=$1; shift
mkdir -p variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-O2,-Wall -shared -ccbin /usr/bin/g++ --fmad=true "$@" -o variants/libcude_b200_$name.so cude_api.cu
