#!/bin/bash
# development aid: tools/build_variant.sh <name> [-DMACRO=..]...  ->  tmp_variants/lib_<name>.so (same ABI, loaded with MISHAGPU_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
src=${MG_VARIANT_SRC:-misha_b200/csrc}
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC,-O3 -cudart static \
  --fmad=false --prec-div=true --prec-sqrt=true --ftz=false -Iinclude "$@" -o tmp_variants/lib_$name.so $src/mishagpu.cu
echo built tmp_variants/lib_$name.so
