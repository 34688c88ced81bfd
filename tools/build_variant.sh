#!/bin/bash
# A/B builds of the chained-linear kernel: tools/build_variant.sh <name> -DCMDI_CHAIN_...=...  ->
# diffusion-motion-inbetweening_b200/build/variants/lib_<name>.so (run with CONDMDI_B200_LIB=<that path>)
set -e
cd "$(dirname "$0")/../diffusion-motion-inbetweening_b200"
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" \
  -c csrc/gemm_chain.cu -o build/variants/gemm_chain_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so $(ls build/*.o | grep -v gemm_chain.o) build/variants/gemm_chain_$name.o
echo build/variants/lib_$name.so
