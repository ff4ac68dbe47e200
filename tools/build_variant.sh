#!/bin/bash
# tools/build_variant.sh NAME [nvcc -D flags...]: build an experimental libgca into build/variants/NAME.so
# (select it with GCA_LIB_PATH=build/variants/NAME.so).  build/ is git-ignored but travels with gpurun.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
C=gym_cellular_automata_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o build/variants/$name.so $C/gca_step64.cu $C/gca_bb.cu $C/gca_hidden.cu $C/gca_tiled.cu $C/gca_windy.cu $C/gca_aux.cu $C/gca_abi.cu
echo built build/variants/$name.so
