#!/bin/bash
# Builds scratch/lib_<tag>.so: conv_umma.cu recompiled with the given -D flags, linked with the objects of the
# regular build (python -m autoencoder_based_image_compression_b200.build). Use with EAE_LIB_PATH=...
#   scripts/build_variant.sh r96 -DEAE_MAXREGS34=96
set -e
tag=$1; shift
pkg=autoencoder_based_image_compression_b200
mkdir -p scratch
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=true -Xcompiler -fPIC -Xptxas -v -cudart static \
    "$@" -c $pkg/csrc/conv_umma.cu -o scratch/conv_umma_$tag.o 2> scratch/nvcc_$tag.log
nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -Xcompiler -fPIC -o scratch/lib_$tag.so \
    $pkg/build/runtime.o $pkg/build/coder.o $pkg/build/glue.o $pkg/build/transforms_simt.o scratch/conv_umma_$tag.o $pkg/build/codec.o \
    -ldl -lpthread -lrt
grep -A2 "umma[34]_kernel" scratch/nvcc_$tag.log | grep -E "spill|Used"
