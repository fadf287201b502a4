#!/bin/bash
# Builds scratch/lib_<tag>.so: coder.cu recompiled with the given -D flags, linked with the objects of the
# regular build (python -m autoencoder_based_image_compression_b200.build). Use with EAE_LIB_PATH=...
#   scripts/build_coder_variant.sh c101 -DEAE_CODER_MULHI=1 -DEAE_CODER_KEQA=0 -DEAE_CODER_PHASEB_VEC=1
set -e
tag=$1; shift
pkg=autoencoder_based_image_compression_b200
mkdir -p scratch
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=true -Xcompiler -fPIC -Xptxas -v -cudart static \
    "$@" -c $pkg/csrc/coder.cu -o scratch/coder_$tag.o 2> scratch/nvcc_$tag.log
nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -Xcompiler -fPIC -o scratch/lib_$tag.so \
    $pkg/build/runtime.o scratch/coder_$tag.o $pkg/build/glue.o $pkg/build/transforms_simt.o $pkg/build/conv_umma.o $pkg/build/codec.o \
    -ldl -lpthread -lrt
