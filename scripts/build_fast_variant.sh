#!/bin/bash
# build a variant of the TOLERANCE build (vimg_fast.cu) with other flags; the bit-exact objects are reused:
#   scripts/build_fast_variant.sh NAME <nvcc flags for vimg_fast.cu ...>
# -> v-img_b200/variants/libvimg_NAME.so (git-ignored, shipped to the GPU box); select with VIMG_CUDA_LIB
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/v-img_b200/csrc
out=$root/v-img_b200/variants
mkdir -p $out/obj_$name
nvcc -std=c++20 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wno-unknown-pragmas -Xptxas -v "$@" \
  -c $src/vimg_fast.cu -o $out/obj_$name/vimg_fast.o 2> $out/obj_$name/ptxas.txt || { cat $out/obj_$name/ptxas.txt; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libvimg_$name.so $src/vimg_api.o $out/obj_$name/vimg_fast.o $src/vimg_lean.o $src/vimg_exact_lean.o $src/vimg_scene_io.o
echo "built $name"
