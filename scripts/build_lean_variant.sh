#!/bin/bash
# build a variant of the LEAN tolerance build (vimg_lean.cu) with other flags; the other objects are reused:
#   scripts/build_lean_variant.sh NAME -DFOO=1 ...   -> v-img_b200/variants/libvimg_NAME.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/v-img_b200/csrc
out=$root/v-img_b200/variants
mkdir -p $out/obj_$name
nvcc -std=c++20 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -use_fast_math -Xcompiler -fPIC,-Wno-unknown-pragmas -Xptxas -v "$@" \
  -c $src/vimg_lean.cu -o $out/obj_$name/vimg_lean.o 2> $out/obj_$name/ptxas.txt || { cat $out/obj_$name/ptxas.txt; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libvimg_$name.so $src/vimg_api.o $src/vimg_fast.o $out/obj_$name/vimg_lean.o $src/vimg_exact_lean.o $src/vimg_scene_io.o
echo "built $name"
