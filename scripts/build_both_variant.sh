#!/bin/bash
# build a variant of BOTH builds (bit-exact vimg_api.cu and tolerance vimg_fast.cu) with the same -D flags:
#   scripts/build_both_variant.sh NAME -DFOO=1 ...   -> v-img_b200/variants/libvimg_NAME.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/v-img_b200/csrc
out=$root/v-img_b200/variants
mkdir -p $out/obj_$name
common="-std=c++20 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xptxas -v"
nvcc $common -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off,-Wno-unknown-pragmas "$@" \
  -c $src/vimg_api.cu -o $out/obj_$name/vimg_api.o 2> $out/obj_$name/ptxas_api.txt &
nvcc $common -use_fast_math -Xcompiler -fPIC,-Wno-unknown-pragmas "$@" \
  -c $src/vimg_fast.cu -o $out/obj_$name/vimg_fast.o 2> $out/obj_$name/ptxas_fast.txt &
nvcc $common -use_fast_math -Xcompiler -fPIC,-Wno-unknown-pragmas "$@" \
  -c $src/vimg_lean.cu -o $out/obj_$name/vimg_lean.o 2> $out/obj_$name/ptxas_lean.txt &
nvcc $common -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off,-Wno-unknown-pragmas "$@" \
  -c $src/vimg_exact_lean.cu -o $out/obj_$name/vimg_exact_lean.o 2> $out/obj_$name/ptxas_exact_lean.txt &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libvimg_$name.so $out/obj_$name/vimg_api.o $out/obj_$name/vimg_fast.o $out/obj_$name/vimg_lean.o $out/obj_$name/vimg_exact_lean.o $src/vimg_scene_io.o
echo "built $name"
