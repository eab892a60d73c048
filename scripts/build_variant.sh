#!/bin/bash
# build a tuning variant of the bit-exact build (vimg_api.cu); the tolerance build is reused: scripts/build_variant.sh NAME -DFOO=1 ...
# -> v-img_b200/variants/libvimg_NAME.so  (git-ignored, shipped to the GPU box)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/v-img_b200/csrc
out=$root/v-img_b200/variants
mkdir -p $out/obj_$name
nvcc -std=c++20 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true \
  -ftz=false -Xcompiler -fPIC,-ffp-contract=off,-Wno-unknown-pragmas -Xptxas -v "$@" \
  -c $src/vimg_api.cu -o $out/obj_$name/vimg_api.o 2> $out/obj_$name/ptxas.txt
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libvimg_$name.so $out/obj_$name/vimg_api.o $src/vimg_fast.o $src/vimg_scene_io.o
grep -A1 "qw_" $out/obj_$name/ptxas.txt | grep -c spill >/dev/null
echo "built $name"
