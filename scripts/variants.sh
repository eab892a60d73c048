#!/bin/bash
# time every variant library under v-img_b200/variants on one workload: stage profile (serial) + frame time (concurrent)
#   scripts/variants.sh [scene] [spp] [streams] [fast]
for lib in v-img_b200/libvimg_cuda.so v-img_b200/variants/*.so; do
  echo "== $lib"
  VIMG_QW_PROFILE=1 VIMG_CUDA_LIB=$PWD/$lib python scripts/prof_run.py ${1:-disney_spheres} mis ${2:-128} 1 ${3:-64} ${4:-fast} 2>&1 | grep "qw profile"
  VIMG_CUDA_LIB=$PWD/$lib python scripts/prof_run.py ${1:-disney_spheres} mis ${2:-128} 3 ${3:-64} ${4:-fast} 2>&1 | tail -1
done
