#!/bin/bash
# time every kernel-tuning variant under v-img_b200/variants on one workload (serial profile + concurrent total)
for lib in v-img_b200/libvimg_cuda.so v-img_b200/variants/*.so; do
  echo "== $lib"
  VIMG_QW_PROFILE=1 VIMG_CUDA_LIB=$PWD/$lib python scripts/prof_run.py ${1:-disney_spheres} ${2:-mis} ${3:-16} 1 2>&1 | grep "qw profile"
  VIMG_CUDA_LIB=$PWD/$lib python scripts/prof_run.py ${1:-disney_spheres} ${2:-mis} ${3:-16} 3 2>&1 | tail -1
done
