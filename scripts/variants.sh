#!/bin/bash
# time every kernel-tuning variant under v-img_b200/variants on one workload
for lib in v-img_b200/libvimg_cuda.so v-img_b200/variants/*.so; do
  echo "== $lib"; VIMG_QW_PROFILE=1 VIMG_CUDA_LIB=$PWD/$lib python scripts/prof_run.py ${1:-disney_spheres} ${2:-mis} ${3:-16} 2 2>&1 | tail -2
done
