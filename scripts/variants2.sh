#!/bin/bash
# frame time of every library variant x trace blocks per SM (streams 16)
for lib in v-img_b200/libvimg_cuda.so v-img_b200/variants/*.so; do
  for tb in 2 3 4; do
    echo -n "$(basename $lib) trace_blocks=$tb: "
    VIMG_QW_TRACE_BLOCKS=$tb VIMG_CUDA_LIB=$PWD/$lib python scripts/repeat_run.py disney_spheres ${1:-128} 16 4 | cut -d: -f2
  done
done
