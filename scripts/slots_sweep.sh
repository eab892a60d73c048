#!/bin/bash
# frame time of the streamed pipeline against the size and number of its slot pools: scripts/slots_sweep.sh [lib] [spp] "slots..." "pools..."
lib=${1:-v-img_b200/libvimg_cuda.so}; spp=${2:-512}
for slots in ${3:-1048576 2097152 4194304 8388608 16777216}; do
  for pools in ${4:-2}; do
    echo -n "slots=$slots pools=$pools: "
    VIMG_QW_SLOTS=$slots VIMG_QW_POOLS=$pools VIMG_CUDA_LIB=$PWD/$lib python scripts/prof_run.py disney_spheres mis $spp 3 64 fast 2>&1 | tail -1
  done
done
