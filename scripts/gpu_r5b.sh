#!/bin/bash
mkdir -p gpurun_out
LIBS="v-img_b200/libvimg_cuda.so $(ls v-img_b200/variants/libvimg_*.so)"
timeout 900 python scripts/fast_bias.py disney_spheres 128 $LIBS > gpurun_out/r5b_bias_disney.log 2>&1; cat gpurun_out/r5b_bias_disney.log
timeout 600 python scripts/fast_bias.py cornell_box_spheres 100 $LIBS > gpurun_out/r5b_bias_cornell.log 2>&1; cat gpurun_out/r5b_bias_cornell.log
for v in g_l3 h_l4; do
  VIMG_CUDA_LIB=$PWD/v-img_b200/variants/libvimg_$v.so timeout 300 python scripts/prof_run.py disney_spheres mis 128 3 64 fast 2>&1 | tail -1
done > gpurun_out/r5b_occ.log 2>&1; cat gpurun_out/r5b_occ.log
