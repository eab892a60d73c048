#!/bin/bash
mkdir -p gpurun_out
LIBS="v-img_b200/libvimg_cuda.so $(ls v-img_b200/variants/libvimg_*.so)"
timeout 600 python scripts/fast_bias.py disney_spheres 128 v-img_b200/libvimg_cuda.so > gpurun_out/r5c_fused_disney.log 2>&1; cat gpurun_out/r5c_fused_disney.log
STREAMS=64 timeout 900 python scripts/fast_bias.py disney_spheres 128 $LIBS > gpurun_out/r5c_fused_disney64.log 2>&1; cat gpurun_out/r5c_fused_disney64.log
for sc in cornell_box_spheres glass_in_box cornell_skylight sphere_light_small cornell_checker cornell_dof; do
  timeout 300 python scripts/fast_bias.py $sc 100 v-img_b200/libvimg_cuda.so 2>&1 | sed "s/^/$sc /"
done > gpurun_out/r5c_fused_others.log 2>&1; cat gpurun_out/r5c_fused_others.log
VIMG_QW_PROFILE=1 timeout 300 python scripts/prof_run.py disney_spheres mis 128 2 64 fast > gpurun_out/r5c_prof.log 2>&1; cat gpurun_out/r5c_prof.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
