#!/bin/bash
mkdir -p gpurun_out
STREAMS=64 timeout 600 python scripts/fast_bias.py disney_spheres 128 v-img_b200/libvimg_cuda.so > gpurun_out/r5e_disney64.log 2>&1; cat gpurun_out/r5e_disney64.log
timeout 600 python scripts/fast_bias.py disney_spheres 128 v-img_b200/libvimg_cuda.so > gpurun_out/r5e_disney1.log 2>&1; cat gpurun_out/r5e_disney1.log
for sc in cornell_box_spheres glass_in_box cornell_skylight sphere_light_small sphere_light_medium; do
  timeout 300 python scripts/fast_bias.py $sc 100 v-img_b200/libvimg_cuda.so 2>&1 | sed "s/^/$sc /"
done > gpurun_out/r5e_others.log 2>&1; cat gpurun_out/r5e_others.log
VIMG_QW_PROFILE=1 timeout 300 python scripts/prof_run.py disney_spheres mis 128 2 64 fast > gpurun_out/r5e_prof.log 2>&1; cat gpurun_out/r5e_prof.log
VIMG_QW_SERIAL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qf_trace_flat|qw_end' -s 8 -c 2 \
  -o gpurun_out/prof_r5e_flat -f python scripts/prof_run.py disney_spheres mis 64 1 64 fast > gpurun_out/r5e_ncu.log 2>&1
tail -2 gpurun_out/r5e_ncu.log
