#!/bin/bash
mkdir -p gpurun_out
STREAMS=64 timeout 600 python scripts/fast_bias.py disney_spheres 128 v-img_b200/libvimg_cuda.so > gpurun_out/r5h_disney64.log 2>&1; cat gpurun_out/r5h_disney64.log
timeout 600 python scripts/fast_bias.py disney_spheres 128 v-img_b200/libvimg_cuda.so > gpurun_out/r5h_disney1.log 2>&1; cat gpurun_out/r5h_disney1.log
for sc in cornell_box_spheres glass_in_box cornell_skylight sphere_light_small; do
  STREAMS=8 timeout 300 python scripts/fast_bias.py $sc 100 v-img_b200/libvimg_cuda.so 2>&1 | sed "s/^/$sc /"
done > gpurun_out/r5h_others.log 2>&1; cat gpurun_out/r5h_others.log
VIMG_QW_PROFILE=1 timeout 300 python scripts/prof_run.py disney_spheres mis 128 2 64 fast > gpurun_out/r5h_prof.log 2>&1; cat gpurun_out/r5h_prof.log
timeout 300 python scripts/prof_run.py disney_spheres mis 512 3 64 fast 2>&1 | tail -2
VIMG_QW_SERIAL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qv_' -s 12 -c 4 \
  -o gpurun_out/prof_r5h_stream -f python scripts/prof_run.py disney_spheres mis 64 1 64 fast > gpurun_out/r5h_ncu.log 2>&1
tail -2 gpurun_out/r5h_ncu.log
timeout 900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_streams.py -q 2>&1 | tail -5
