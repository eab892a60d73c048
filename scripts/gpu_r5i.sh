#!/bin/bash
mkdir -p gpurun_out
STREAMS=64 timeout 600 python scripts/fast_bias.py disney_spheres 128 v-img_b200/libvimg_cuda.so 2>&1 | grep "fast_stream\|Error\|error" ; 
timeout 300 python scripts/fast_bias.py cornell_box_spheres 100 v-img_b200/libvimg_cuda.so 2>&1 | grep "fast_stream\|Error\|error"
timeout 300 python scripts/fast_bias.py glass_in_box 100 v-img_b200/libvimg_cuda.so 2>&1 | grep "fast_stream\|Error\|error"
for lib in v-img_b200/libvimg_cuda.so v-img_b200/variants/libvimg_*.so; do
  for ib in 8 4 2; do
    echo "== $lib item_blocks=$ib"
    VIMG_QW_ITEM_BLOCKS=$ib VIMG_CUDA_LIB=$PWD/$lib timeout 300 python scripts/prof_run.py disney_spheres mis 256 3 64 fast 2>&1 | tail -1
  done
  VIMG_QW_PROFILE=1 VIMG_CUDA_LIB=$PWD/$lib timeout 300 python scripts/prof_run.py disney_spheres mis 128 1 64 fast 2>&1 | grep "qw profile"
done > gpurun_out/r5i_variants.log 2>&1; cat gpurun_out/r5i_variants.log
timeout 900 python -m pytest tests/test_gpu_fast.py -q 2>&1 | tail -3
