#!/bin/bash
# per-stage times of every variant library on the synthetic scenes (tolerance mode): scripts/range_synth.sh "G W H spp" ...
for cfg in "$@"; do
  for lib in v-img_b200/libvimg_cuda.so v-img_b200/variants/*.so; do
    echo "== $cfg $lib"
    FAST=1 SPP_STREAMS=16 VIMG_QW_PROFILE=1 VIMG_CUDA_LIB=$PWD/$lib python scripts/bench_synth.py $cfg mis 2>&1 | grep "qw profile" | tail -2 | head -1 | cut -c1-330
  done
done
