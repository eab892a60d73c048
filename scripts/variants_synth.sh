#!/bin/bash
# time every variant on the synthetic 2 M-triangle scene (C4) and on disney_spheres
for lib in v-img_b200/libvimg_cuda.so v-img_b200/variants/*.so; do
  echo "== $lib"
  VIMG_QW_PROFILE=1 VIMG_CUDA_LIB=$PWD/$lib python scripts/bench_synth.py 1000 1366 768 ${1:-8} mis 2>&1 | grep "qw profile" | tail -1 | cut -c1-260
  VIMG_CUDA_LIB=$PWD/$lib python scripts/bench_synth.py 1000 1366 768 ${1:-8} mis 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4 kernel_ms %.2f mrays_s %.1f' % (d['kernel_ms'], d['mrays_s']))"
  VIMG_CUDA_LIB=$PWD/$lib python scripts/prof_run.py disney_spheres mis 64 3 2>&1 | tail -1
done
