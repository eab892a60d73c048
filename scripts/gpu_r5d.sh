#!/bin/bash
mkdir -p gpurun_out
VIMG_QW_SERIAL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qf_path|qw_shade' -s 24 -c 3 \
  -o gpurun_out/prof_r5d_fused -f python scripts/prof_run.py disney_spheres mis 64 1 64 fast > gpurun_out/r5d_ncu.log 2>&1
tail -5 gpurun_out/r5d_ncu.log
VIMG_FLAT_MODE=1 VIMG_QW_SERIAL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qf_trace_flat' -s 8 -c 1 \
  -o gpurun_out/prof_r5d_flat -f python scripts/prof_run.py disney_spheres mis 64 1 64 fast > gpurun_out/r5d_ncu2.log 2>&1
tail -3 gpurun_out/r5d_ncu2.log
