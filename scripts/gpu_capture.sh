#!/bin/bash
# one `ncu --set full` capture per kernel the bench's roofline quotes traffic for (each after the plain command ran)
mkdir -p gpurun_out
python scripts/prof_run.py disney_spheres mis 64 1 64 fast > /dev/null 2>&1 || exit 1
VIMG_QW_SERIAL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qv_' -s 40 -c 4 \
  -o gpurun_out/prof_r5p_stream -f python scripts/prof_run.py disney_spheres mis 64 1 64 fast > gpurun_out/r5p_ncu1.log 2>&1; tail -1 gpurun_out/r5p_ncu1.log
VIMG_QW_SERIAL=1 SPP_STREAMS=16 FAST=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qw_trace_wide' -s 6 -c 1 \
  -o gpurun_out/prof_r5p_c4wide -f python scripts/bench_synth.py 1000 1366 768 16 > gpurun_out/r5p_ncu2.log 2>&1; tail -1 gpurun_out/r5p_ncu2.log
