#!/bin/bash
# one `ncu --set full` capture per kernel the bench's roofline quotes traffic for (each after the plain command ran)
tag=${1:-r6}
mkdir -p gpurun_out
python scripts/prof_run.py disney_spheres mis 64 1 64 fast > /dev/null 2>&1 || exit 1
# streamed pipeline (tolerance mode): one loaded iteration = shade Principled, shade Lambertian, END
VIMG_QW_SERIAL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qv_' -s 30 -c 3 \
  -o gpurun_out/prof_${tag}_stream -f python scripts/prof_run.py disney_spheres mis 64 1 64 fast > gpurun_out/${tag}_ncu1.log 2>&1; tail -1 gpurun_out/${tag}_ncu1.log
# slot-record pipeline (bit-identical mode): trace, classify, shade x 2, END of one loaded iteration
VIMG_QW_SERIAL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qw_' -s 50 -c 5 \
  -o gpurun_out/prof_${tag}_exact -f python scripts/prof_run.py disney_spheres mis 64 1 1 > gpurun_out/${tag}_ncu3.log 2>&1; tail -1 gpurun_out/${tag}_ncu3.log
# 4-wide BVH trace on the 2 M-triangle scene
VIMG_QW_SERIAL=1 SPP_STREAMS=16 FAST=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qw_trace_wide' -s 6 -c 1 \
  -o gpurun_out/prof_${tag}_c4wide -f python scripts/bench_synth.py 1000 1366 768 16 > gpurun_out/${tag}_ncu2.log 2>&1; tail -1 gpurun_out/${tag}_ncu2.log
