#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/fast_stats.py disney_spheres 512 > gpurun_out/r5g_stats.log 2>&1; cat gpurun_out/r5g_stats.log
timeout 300 python scripts/fast_stats.py cornell_box_spheres 400 >> gpurun_out/r5g_stats.log 2>&1; tail -7 gpurun_out/r5g_stats.log
for spp in 64 256; do
SPP_STREAMS=16 timeout 600 python scripts/bench_synth.py 1000 1366 768 $spp 2>/dev/null | tail -1
SPP_STREAMS=16 FAST=1 timeout 600 python scripts/bench_synth.py 1000 1366 768 $spp 2>/dev/null | tail -1
done > gpurun_out/r5g_c4.log 2>&1; cat gpurun_out/r5g_c4.log
timeout 2400 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r5g_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r5g_tests.log
tail -40 gpurun_out/r5g_tests.log
