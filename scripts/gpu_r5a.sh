#!/bin/bash
# round 2, first GPU call: parity suite after the kernel-table refactor, then the tolerance mode's timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r5a_smi.txt
nproc >> gpurun_out/r5a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r5a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r5a_tests.log
tail -3 gpurun_out/r5a_tests.log
timeout 600 python scripts/fast_eval.py disney_spheres 128 64 > gpurun_out/r5a_fast_disney.log 2>&1; cat gpurun_out/r5a_fast_disney.log
timeout 300 python scripts/fast_eval.py cornell_box_spheres 100 50 > gpurun_out/r5a_fast_cornell.log 2>&1; cat gpurun_out/r5a_fast_cornell.log
timeout 300 python scripts/fast_eval.py glass_in_box 200 50 > gpurun_out/r5a_fast_glass.log 2>&1; cat gpurun_out/r5a_fast_glass.log
timeout 300 python scripts/fast_eval.py synthetic_small 64 16 > gpurun_out/r5a_fast_synth.log 2>&1; cat gpurun_out/r5a_fast_synth.log
