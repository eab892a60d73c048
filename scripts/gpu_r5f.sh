#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r5f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r5f_tests.log
tail -30 gpurun_out/r5f_tests.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r5f_bench.json 2> gpurun_out/r5f_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r5f_bench.err; head -c 3000 gpurun_out/r5f_bench.json
