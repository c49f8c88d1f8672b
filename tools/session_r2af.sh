#!/bin/bash
# Round 2, GPU session AF (1 GPU): per-bin grid sizes from the previous batches' counts (hint) against the final build (base)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_golden.py tests/test_gpu_counters.py -m gpu -x -q > $O/r2af_gpu_tests.log 2>&1; tail -2 $O/r2af_gpu_tests.log
WORKLOAD=c4 STEPS=3 WARMUP=2 timeout 600 tools/ab_libs.sh 2 base hint > $O/r2af_ab_c4.log 2>&1; grep "^\[" $O/r2af_ab_c4.log
WORKLOAD=c4-small STEPS=5 WARMUP=3 timeout 300 tools/ab_libs.sh 1 base hint > $O/r2af_ab_c4small.log 2>&1; grep "^\[" $O/r2af_ab_c4small.log
