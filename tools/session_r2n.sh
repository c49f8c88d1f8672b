#!/bin/bash
# Round 2, GPU session N (1 GPU): mixed host/device tree tests, A/B of the deferred mesh-queue append and the
# step-header prefetch in the tabulated top-level walk
mkdir -p gpurun_out
O=gpurun_out
T=r2n
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; tail -5 $O/${T}_gpu_tests.log
WORKLOAD=c4-1080p timeout 900 tools/ab_libs.sh 2 base defer deferpf > $O/${T}_ab_c4.log 2>&1; grep "^\[" $O/${T}_ab_c4.log
WORKLOAD=c5-64spp timeout 600 tools/ab_libs.sh 1 base defer deferpf > $O/${T}_ab_c5.log 2>&1; grep "^\[" $O/${T}_ab_c5.log
