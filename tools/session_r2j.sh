#!/bin/bash
# Round 2, GPU session J: parity tests, A/B prev / work2 / work3 (k_shade per shape, pow2 sample split) / shadegen
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2j_gpu_tests.log 2>&1; tail -3 $O/r2j_gpu_tests.log
WORKLOAD=c4-1080p timeout 900 tools/ab_libs.sh 2 prev work2 work3 shadegen > $O/r2j_ab_c4.log 2>&1; grep "^\[" $O/r2j_ab_c4.log
WORKLOAD=c5-64spp timeout 900 tools/ab_libs.sh 1 work2 work3 > $O/r2j_ab_c5.log 2>&1; grep "^\[" $O/r2j_ab_c5.log
