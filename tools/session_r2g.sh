#!/bin/bash
# Round 2, GPU session G: parity tests (working tree), A/B: prev / c3 (staged + packed + one-sector rays) / work / staged0 / direct0
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2g_gpu_tests.log 2>&1; tail -3 $O/r2g_gpu_tests.log
WORKLOAD=c4-1080p timeout 900 tools/ab_libs.sh 2 prev c3 work staged0 direct0 > $O/r2g_ab_c4.log 2>&1; grep "^\[" $O/r2g_ab_c4.log
WORKLOAD=c5-64spp timeout 900 tools/ab_libs.sh 2 prev c3 work staged0 direct0 > $O/r2g_ab_c5.log 2>&1; grep "^\[" $O/r2g_ab_c5.log
