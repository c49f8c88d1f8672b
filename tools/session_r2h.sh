#!/bin/bash
# Round 2, GPU session H: parity tests, A/B prev / rayrec (prev + one-sector ray records) / work / lightgen (no light-sample specialisation)
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2h_gpu_tests.log 2>&1; tail -3 $O/r2h_gpu_tests.log
WORKLOAD=c4-1080p timeout 900 tools/ab_libs.sh 2 prev rayrec work lightgen > $O/r2h_ab_c4.log 2>&1; grep "^\[" $O/r2h_ab_c4.log
WORKLOAD=c5-64spp timeout 900 tools/ab_libs.sh 2 prev rayrec work lightgen > $O/r2h_ab_c5.log 2>&1; grep "^\[" $O/r2h_ab_c5.log
