#!/bin/bash
# Round 2, GPU session I: parity tests, A/B prev / rayrec / work1 (64-byte packed record) / work2 (128-byte record by sectors + pow2 CMJ) / nopow2
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2i_gpu_tests.log 2>&1; tail -3 $O/r2i_gpu_tests.log
WORKLOAD=c4-1080p timeout 900 tools/ab_libs.sh 2 prev rayrec work1 work2 nopow2 > $O/r2i_ab_c4.log 2>&1; grep "^\[" $O/r2i_ab_c4.log
WORKLOAD=c5-64spp timeout 900 tools/ab_libs.sh 2 prev work1 work2 > $O/r2i_ab_c5.log 2>&1; grep "^\[" $O/r2i_ab_c5.log
