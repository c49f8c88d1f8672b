#!/bin/bash
# Round 2, GPU session P (1 GPU): A/B of the gather prefetch in the shading kernels (gpf), of the hoisted record loads
# alone (nogpf), against the head build (base); parity tests on the gpf build first
mkdir -p gpurun_out
O=gpurun_out
T=r2p
cp rayito_b200/csrc/librayito_b200.so /tmp/keep.so; cp rayito_b200/csrc/_ab/lib_gpf.so rayito_b200/csrc/librayito_b200.so
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_golden.py tests/test_gpu_counters.py -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; tail -3 $O/${T}_gpu_tests.log
cp /tmp/keep.so rayito_b200/csrc/librayito_b200.so
WORKLOAD=c4-1080p timeout 900 tools/ab_libs.sh 2 base gpf nogpf > $O/${T}_ab_c4.log 2>&1; grep "^\[" $O/${T}_ab_c4.log
WORKLOAD=c5-64spp timeout 600 tools/ab_libs.sh 1 base gpf nogpf > $O/${T}_ab_c5.log 2>&1; grep "^\[" $O/${T}_ab_c5.log
