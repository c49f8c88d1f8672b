#!/bin/bash
# Round 2, GPU session B: parity tests on the sibling-pair face-BVH pass, then same-session A/B of
# the old pass and pair variants on C5 (64 spp) and C4 (1080p)
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2b_gpu_tests.log 2>&1; tail -4 $O/r2b_gpu_tests.log
WORKLOAD=c5-64spp timeout 900 tools/ab_libs.sh 2 old pair3 pair2 pair5 pair3s12 pair3svc8 > $O/r2b_ab_c5.log 2>&1; cat $O/r2b_ab_c5.log
WORKLOAD=c4-1080p timeout 600 tools/ab_libs.sh 2 old pair3 pair2 pair5 pair3s12 pair3svc8 > $O/r2b_ab_c4.log 2>&1; cat $O/r2b_ab_c4.log
