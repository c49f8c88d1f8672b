#!/bin/bash
# Round 2, GPU session C: parity tests on pair pass + per-sample transform cache, then A/B of the cache
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c_gpu_tests.log 2>&1; tail -4 $O/r2c_gpu_tests.log
WORKLOAD=c4-1080p timeout 600 tools/ab_env.sh 2 "RAYITO_B200_NO_XFORM_CACHE=1" "X=1" > $O/r2c_ab_c4.log 2>&1; cat $O/r2c_ab_c4.log
WORKLOAD=c5-64spp timeout 600 tools/ab_env.sh 1 "RAYITO_B200_NO_XFORM_CACHE=1" "X=1" > $O/r2c_ab_c5.log 2>&1; cat $O/r2c_ab_c5.log
WORKLOAD=scene2 STEPS=1 WARMUP=1 timeout 600 tools/ab_env.sh 1 "RAYITO_B200_NO_XFORM_CACHE=1" "X=1" > $O/r2c_ab_s2.log 2>&1; cat $O/r2c_ab_s2.log
