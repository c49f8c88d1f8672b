#!/bin/bash
# Round 2, GPU session AC (1 GPU): larger grids for the grid-stride kernels (blocks per SM)
mkdir -p gpurun_out
O=gpurun_out
WORKLOAD=c4-1080p timeout 1200 tools/ab_env.sh 2 "X=1" "RAYITO_B200_WIDE_SHADE=48" "RAYITO_B200_WIDE_SHADE=64" "RAYITO_B200_WIDE_SHADE=128" "RAYITO_B200_WIDE_SHADE=512" "RAYITO_B200_WIDE_SHADE=100000" "RAYITO_B200_WIDE_SHADE=128 RAYITO_B200_WIDE_GEN=128" > $O/r2ac_ab_c4.log 2>&1; cat $O/r2ac_ab_c4.log
WORKLOAD=c5-64spp timeout 600 tools/ab_env.sh 1 "X=1" "RAYITO_B200_WIDE_SHADE=64" "RAYITO_B200_WIDE_SHADE=128" "RAYITO_B200_WIDE_SHADE=512" > $O/r2ac_ab_c5.log 2>&1; cat $O/r2ac_ab_c5.log
