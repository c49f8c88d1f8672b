#!/bin/bash
# Round 2, GPU session AB (1 GPU): grid size of the grid-stride (shading / ray generation) kernels in blocks per SM: 32
# (5.3 waves of the 6 resident 80-register blocks) against whole numbers of waves
mkdir -p gpurun_out
O=gpurun_out
WORKLOAD=c4-1080p timeout 1200 tools/ab_env.sh 1 "X=1" "RAYITO_B200_WIDE_SHADE=6" "RAYITO_B200_WIDE_SHADE=12" "RAYITO_B200_WIDE_SHADE=18" "RAYITO_B200_WIDE_SHADE=24" "RAYITO_B200_WIDE_SHADE=48" "X=2" "RAYITO_B200_WIDE_GEN=10" "RAYITO_B200_WIDE_GEN=20" "RAYITO_B200_WIDE_GEN=30" > $O/r2ab_ab_c4.log 2>&1; cat $O/r2ab_ab_c4.log
