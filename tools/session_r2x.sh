#!/bin/bash
# Round 2, GPU session X (1 GPU): persisting-L2 window over a small scene's arena (RAYITO_B200_L2_PERSIST=1)
mkdir -p gpurun_out
O=gpurun_out
WORKLOAD=c4-1080p timeout 900 tools/ab_env.sh 2 "X=1" "RAYITO_B200_L2_PERSIST=1" > $O/r2x_ab_c4.log 2>&1; cat $O/r2x_ab_c4.log
WORKLOAD=c3 timeout 600 tools/ab_env.sh 1 "X=1" "RAYITO_B200_L2_PERSIST=1" > $O/r2x_ab_c3.log 2>&1; cat $O/r2x_ab_c3.log
