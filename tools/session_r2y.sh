#!/bin/bash
# Round 2, GPU session Y (1 GPU): L2 prefetch of the pair behind the one being expanded (the near child's own children
# when the near child is the left one) in the face-BVH pass
mkdir -p gpurun_out
O=gpurun_out
WORKLOAD=c5-64spp timeout 900 tools/ab_libs.sh 2 base pfnext > $O/r2y_ab_c5.log 2>&1; grep "^\[" $O/r2y_ab_c5.log
WORKLOAD=c4-1080p timeout 600 tools/ab_libs.sh 1 base pfnext > $O/r2y_ab_c4.log 2>&1; grep "^\[" $O/r2y_ab_c4.log
