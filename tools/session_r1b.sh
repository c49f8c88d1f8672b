#!/bin/bash
# GPU session: parity tests, C5 layout A/B, default bench, C5 end to end with host timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_gpu_tests.log 2>&1; tail -3 gpurun_out/r1b_gpu_tests.log
WORKLOAD=c5-64spp timeout 300 tools/ab_env.sh 2 "RAYITO_B200_NODE_ALIGN=0" "RAYITO_B200_NODE_ALIGN=1" "RAYITO_B200_NODE_ALIGN=1 RAYITO_B200_L2_FETCH=128" > gpurun_out/r1b_ab_c5.log 2>&1
cat gpurun_out/r1b_ab_c5.log
WORKLOAD=c4-1080p timeout 200 tools/ab_env.sh 2 "RAYITO_B200_NODE_ALIGN=0" "RAYITO_B200_NODE_ALIGN=1" > gpurun_out/r1b_ab_c4.log 2>&1
cat gpurun_out/r1b_ab_c4.log
timeout 300 python bench.py > gpurun_out/r1b_bench_c4.json 2> gpurun_out/r1b_bench_c4.err; cat gpurun_out/r1b_bench_c4.json | cut -c1-300
RAYITO_B200_TIMING=1 timeout 300 python bench.py --workload c5-64spp --steps 2 > gpurun_out/r1b_bench_c5.json 2> gpurun_out/r1b_bench_c5.err; cut -c1-300 gpurun_out/r1b_bench_c5.json; grep rayito_b200 gpurun_out/r1b_bench_c5.err | tail -8
