#!/bin/bash
# GPU session: parity tests, default bench (C4) with host timing, C5 at 64 spp and at the full 1024 spp
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r1c_gpu_tests.log 2>&1; tail -3 gpurun_out/r1c_gpu_tests.log
RAYITO_B200_TIMING=1 timeout 300 python bench.py > gpurun_out/r1c_bench_c4.json 2> gpurun_out/r1c_bench_c4.err; cut -c1-200 gpurun_out/r1c_bench_c4.json; grep "rayito_b200\]" gpurun_out/r1c_bench_c4.err | tail -4
RAYITO_B200_TIMING=1 timeout 300 python bench.py --workload c5-64spp --steps 2 > gpurun_out/r1c_bench_c5_64.json 2> gpurun_out/r1c_bench_c5_64.err; cut -c1-200 gpurun_out/r1c_bench_c5_64.json; grep "rayito_b200\]" gpurun_out/r1c_bench_c5_64.err | tail -4
timeout 400 python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_c5_full.json 2> gpurun_out/r1c_bench_c5_full.err; cut -c1-200 gpurun_out/r1c_bench_c5_full.json
