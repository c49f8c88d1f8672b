#!/bin/bash
# Final validation of the round: GPU parity tests, smoke, default bench, C3 and C2 lines
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r1d_gpu_tests.log 2>&1; tail -3 gpurun_out/r1d_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
RAYITO_B200_TIMING=1 timeout 300 python bench.py > gpurun_out/r1d_bench_c4.json 2> gpurun_out/r1d_bench_c4.err; cut -c1-160 gpurun_out/r1d_bench_c4.json; grep "rayito_b200\] raytrace" gpurun_out/r1d_bench_c4.err | tail -2
timeout 200 python bench.py --workload c3 > gpurun_out/r1d_bench_c3.json 2> gpurun_out/r1d_bench_c3.err; cut -c1-160 gpurun_out/r1d_bench_c3.json
timeout 200 python bench.py --workload c2 > gpurun_out/r1d_bench_c2.json 2> gpurun_out/r1d_bench_c2.err; cut -c1-160 gpurun_out/r1d_bench_c2.json
