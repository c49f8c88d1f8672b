#!/bin/bash
# Round 2, GPU session F: parity tests (staged queue pushes, packed suspended state, one-sector ray records), A/B vs previous build
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2f_gpu_tests.log 2>&1; tail -3 $O/r2f_gpu_tests.log
WORKLOAD=c5-64spp timeout 900 tools/ab_libs.sh 2 prev new > $O/r2f_ab_c5.log 2>&1; cat $O/r2f_ab_c5.log
WORKLOAD=c4-1080p timeout 600 tools/ab_libs.sh 2 prev new > $O/r2f_ab_c4.log 2>&1; cat $O/r2f_ab_c4.log
timeout 600 python bench.py --no-cpu-baseline > $O/r2f_bench_c4.json 2> $O/r2f_bench_c4.err; cut -c1-200 $O/r2f_bench_c4.json
