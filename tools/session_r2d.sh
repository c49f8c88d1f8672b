#!/bin/bash
# Round 2, GPU session D: C5 launch list + full capture of the pair face-BVH pass, A/B of two-leaf parking and cache modes
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_deep.py tests/test_gpu_counters.py tests/test_gpu_render.py -m gpu -x -q > $O/r2d_gpu_tests.log 2>&1; tail -3 $O/r2d_gpu_tests.log
ARGS5="--workload c5-small --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-also"
timeout 300 python bench.py $ARGS5 > $O/r2d_c5small.json 2>/dev/null || exit 1
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/r02_v1_launches_c5small.csv python bench.py $ARGS5 > $O/ncu_r2d_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_split_mesh -s 9 -c 9 -f -o /tmp/r2d_mesh5 python bench.py $ARGS5 > $O/ncu_r2d_mesh5.log 2>&1
ncu -i /tmp/r2d_mesh5.ncu-rep --page raw --csv > $O/r2d_mesh_c5small_raw.csv 2>/dev/null
ncu -i /tmp/r2d_mesh5.ncu-rep --page source --csv > $O/r2d_mesh_c5small_source.csv 2>/dev/null
WORKLOAD=c5-64spp timeout 600 tools/ab_libs.sh 2 one two > $O/r2d_ab_c5.log 2>&1; cat $O/r2d_ab_c5.log
WORKLOAD=c4-1080p timeout 600 tools/ab_libs.sh 2 one two > $O/r2d_ab_c4.log 2>&1; cat $O/r2d_ab_c4.log
WORKLOAD=c4-1080p timeout 600 tools/ab_env.sh 2 "RAYITO_B200_XFORM_CACHE=none" "RAYITO_B200_XFORM_CACHE=rotations" "RAYITO_B200_XFORM_CACHE=all" > $O/r2d_ab_cache_c4.log 2>&1; cat $O/r2d_ab_cache_c4.log
WORKLOAD=c5-64spp timeout 600 tools/ab_env.sh 1 "RAYITO_B200_XFORM_CACHE=none" "RAYITO_B200_XFORM_CACHE=rotations" > $O/r2d_ab_cache_c5.log 2>&1; cat $O/r2d_ab_cache_c5.log
