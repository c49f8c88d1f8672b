#!/bin/bash
# Round 2, GPU session E: parity tests (state shrink, cache heuristic), C5 top-level capture, A/B of pair-pass variants
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2e_gpu_tests.log 2>&1; tail -3 $O/r2e_gpu_tests.log
WORKLOAD=c5-64spp timeout 900 tools/ab_libs.sh 2 base pf s4 s6 adv5 > $O/r2e_ab_c5.log 2>&1; cat $O/r2e_ab_c5.log
WORKLOAD=c4-1080p timeout 600 tools/ab_libs.sh 2 base pf s4 s6 adv5 > $O/r2e_ab_c4.log 2>&1; cat $O/r2e_ab_c4.log
WORKLOAD=c4-1080p timeout 300 tools/ab_env.sh 1 "RAYITO_B200_XFORM_CACHE=none" "RAYITO_B200_XFORM_CACHE=auto" > $O/r2e_ab_cache_c4.log 2>&1; cat $O/r2e_ab_cache_c4.log
WORKLOAD=scene2 STEPS=1 WARMUP=1 timeout 300 tools/ab_env.sh 1 "RAYITO_B200_XFORM_CACHE=none" "RAYITO_B200_XFORM_CACHE=auto" "RAYITO_B200_XFORM_CACHE=rotations" > $O/r2e_ab_cache_s2.log 2>&1; cat $O/r2e_ab_cache_s2.log
ARGS5="--workload c5-small --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-also"
timeout 300 python bench.py $ARGS5 > $O/r2e_c5small.json 2>/dev/null || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_split_top" -s 18 -c 18 -f -o /tmp/r2e_top5 python bench.py $ARGS5 > $O/ncu_r2e_top5.log 2>&1
ncu -i /tmp/r2e_top5.ncu-rep --page raw --csv > $O/r2e_top_c5small_raw.csv 2>/dev/null
ncu -i /tmp/r2e_top5.ncu-rep --page source --csv > $O/r2e_top_c5small_source.csv 2>/dev/null
ls -la $O/r2e_*
