#!/bin/bash
# Round 2, GPU session L (1 GPU): evidence for the build at the head of the round: parity tests, the default bench
# (C4 + C5 also-leg + e2e + cpu baseline), launch lists of c4-small and c5-small with lanes / issue / DRAM metrics,
# `ncu --set full` captures of the tabulated top-level walk (c4-small), the pair face-BVH pass (c5-small) and the
# shading kernels (c4-small).  ncu reports stay on the box; raw / source pages come back as csv.
mkdir -p gpurun_out
O=gpurun_out
T=r2l
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; tail -3 $O/${T}_gpu_tests.log
RAYITO_B200_TIMING=1 timeout 900 python bench.py > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err; cut -c1-300 $O/${T}_bench_c4.json
export_rep() {   # export_rep /tmp/name tag
  ncu -i $1.ncu-rep --page raw --csv > $O/$2_raw.csv 2>/dev/null
  ncu -i $1.ncu-rep --page source --csv > $O/$2_source.csv 2>/dev/null
  ls -la $1.ncu-rep $O/$2_raw.csv $O/$2_source.csv
}
ARGS4="--workload c4-small --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-also"
ARGS5="--workload c5-small --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-also"
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 200 python bench.py $ARGS4 > $O/${T}_c4small.json 2>/dev/null || exit 1
timeout 500 ncu --metrics $M --clock-control none -c 900 --csv --log-file $O/r02_v3_launches_c4small.csv python bench.py $ARGS4 > $O/ncu_${T}_list4.log 2>&1
timeout 300 python bench.py $ARGS5 > $O/${T}_c5small.json 2>/dev/null || exit 1
timeout 700 ncu --metrics $M --clock-control none -c 900 --csv --log-file $O/r02_v3_launches_c5small.csv python bench.py $ARGS5 > $O/ncu_${T}_list5.log 2>&1
# frame 0 is the counted step (COUNT=1 instantiations): skip its launches, capture the warm-up frame
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_split_top_static -s 9 -c 9 -f -o /tmp/${T}_static python bench.py $ARGS4 > $O/ncu_${T}_static.log 2>&1
export_rep /tmp/${T}_static ${T}_static_c4small
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_split_mesh -s 9 -c 9 -f -o /tmp/${T}_mesh5 python bench.py $ARGS5 > $O/ncu_${T}_mesh5.log 2>&1
export_rep /tmp/${T}_mesh5 ${T}_mesh_c5small
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_split_mesh -s 18 -c 18 -f -o /tmp/${T}_mesh4 python bench.py $ARGS4 > $O/ncu_${T}_mesh4.log 2>&1
export_rep /tmp/${T}_mesh4 ${T}_mesh_c4small
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_shade|k_light_sample|k_resolve|k_raygen" -s 12 -c 14 -f -o /tmp/${T}_shade python bench.py $ARGS4 > $O/ncu_${T}_shade.log 2>&1
export_rep /tmp/${T}_shade ${T}_shade_c4small
ls -la $O | tail -20
