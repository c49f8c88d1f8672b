#!/bin/bash
# Round 2, GPU session A: parity tests (new deep-tree and counter tests), default bench on the
# round-1 kernels (baseline of this round, with the C5 also-leg), and source-level ncu captures of
# the final round-1 build: tabulated top-level walk, face-BVH pass (C4 and C5), shading kernels.
# ncu reports stay on the box; their raw / source pages come back as csv.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2a_gpu_tests.log 2>&1; tail -4 $O/r2a_gpu_tests.log
RAYITO_B200_TIMING=1 timeout 600 python bench.py > $O/r2a_bench_c4.json 2> $O/r2a_bench_c4.err; cut -c1-400 $O/r2a_bench_c4.json

export_rep() {   # export_rep /tmp/name tag
  ncu -i $1.ncu-rep --page raw --csv > $O/$2_raw.csv 2>/dev/null
  ncu -i $1.ncu-rep --page source --csv > $O/$2_source.csv 2>/dev/null
  ls -la $1.ncu-rep $O/$2_raw.csv $O/$2_source.csv
}
ARGS4="--workload c4-small --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-also"
ARGS5="--workload c5-small --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-also"
timeout 200 python bench.py $ARGS4 > $O/r2a_c4small.json 2>/dev/null || exit 1
# frame 0 is the counted step (COUNT=1 instantiations): skip its launches, capture the warm-up frame
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_split_top_static -s 9 -c 9 -f -o /tmp/r2a_static python bench.py $ARGS4 > $O/ncu_r2a_static.log 2>&1
export_rep /tmp/r2a_static r2a_static_c4small
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_split_mesh -s 18 -c 18 -f -o /tmp/r2a_mesh4 python bench.py $ARGS4 > $O/ncu_r2a_mesh4.log 2>&1
export_rep /tmp/r2a_mesh4 r2a_mesh_c4small
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_shade|k_light_sample|k_resolve|k_raygen" -s 10 -c 10 -f -o /tmp/r2a_shade python bench.py $ARGS4 > $O/ncu_r2a_shade.log 2>&1
export_rep /tmp/r2a_shade r2a_shade_c4small
timeout 300 python bench.py $ARGS5 > $O/r2a_c5small.json 2>/dev/null || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_split_mesh -s 9 -c 9 -f -o /tmp/r2a_mesh5 python bench.py $ARGS5 > $O/ncu_r2a_mesh5.log 2>&1
export_rep /tmp/r2a_mesh5 r2a_mesh_c5small
# launch list of one c4-small frame with lanes / issue metrics (all kernels)
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/r02_v0_launches_c4small.csv python bench.py $ARGS4 > $O/ncu_r2a_list.log 2>&1
ls -la $O | tail -20
