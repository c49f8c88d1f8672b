#!/bin/bash
# ncu evidence for the final round-1 build: launch list (c4-small) and a full capture of the
# face-BVH passes on the 10 M-triangle scene (c5-small, 16 Mi-sample batches like r01_v6)
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
ARGS4="--workload c4-small --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
ARGS5="--workload c5-small --batch 16777216 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 120 python bench.py $ARGS4 > gpurun_out/r1c_c4small.json 2>/dev/null || exit 1
timeout 300 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r01_v11_launches_c4small.csv python bench.py $ARGS4 > gpurun_out/ncu_v11.log 2>&1
timeout 120 python bench.py $ARGS5 > gpurun_out/r1c_c5small.json 2>/dev/null || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_split_mesh -c 4 -f -o gpurun_out/r01_v11_mesh_c5small python bench.py $ARGS5 > gpurun_out/ncu_v11_c5.log 2>&1
ls -la gpurun_out | tail -8
