#!/bin/bash
# Round 2, GPU session AE (1 GPU): launch list of the FINAL build (v4: grid change included), c4-small
mkdir -p gpurun_out
O=gpurun_out
ARGS4="--workload c4-small --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-also"
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 200 python bench.py $ARGS4 > $O/r2ae_c4small.json 2>/dev/null || exit 1
timeout 500 ncu --metrics $M --clock-control none -c 900 --csv --log-file $O/r02_v4_launches_c4small.csv python bench.py $ARGS4 > $O/ncu_r2ae_list4.log 2>&1
ls -la $O/r02_v4_launches_c4small.csv; python -c "import json; d=json.load(open('$O/r2ae_c4small.json')); print('c4-small live', d['ms_per_step'], d['value'], d['roofline']['trace_share_of_step'])"
