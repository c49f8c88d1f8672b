#!/bin/bash
# Round 2, GPU session Q (1 GPU): what one rank of an 8-GPU run does, on one GPU (bench.py --shard R/8): per-ray cost of
# the diagonal tile interleave at tile sizes 16 / 32 / 64 / 128 against the whole frame, and the spread between ranks
mkdir -p gpurun_out
O=gpurun_out/r2q_shard.log
: > $O
run() {   # run LABEL ARGS...
  label=$1; shift
  python bench.py --workload c4 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-also "$@" 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$label]', 'rays/step %.0f M' % (d['config']['rays_per_step']/1e6), 'ms %.1f' % d['ms_per_step'], 'Mrays/s %.0f' % d['value'], 'trace %.0f' % r['trace_mrays_per_s_per_gpu'], 'share %.3f' % r['trace_share_of_step'], 'avg launch %.3f ms' % r['avg_launch_ms'], 'launches %d' % r['launches_per_step'])" | tee -a $O
}
run "whole frame"
run "0/8 tile 32" --shard 0/8
run "3/8 tile 32" --shard 3/8
run "5/8 tile 32" --shard 5/8
run "0/8 tile 16" --shard 0/8 --tile 16
run "0/8 tile 64" --shard 0/8 --tile 64
run "0/8 tile 128" --shard 0/8 --tile 128
run "0/8 tile 32 batch 16Mi" --shard 0/8 --batch 16777216
run "whole frame batch 16Mi" --batch 16777216
run "0/2 tile 32" --shard 0/2
