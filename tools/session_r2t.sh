#!/bin/bash
# Round 2, GPU session T (1 GPU): every rank of an 8-way partition, lattice dealing (tx + 3 ty) mod 8 against the diagonal
# interleave (tx + ty) mod 8 of round 1 (RAYITO_B200_TILE_DEAL=diagonal); max over ranks = the 8-GPU frame time
mkdir -p gpurun_out
O=gpurun_out/r2t_shard.log
: > $O
run() {   # run LABEL ARGS...
  label=$1; shift
  python bench.py --workload c4 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-also "$@" 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$label]', 'rays/step %.0f M' % (d['config']['rays_per_step']/1e6), 'ms %.1f' % d['ms_per_step'], 'Mrays/s %.0f' % d['value'], 'trace %.0f' % r['trace_mrays_per_s_per_gpu'], 'share %.3f' % r['trace_share_of_step'])" | tee -a $O
}
run "whole frame"
for r in 0 1 2 3 4 5 6 7; do run "$r/8 lattice" --shard $r/8; done
export RAYITO_B200_TILE_DEAL=diagonal
for r in 0 1 2 3 4 5 6 7; do run "$r/8 diagonal" --shard $r/8; done
unset RAYITO_B200_TILE_DEAL
for r in 0 1 2 3; do run "$r/4 lattice" --shard $r/4; done
