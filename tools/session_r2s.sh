#!/bin/bash
# Round 2, GPU session S (1 GPU): every rank of an 8-way partition at tile sizes 32 / 128 (max over ranks = the 8-GPU
# frame time), and the whole frame at tile sizes 64 / 128
mkdir -p gpurun_out
O=gpurun_out/r2s_shard.log
: > $O
run() {   # run LABEL ARGS...
  label=$1; shift
  python bench.py --workload c4 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-also "$@" 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$label]', 'rays/step %.0f M' % (d['config']['rays_per_step']/1e6), 'ms %.1f' % d['ms_per_step'], 'Mrays/s %.0f' % d['value'], 'trace %.0f' % r['trace_mrays_per_s_per_gpu'], 'share %.3f' % r['trace_share_of_step'])" | tee -a $O
}
run "whole frame tile 64" --tile 64
run "whole frame tile 128" --tile 128
run "whole frame tile 256" --tile 256
for r in 0 1 2 3 4 5 6 7; do run "$r/8 tile 128" --shard $r/8 --tile 128; done
for r in 1 2 4 6 7; do run "$r/8 tile 32" --shard $r/8; done
for r in 0 3 6; do run "$r/8 tile 256" --shard $r/8 --tile 256; done
