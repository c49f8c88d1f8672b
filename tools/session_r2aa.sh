#!/bin/bash
# Round 2, GPU session AA (1 GPU): grid sizes cached per device (no occupancy queries at the start of a frame): parity
# tests, per-frame fixed cost, one rank's shard against the whole frame
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_deep.py tests/test_golden.py -m gpu -x -q > $O/r2aa_gpu_tests.log 2>&1; tail -2 $O/r2aa_gpu_tests.log
python tools/frame_fixed_cost.py 2>&1 | tail -5
run() {
  label=$1; shift
  python bench.py --workload c4 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-also "$@" 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$label]', 'ms %.1f' % d['ms_per_step'], 'Mrays/s %.0f' % d['value'], 'trace %.0f' % r['trace_mrays_per_s_per_gpu'])"
}
run "whole frame"
run "0/8" --shard 0/8
run "5/8" --shard 5/8
