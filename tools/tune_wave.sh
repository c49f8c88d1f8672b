#!/bin/bash
# Rebuild the CUDA core with different wave-traversal scheduling constants and time
# the 1080p workload (run on the GPU box through gpurun).
# usage: tools/tune_wave.sh "REFILL ADVANCE TRI SHAPE" ...
for cfg in "$@"; do
  set -- $cfg
  export RT_NVCC_EXTRA="-DRT_REFILL_MIN=$1 -DRT_ADVANCE_STEPS=$2 -DRT_SERVICE_MIN_TRI=$3 -DRT_SERVICE_MIN_SHAPE=$4"
  python -c "from rayito_b200 import build; build.build_core(force=True)" || exit 1
  python bench.py --workload ${WORKLOAD:-c4-1080p} --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$cfg', 'Mrays/s %.0f' % d['value'], 'trace Mrays/s %.0f' % r['trace_mrays_per_s_per_gpu'], 'trace share %.2f' % r['trace_share_of_step'])"
done
