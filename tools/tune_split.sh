#!/bin/bash
# usage: tools/tune_split.sh "TOP_REFILL TOP_ADV TOP_SVC MESH_REFILL MESH_ADV MESH_SVC" ...
for cfg in "$@"; do
  set -- $cfg
  export RT_NVCC_EXTRA="-DRT_TOP_REFILL_MIN=$1 -DRT_TOP_ADVANCE_STEPS=$2 -DRT_TOP_SERVICE_MIN=$3 -DRT_MESH_REFILL_MIN=$4 -DRT_MESH_ADVANCE_STEPS=$5 -DRT_MESH_SERVICE_MIN=$6"
  python -c "from rayito_b200 import build; build.build_core(force=True)" || exit 1
  python bench.py --workload ${WORKLOAD:-c4-1080p} --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$cfg', 'Mrays/s %.0f' % d['value'], 'trace Mrays/s %.0f' % r['trace_mrays_per_s_per_gpu'], 'trace share %.2f' % r['trace_share_of_step'])"
done
