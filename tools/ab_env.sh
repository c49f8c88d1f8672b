#!/bin/bash
# usage: tools/ab_env.sh ROUNDS "ENV_ASSIGNMENTS_1" "ENV_ASSIGNMENTS_2" ...
# Interleaved A/B of run-time switches (one library, different environments), e.g.
#   WORKLOAD=c5-64spp tools/ab_env.sh 2 "RAYITO_B200_NODE_ALIGN=0" "RAYITO_B200_NODE_ALIGN=1"
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for cfg in "$@"; do
    env $cfg python bench.py --workload ${WORKLOAD:-c4-1080p} --steps ${STEPS:-2} --warmup ${WARMUP:-2} --no-e2e --no-cpu-baseline --no-also 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$cfg]', 'Mrays/s %.0f' % d['value'], 'trace %.0f' % r['trace_mrays_per_s_per_gpu'], 'share %.2f' % r['trace_share_of_step'], 'frac %.3f' % r['frac'])"
  done
done
