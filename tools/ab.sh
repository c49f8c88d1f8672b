#!/bin/bash
# usage: tools/ab.sh ROUNDS "ENV1" "ENV2" ...   -- interleaved same-session A/B of bench.py (WORKLOAD env, default c4-1080p)
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for cfg in "$@"; do
    env $cfg python bench.py --workload ${WORKLOAD:-c4-1080p} --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-also 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$cfg]', 'Mrays/s %.0f' % d['value'], 'trace %.0f' % r['trace_mrays_per_s_per_gpu'], 'share %.2f' % r['trace_share_of_step'])"
  done
done
