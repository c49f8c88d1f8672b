#!/bin/bash
# usage: tools/ab_args.sh ROUNDS "ARGS1" "ARGS2" ...  -- interleaved bench.py runs differing in command-line arguments
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for a in "$@"; do
    python bench.py --workload ${WORKLOAD:-c4-1080p} --steps 2 --warmup 2 --no-e2e --no-cpu-baseline $a 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$a]', 'Mrays/s %.0f' % d['value'], 'trace %.0f' % r['trace_mrays_per_s_per_gpu'], 'share %.2f' % r['trace_share_of_step'])"
  done
done
