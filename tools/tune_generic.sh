#!/bin/bash
# usage: tools/tune_generic.sh "-DFOO=1 -DBAR=2" "-DFOO=3" ...   (each argument = one RT_NVCC_EXTRA setting)
for cfg in "$@"; do
  export RT_NVCC_EXTRA="$cfg"
  python -c "from rayito_b200 import build; build.build_core(force=True)" || exit 1
  python bench.py --workload ${WORKLOAD:-c4-1080p} --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-also 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$cfg]', 'Mrays/s %.0f' % d['value'], 'trace Mrays/s %.0f' % r['trace_mrays_per_s_per_gpu'], 'trace share %.2f' % r['trace_share_of_step'])"
done
