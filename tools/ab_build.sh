#!/bin/bash
# usage: tools/ab_build.sh ROUNDS "NVCC_EXTRA_1" "NVCC_EXTRA_2" ...  -- builds one library per setting, then interleaves runs
rounds=$1; shift
i=0
for cfg in "$@"; do
  RT_NVCC_EXTRA="$cfg" python -c "from rayito_b200 import build; build.build_core(force=True)" || exit 1
  cp rayito_b200/csrc/librayito_b200.so /tmp/librt_$i.so
  i=$((i+1))
done
for r in $(seq 1 $rounds); do
  i=0
  for cfg in "$@"; do
    cp /tmp/librt_$i.so rayito_b200/csrc/librayito_b200.so
    python bench.py --workload ${WORKLOAD:-c4-1080p} --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-also 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$cfg]', 'Mrays/s %.0f' % d['value'], 'trace %.0f' % r['trace_mrays_per_s_per_gpu'], 'share %.2f' % r['trace_share_of_step'])"
    i=$((i+1))
  done
done
