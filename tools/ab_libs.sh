#!/bin/bash
# usage: tools/ab_libs.sh ROUNDS name1 name2 ...  -- interleaved bench of prebuilt rayito_b200/csrc/_ab/lib_<name>.so
# (WORKLOAD env, default c4-1080p; STEPS / WARMUP env)
rounds=$1; shift
cp rayito_b200/csrc/librayito_b200.so /tmp/librt_keep.so
for r in $(seq 1 $rounds); do
  for name in "$@"; do
    cp rayito_b200/csrc/_ab/lib_$name.so rayito_b200/csrc/librayito_b200.so
    python bench.py --workload ${WORKLOAD:-c4-1080p} --steps ${STEPS:-2} --warmup ${WARMUP:-2} --no-e2e --no-cpu-baseline --no-also 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[${WORKLOAD:-c4-1080p} $name]', 'Mrays/s %.0f' % d['value'], 'trace %.0f' % r['trace_mrays_per_s_per_gpu'], 'share %.3f' % r['trace_share_of_step'], 'frac %.3f' % r['frac'])"
  done
done
cp /tmp/librt_keep.so rayito_b200/csrc/librayito_b200.so
