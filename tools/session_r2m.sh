#!/bin/bash
# Round 2, GPU session M (1 GPU): device-side face-BVH build (node-for-node tests), perf-mode tree (measured parity),
# and what each buys end to end on the 10 M-triangle mesh (RAYITO_B200_TIMING=1 prints the phases of raytrace())
mkdir -p gpurun_out
O=gpurun_out
T=r2m
timeout 900 python -m pytest tests/test_gpu_build.py tests/test_perf_tree.py -m gpu -x -q -s > $O/${T}_gpu_tests.log 2>&1; tail -25 $O/${T}_gpu_tests.log
for w in c5-64spp c5-64spp-hostbuild c5-64spp-sah; do
  RAYITO_B200_TIMING=1 timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_$w.json 2> $O/${T}_$w.err
  python -c "import json,sys; d=json.load(open('$O/${T}_$w.json')); r=d['roofline']; print('[$w]', 'Mrays/s %.0f' % d['value'], 'e2e %.0f (%.0f ms)' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'trace %.0f' % r['trace_mrays_per_s_per_gpu'], 'frac %.3f' % r['frac'], 'pops %.2f tris %.2f' % (r['per_ray']['node_pops'], r['per_ray']['tri_tests']))"
  grep "raytrace:\|built on the device\|host clock" $O/${T}_$w.err | tail -4
done
