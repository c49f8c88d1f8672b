#!/bin/bash
# Round 2, GPU session V (1 GPU): wavefront state parked whole between scenes (e2e "reserve" phase), every GPU test on
# that build, C3 through the default tree mode, then the L2 fetch granularity A/B of session U
mkdir -p gpurun_out
O=gpurun_out
T=r2v
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; tail -3 $O/${T}_gpu_tests.log
RAYITO_B200_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-also --no-cpu-baseline > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err
python -c "import json; d=json.load(open('$O/${T}_bench_c4.json')); print('[c4]', 'Mrays/s %.0f' % d['value'], 'e2e %.0f (%.1f ms)' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'traffic', d['roofline']['traffic'])"
grep -h "raytrace:" $O/${T}_bench_c4.err | tail -6; grep -h "rt_render host clock" $O/${T}_bench_c4.err | tail -6
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
python -c "import json; d=json.load(open('$O/${T}_bench_c3.json')); print('[c3]', 'Mrays/s %.0f' % d['value'], 'ms %.1f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'], 'cpu', d.get('cpu_baseline', {}).get('value'))"
bash tools/session_r2u.sh
