#!/bin/bash
# Round 2, GPU session R (1 GPU): the state the round ends in -- every GPU test, smoke(), the default bench, the other
# configs (C3, C2, scene 2), and the measured DRAM traffic of the traversal launches at the bench's own batch size
mkdir -p gpurun_out
O=gpurun_out
T=r2r
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; tail -3 $O/${T}_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -5 $O/${T}_smoke.log
RAYITO_B200_TIMING=1 timeout 900 python bench.py > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err; cut -c1-300 $O/${T}_bench_c4.json
for w in c3 c2 scene2; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > $O/${T}_bench_$w.json 2> $O/${T}_bench_$w.err
  python -c "import json; d=json.load(open('$O/${T}_bench_$w.json')); print('[$w]', 'Mrays/s %.0f' % d['value'], 'ms %.1f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'], 'cpu', d.get('cpu_baseline', {}).get('value'))"
done
timeout 600 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > $O/${T}_bench_c5_1024spp.json 2> $O/${T}_bench_c5_1024spp.err
python -c "import json; d=json.load(open('$O/${T}_bench_c5_1024spp.json')); print('[c5 1024spp]', 'Mrays/s %.0f' % d['value'], 'ms %.1f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'])"
timeout 120 python bench.py --impl reference --steps 1 --warmup 0 > $O/${T}_bench_ref.json 2>/dev/null; cut -c1-200 $O/${T}_bench_ref.json
# traversal DRAM traffic at the bench's batch size: skip the counted frame's 768 traversal launches, capture the timed frame's
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_split -s 768 -c 768 \
    --csv --log-file $O/r02_v3_traffic_c4.csv python bench.py --workload c4 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-also > $O/ncu_${T}_traffic.log 2>&1
ls -la $O/r02_v3_traffic_c4.csv; tail -2 $O/ncu_${T}_traffic.log
