#!/bin/bash
# Round 2, GPU session Z (1 GPU): the tree the round ends with -- every GPU test, smoke(), the default bench
mkdir -p gpurun_out
O=gpurun_out
T=r2z
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; tail -3 $O/${T}_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -5 $O/${T}_smoke.log
timeout 900 python bench.py > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err
python -c "import json; d=json.load(open('$O/${T}_bench_c4.json')); r=d['roofline']; a=d['config']['also']['c5']; print('[c4]', 'Mrays/s %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'frac %.3f strict %.3f' % (r['frac'], r['frac_strict']), 'traffic', r['traffic'], 'cpu %.1f' % d['cpu_baseline']['value'], d['clocks'], '| c5 %.0f frac %.3f' % (a['value'], a['roofline']['frac']), 'launches', d['gpu_launches'])"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.json 2>/dev/null; cut -c1-160 $O/${T}_bench_ref.json
