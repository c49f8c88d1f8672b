#!/bin/bash
# Round 2: config C5 as stated (10 M triangles, 3840x2160, 1024 spp) on N GPUs (first argument), device-timed
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 800 $TR bench.py --gpus $N --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2c5_${N}gpu.json 2> $O/r2c5_${N}gpu.err; tail -2 $O/r2c5_${N}gpu.err
python -c "import json; d=json.load(open('$O/r2c5_${N}gpu.json')); r=d['roofline']; print('[c5 1024spp N=$N]', 'Mrays/s %.0f' % d['value'], 'ms %.1f' % d['ms_per_step'], 'frac %.3f' % r['frac'], d['clocks'])"
