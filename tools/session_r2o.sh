#!/bin/bash
# Round 2, GPU session O (N GPUs, N = first argument): the default bench (C4 + C5 also-leg + e2e) on N GPUs through the
# product's tile assembly (rt_render_multi / raytraceMulti)
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > $O/r2o_bench_c4_${N}gpu.json 2> $O/r2o_bench_c4_${N}gpu.err; cut -c1-200 $O/r2o_bench_c4_${N}gpu.json; tail -2 $O/r2o_bench_c4_${N}gpu.err
python - <<PY
import json
d=json.load(open("$O/r2o_bench_c4_${N}gpu.json"))
r=d["roofline"]; a=d["config"]["also"]["c5"]; ra=a["roofline"]
print("[N=$N] C4 %.0f Mrays/s (%.1f ms) e2e %.0f frac %.3f avg launch %.3f ms | C5 %.0f Mrays/s frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["avg_launch_ms"], a["value"], ra["frac"]))
print(d["config"]["parallelism"])
PY
