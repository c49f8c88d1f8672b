#!/bin/bash
# Round 2, GPU session K (2 GPUs): the product tile assembly (rt_render_multi / raytraceMulti over NCCL) checked
# bit for bit against the single-GPU image, then the default bench (C4 + C5 also-leg + e2e) on 2 GPUs
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multi_check.py > $O/r2k_multi_check.log 2>&1; grep -h "MULTI_CHECK\|Error\|error" $O/r2k_multi_check.log | head -5
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r2k_gpu_multi_tests.log 2>&1; tail -2 $O/r2k_gpu_multi_tests.log
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2k_bench_c4_2gpu.json 2> $O/r2k_bench_c4_2gpu.err; cut -c1-300 $O/r2k_bench_c4_2gpu.json; tail -3 $O/r2k_bench_c4_2gpu.err
